// oracle/ref_glm_fix.h — TEST INFRASTRUCTURE; force-included (nvcc -include) when oracle/build_ref.py
// compiles the reference's own .cu files.  It does not change any reference source.
//
// Why it exists: the glm snapshot vendored by the reference declares the functor
//     template<typename T> struct TMax { T operator()(const T& a, const T& b) { return max(a, b); } };
// (third_party/glm/glm/detail/func_common.inl:37-40) without GLM_FUNC_QUALIFIER, i.e. host-only.
// computeColorFromSH ends with `glm::max(result, 0.0f)` (RAST/cuda_rasterizer/forward.cu:70), which
// calls that functor from device code.  nvcc 12.9 only warns (#20011-D) and then treats the call as
// undefined: the three `rgb[idx*C+ch] = ...` stores of preprocessCUDA vanish from the PTX, so the
// reference binary built with this toolchain renders the SH path with uninitialised colours.
// The explicit specialisation below has the same body, merely callable on the device, so the
// reference computes what its source says.  (The colors_precomp path never touches TMax.)
#pragma once
#include <cuda.h>
#include "cuda_runtime.h"
#define GLM_FORCE_CUDA
#include <glm/glm.hpp>
namespace glm {
template <>
struct TMax<float> {
  __host__ __device__ float operator()(const float& a, const float& b) { return (a < b) ? b : a; }
};
}  // namespace glm
