// sfgs_activations.cu — the per-Gaussian activations the reference evaluates with torch ops in front of every
// render() call, as one kernel forward and one backward:
//
//   scales    = sqrt(exp(s)^2 + f^2)                              get_scaling_with_3D_filter, scene/gaussian_model.py:207-213
//   opacity   = sigmoid(o) * sqrt(prod exp(s)^2 / prod(exp(s)^2 + f^2))   get_opacity_with_3D_filter, :237-249
//   rotations = q / max(|q|, 1e-12)                               get_rotation, :215-217 (F.normalize, :89)
//
// Precision follows the reference's dtype promotion: the parameters are float32, `filter_3D` is float64
// (compute_3D_filter, :254-308), so exp / square / sigmoid / prod(exp(s)^2) are float32 operations and
// everything that touches f is float64, cast to float32 at the end (gaussian_renderer/__init__.py:137-138).
// The reference spends ~14 elementwise kernels forward and ~25 backward on this; both kernels here are pure
// bandwidth (72 B and 136 B per Gaussian).
#include "sfgs_common.cuh"

namespace {

constexpr int ACT_THREADS = 256;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(ACT_THREADS)
activations_fwd_kernel(int P, const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
                       const float* __restrict__ rotation_raw, const double* __restrict__ filter_3D,
                       float* __restrict__ opacity, float* __restrict__ scales, float* __restrict__ rotations) {
  const int i = blockIdx.x * ACT_THREADS + threadIdx.x;
  if (i >= P) return;
  const double f = filter_3D[i], f2 = f * f;
  float sq[3];
  double a[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float s = expf(scaling_raw[3 * i + k]);
    sq[k] = s * s;
    a[k] = (double)sq[k] + f2;
    scales[3 * i + k] = (float)sqrt(a[k]);
  }
  const float det1 = sq[0] * sq[1] * sq[2];          // float32 product, like torch.prod on a float32 tensor
  const double det2 = a[0] * a[1] * a[2];
  const double coef = sqrt((double)det1 / det2);
  opacity[i] = (float)((double)sigmoid_f(opacity_raw[i]) * coef);
  const float4 q = *reinterpret_cast<const float4*>(rotation_raw + 4 * (size_t)i);
  const float d = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
  *reinterpret_cast<float4*>(rotations + 4 * (size_t)i) = make_float4(q.x / d, q.y / d, q.z / d, q.w / d);
}

__global__ void __launch_bounds__(ACT_THREADS)
activations_bwd_kernel(int P, const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
                       const float* __restrict__ rotation_raw, const double* __restrict__ filter_3D,
                       const float* __restrict__ g_opacity, const float* __restrict__ g_scales,
                       const float* __restrict__ g_rotations, float* __restrict__ g_opacity_raw,
                       float* __restrict__ g_scaling_raw, float* __restrict__ g_rotation_raw) {
  const int i = blockIdx.x * ACT_THREADS + threadIdx.x;
  if (i >= P) return;
  const double f = filter_3D[i], f2 = f * f;
  float s[3], sq[3];
  double a[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s[k] = expf(scaling_raw[3 * i + k]);
    sq[k] = s[k] * s[k];
    a[k] = (double)sq[k] + f2;
  }
  const float det1 = sq[0] * sq[1] * sq[2];
  const double det2 = a[0] * a[1] * a[2];
  const double coef = sqrt((double)det1 / det2);
  const float sg = sigmoid_f(opacity_raw[i]);
  const double go = (double)g_opacity[i];
  // opacity = sg * coef
  g_opacity_raw[i] = (float)(go * coef) * sg * (1.0f - sg);
  // d coef / d sq_k = coef/2 * (1/sq_k - 1/a_k);   d sq_k / d raw_k = 2 sq_k
  // scales_k = sqrt(a_k):  d scales_k / d sq_k = 1 / (2 sqrt(a_k))
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double g_sq = go * (double)sg * coef * 0.5 * (1.0 / (double)sq[k] - 1.0 / a[k]) +
                        (double)g_scales[3 * i + k] * 0.5 / sqrt(a[k]);
    g_scaling_raw[3 * i + k] = (float)(g_sq * 2.0 * (double)sq[k]);
  }
  // rotations = q / d, d = max(|q|, eps): g/d - [|q| >= eps] (g.q)/d^2 * q/|q|
  const float4 q = *reinterpret_cast<const float4*>(rotation_raw + 4 * (size_t)i);
  const float4 g = *reinterpret_cast<const float4*>(g_rotations + 4 * (size_t)i);
  const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  const float d = fmaxf(n, 1e-12f);
  float4 r = make_float4(g.x / d, g.y / d, g.z / d, g.w / d);
  if (n >= 1e-12f) {
    const float t = (g.x * q.x + g.y * q.y + g.z * q.z + g.w * q.w) / (d * d) / n;
    r.x -= t * q.x; r.y -= t * q.y; r.z -= t * q.z; r.w -= t * q.w;
  }
  *reinterpret_cast<float4*>(g_rotation_raw + 4 * (size_t)i) = r;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int sfgs_launch_activations_fwd(int P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                                const double* filter_3D, float* opacity, float* scales, float* rotations,
                                cudaStream_t st) {
  if (!aligned16(rotation_raw) || !aligned16(rotations)) return 1;
  SFGS_COUNT_LAUNCH();
  activations_fwd_kernel<<<(P + ACT_THREADS - 1) / ACT_THREADS, ACT_THREADS, 0, st>>>(
      P, opacity_raw, scaling_raw, rotation_raw, filter_3D, opacity, scales, rotations);
  return 0;
}

int sfgs_launch_activations_bwd(int P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                                const double* filter_3D, const float* g_opacity, const float* g_scales,
                                const float* g_rotations, float* g_opacity_raw, float* g_scaling_raw,
                                float* g_rotation_raw, cudaStream_t st) {
  if (!aligned16(rotation_raw) || !aligned16(g_rotations) || !aligned16(g_rotation_raw)) return 1;
  SFGS_COUNT_LAUNCH();
  activations_bwd_kernel<<<(P + ACT_THREADS - 1) / ACT_THREADS, ACT_THREADS, 0, st>>>(
      P, opacity_raw, scaling_raw, rotation_raw, filter_3D, g_opacity, g_scales, g_rotations, g_opacity_raw,
      g_scaling_raw, g_rotation_raw);
  return 0;
}
