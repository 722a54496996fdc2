// sfgs_filter3d.cu — GaussianModel.compute_3D_filter as one pass over the Gaussians.
//
// The reference (scene/gaussian_model.py:254-308) loops over every training camera in Python and, per camera, runs
// ~15 float64 torch kernels over all P Gaussians (transform, norm, clamp, project, four comparisons, masked min, or):
// C cameras x 15 kernels x P x 8-byte temporaries, every 100 iterations.  Here one thread owns one Gaussian, keeps its
// position in registers and walks the camera table (144 bytes per camera, read through the constant/L1 path by the
// whole warp at once): P x 12 bytes in, P x 8 bytes out, no temporaries.  Arithmetic is float64 like the reference's.
//
//   per camera:  p = xyz @ R + T;  valid = p.z > 0.2 and -0.15 w <= x/z fx + cx <= 1.15 w and likewise in y  (z clamped at 0.001)
//                distance = min(distance, z) over valid cameras
//   afterwards:  distance of never-valid Gaussians = max distance of the valid ones;  filter_3D = distance / max focal * sqrt(0.2)
#include "sfgs_common.cuh"

namespace {

constexpr int CAM_DOUBLES = 18;   // R[9] (row-major, camera.R), T[3], focal_x, focal_y, cx_ori, cy_ori, width, height

__global__ void __launch_bounds__(256)
filter3d_distance_kernel(int P, const float* __restrict__ xyz, int C, const double* __restrict__ cams,
                         double* __restrict__ dist /* [P]: min z over valid cameras, or -1 */,
                         unsigned long long* __restrict__ max_bits) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  double best = 1e8;
  bool any = false;
  if (i < P) {
    const double x = (double)xyz[3 * (size_t)i], y = (double)xyz[3 * (size_t)i + 1], z = (double)xyz[3 * (size_t)i + 2];
    for (int c = 0; c < C; c++) {
      const double* K = cams + (size_t)c * CAM_DOUBLES;
      // xyz @ R + T  (row vector times matrix: out[j] = sum_i xyz[i] R[i][j])
      const double cx_ = x * K[0] + y * K[3] + z * K[6] + K[9];
      const double cy_ = x * K[1] + y * K[4] + z * K[7] + K[10];
      const double cz_ = x * K[2] + y * K[5] + z * K[8] + K[11];
      const bool valid_depth = cz_ > 0.2;
      const double zc = cz_ < 0.001 ? 0.001 : cz_;
      const double sx = cx_ / zc * K[12] + K[14];
      const double sy = cy_ / zc * K[13] + K[15];
      const double w = K[16], h = K[17];
      const bool in_screen = sx >= -0.15 * w && sx <= w * 1.15 && sy >= -0.15 * h && sy <= 1.15 * h;
      if (valid_depth && in_screen) { best = zc < best ? zc : best; any = true; }
    }
    dist[i] = any ? best : -1.0;
  }
  // block maximum of the valid distances -> one atomic per block (positive doubles order like their bit patterns)
  double m = (i < P && any) ? best : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m; }
  __shared__ double s_m[8];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mm = 0.0;
    for (int k = 0; k < 8; k++) mm = s_m[k] > mm ? s_m[k] : mm;
    if (mm > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(mm));
  }
}

__global__ void __launch_bounds__(256)
filter3d_finish_kernel(int P, double* __restrict__ dist, const unsigned long long* __restrict__ max_bits, double focal_max) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const double dmax = __longlong_as_double((long long)*max_bits);
  const double d = dist[i];
  dist[i] = (d < 0.0 ? dmax : d) / focal_max * sqrt(0.2);     // python: distance / focal_length * (0.2 ** 0.5)
}

}  // namespace

extern "C" int sfgs_set_error(int code, const char* what, int cuda_error);   // sfgs_api.cu

// cams: DEVICE array [C][18] doubles (layout above); focal_max: the largest focal_x over the cameras (host value);
// filter_3D: [P] doubles (the reference's [P,1] float64 tensor); scratch8: 8 bytes of device memory.
extern "C" int sfgs_compute_3d_filter(int P, const float* xyz, int C, const double* cams, double focal_max,
                                      double* filter_3D, void* scratch8, void* stream) {
  if (P < 0 || C < 0) return sfgs_set_error(SFGS_E_BADARG, "compute_3d_filter: negative size", 0);
  if (P == 0) return SFGS_OK;
  if (!xyz || !filter_3D || !scratch8 || (C > 0 && !cams)) return sfgs_set_error(SFGS_E_BADARG, "compute_3d_filter: null pointer", 0);
  if (!(focal_max > 0.0)) return sfgs_set_error(SFGS_E_BADARG, "compute_3d_filter: focal_max must be positive", 0);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch8, 0, 8, st);
  if (e != cudaSuccess) return sfgs_set_error(SFGS_E_CUDA, "compute_3d_filter: memset", (int)e);
  const int blocks = (P + 255) / 256;
  SFGS_COUNT_LAUNCH();
  filter3d_distance_kernel<<<blocks, 256, 0, st>>>(P, xyz, C, cams, filter_3D, (unsigned long long*)scratch8);
  SFGS_COUNT_LAUNCH();
  filter3d_finish_kernel<<<blocks, 256, 0, st>>>(P, filter_3D, (const unsigned long long*)scratch8, focal_max);
  e = cudaGetLastError();
  return e == cudaSuccess ? SFGS_OK : sfgs_set_error(SFGS_E_CUDA, "compute_3d_filter", (int)e);
}
