// sfgs_knn.cu — mean squared distance to the 3 nearest neighbours of every point.
//
// Replaces SimpleKNN::knn (KNN/simple_knn.cu:186-222: Morton sort + 1024-point boxes + box rejection,
// with thrust allocations, cudaMalloc/cudaFree and two blocking cudaMemcpy per call).  The result is the
// exact 3-NN set, so any exact search gives the same floats; here the points are binned into a uniform
// grid (about two points per cell, dimensions chosen on the device from the bounding box — no host
// round trip), and each point searches rings of cells until the third-best distance is provably final.
// Everything runs on the caller's stream from one scratch allocation.
#include <cfloat>
#include "sfgs_common.cuh"

namespace {

struct KnnParams {
  float minx, miny, minz;
  float h, inv_h;
  int nx, ny, nz;
  unsigned ncells;
  float margin;
};

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void knn_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;        // min (ordered ints)
  else if (threadIdx.x < 6) bbox[threadIdx.x] = (int)0x80000000;   // max
}

__global__ void __launch_bounds__(256)
knn_bbox_kernel(int P, const float* __restrict__ pts, int* __restrict__ bbox) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
    for (int k = 0; k < 3; k++) { const float v = pts[3 * (size_t)i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { atomicMin(&bbox[k], f2ord(mn[k])); atomicMax(&bbox[3 + k], f2ord(mx[k])); }
  }
}

// choose the grid on the device: cubic cells, ~2 points per cell, at most `cell_cap` cells
__global__ void knn_params_kernel(int P, const int* __restrict__ bbox, unsigned cell_cap, KnnParams* prm) {
  if (threadIdx.x != 0) return;
  const float mn[3] = {ord2f(bbox[0]), ord2f(bbox[1]), ord2f(bbox[2])};
  const float mx[3] = {ord2f(bbox[3]), ord2f(bbox[4]), ord2f(bbox[5])};
  float ext[3], emax = 0.f;
  for (int k = 0; k < 3; k++) { ext[k] = fmaxf(mx[k] - mn[k], 0.f); emax = fmaxf(emax, ext[k]); }
  if (!(emax > 0.f)) emax = 1.f;
  const float target_cells = fminf(fmaxf((float)P * 0.5f, 1.f), (float)cell_cap);
  // volume of the box with degenerate axes replaced by one cell's thickness; solve for h iteratively
  float h = emax;
  for (int it = 0; it < 40; it++) {
    double cells = 1.0;
    for (int k = 0; k < 3; k++) cells *= floor((double)ext[k] / h) + 1.0;
    if (cells >= target_cells) break;
    h *= 0.8f;
  }
  int n[3];
  for (;;) {
    double cells = 1.0;
    for (int k = 0; k < 3; k++) { n[k] = (int)floor((double)ext[k] / h) + 1; cells *= n[k]; }
    if (cells <= (double)cell_cap) break;
    h *= 1.25f;
  }
  prm->minx = mn[0]; prm->miny = mn[1]; prm->minz = mn[2];
  prm->h = h; prm->inv_h = 1.0f / h;
  prm->nx = n[0]; prm->ny = n[1]; prm->nz = n[2];
  prm->ncells = (unsigned)n[0] * (unsigned)n[1] * (unsigned)n[2];
  prm->margin = 1e-4f * h + 1e-6f * emax;
}

__device__ __forceinline__ void cell_of(const KnnParams& p, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = min(p.nx - 1, max(0, (int)((x - p.minx) * p.inv_h)));
  cy = min(p.ny - 1, max(0, (int)((y - p.miny) * p.inv_h)));
  cz = min(p.nz - 1, max(0, (int)((z - p.minz) * p.inv_h)));
}

__global__ void __launch_bounds__(256)
knn_count_kernel(int P, const float* __restrict__ pts, const KnnParams* __restrict__ prm, uint32_t* __restrict__ cell_count,
                 uint32_t* __restrict__ cell_of_point) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const KnnParams p = *prm;
  int cx, cy, cz;
  cell_of(p, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], cx, cy, cz);
  const uint32_t c = ((uint32_t)cz * p.ny + cy) * p.nx + cx;
  cell_of_point[i] = c;
  atomicAdd(&cell_count[c], 1u);
}

// exclusive scan of cell_count[0..ncells) into cell_start[0..ncells], single CTA
__global__ void __launch_bounds__(1024)
knn_scan_kernel(const KnnParams* __restrict__ prm, const uint32_t* __restrict__ cell_count, uint32_t* __restrict__ cell_start,
                uint32_t* __restrict__ cell_cursor) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const unsigned n = prm->ncells;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (unsigned base = 0; base < n; base += 1024) {
    const unsigned i = base + tid;
    const uint32_t v = i < n ? cell_count[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t incl = carry_s + x + (wid > 0 ? warp_sums[wid - 1] : 0u);
    if (i < n) { cell_start[i] = incl - v; cell_cursor[i] = 0u; }
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) cell_start[n] = carry_s;
}

__global__ void __launch_bounds__(256)
knn_fill_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ cell_of_point,
                const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_cursor, float4* __restrict__ sorted) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const uint32_t c = cell_of_point[i];
  const uint32_t slot = cell_start[c] + atomicAdd(&cell_cursor[c], 1u);
  sorted[slot] = make_float4(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], __int_as_float(i));
}

__device__ __forceinline__ void update3(float dist, float best[3]) {
  // same insertion network as updateKBest<3> (KNN/simple_knn.cu:131-146)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (best[j] > dist) { const float t = best[j]; best[j] = dist; dist = t; }
  }
}

__global__ void __launch_bounds__(128)
knn_query_kernel(int P, const KnnParams* __restrict__ prm, const uint32_t* __restrict__ cell_start,
                 const float4* __restrict__ sorted, float* __restrict__ out) {
  const int s = blockIdx.x * 128 + threadIdx.x;   // iterate in cell order for locality
  if (s >= P) return;
  const KnnParams p = *prm;
  const float4 me = sorted[s];
  const int my_id = __float_as_int(me.w);
  int cx, cy, cz;
  cell_of(p, me.x, me.y, me.z, cx, cy, cz);
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  const int rmax = max(max(p.nx, p.ny), p.nz);
  for (int r = 0; r <= rmax; r++) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, p.nz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, p.ny - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, p.nx - 1);
    for (int z = z0; z <= z1; z++) {
      const bool zface = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; y++) {
        const bool yface = (y == cy - r) || (y == cy + r);
        const uint32_t row = ((uint32_t)z * p.ny + y) * p.nx;
        if (zface || yface) {
          // whole x-run of this row belongs to the shell: cells are contiguous in memory
          const uint32_t a = cell_start[row + x0], b = cell_start[row + x1 + 1];
          for (uint32_t k = a; k < b; k++) {
            const float4 q = sorted[k];
            if (__float_as_int(q.w) == my_id) continue;
            const float dx = q.x - me.x, dy = q.y - me.y, dz = q.z - me.z;
            update3(dx * dx + dy * dy + dz * dz, best);
          }
        } else {
          // interior row: only the two end cells are on the shell
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const int x = side == 0 ? cx - r : cx + r;
            if (x < 0 || x >= p.nx || (side == 1 && r == 0)) continue;
            const uint32_t a = cell_start[row + x], b = cell_start[row + x + 1];
            for (uint32_t k = a; k < b; k++) {
              const float4 q = sorted[k];
              if (__float_as_int(q.w) == my_id) continue;
              const float dx = q.x - me.x, dy = q.y - me.y, dz = q.z - me.z;
              update3(dx * dx + dy * dy + dz * dz, best);
            }
          }
        }
      }
    }
    // every unvisited point lies outside the cube of cells [c-r, c+r]^3: lower-bound its distance
    float bound = FLT_MAX;
    if (cx - r > 0) bound = fminf(bound, me.x - (p.minx + (float)(cx - r) * p.h));
    if (cx + r + 1 < p.nx) bound = fminf(bound, (p.minx + (float)(cx + r + 1) * p.h) - me.x);
    if (cy - r > 0) bound = fminf(bound, me.y - (p.miny + (float)(cy - r) * p.h));
    if (cy + r + 1 < p.ny) bound = fminf(bound, (p.miny + (float)(cy + r + 1) * p.h) - me.y);
    if (cz - r > 0) bound = fminf(bound, me.z - (p.minz + (float)(cz - r) * p.h));
    if (cz + r + 1 < p.nz) bound = fminf(bound, (p.minz + (float)(cz + r + 1) * p.h) - me.z);
    if (bound == FLT_MAX) break;                       // the cube covers the whole grid
    bound = fmaxf(bound - p.margin, 0.f);
    if (best[2] <= bound * bound) break;
  }
  out[my_id] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace

extern "C" int sfgs_set_error(int code, const char* what, int cuda_error);   // sfgs_api.cu

extern "C" int sfgs_dist2_knn3(int P, const float* points, float* mean_dist2, sfgs_alloc_fn scratch_alloc,
                               void* scratch_user, void* stream) {
  if (P < 0) return sfgs_set_error(SFGS_E_BADARG, "dist2_knn3: P < 0", 0);
  if (P == 0) return SFGS_OK;
  if (!points || !mean_dist2 || !scratch_alloc) return sfgs_set_error(SFGS_E_BADARG, "dist2_knn3: null pointer", 0);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long cap = (unsigned long long)P * 2ull;
  if (cap < 4096ull) cap = 4096ull;
  if (cap > (1ull << 24)) cap = 1ull << 24;
  const unsigned cell_cap = (unsigned)cap;
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = sfgs_align_up(off + bytes); return o; };
  const size_t o_bbox = carve(8 * sizeof(int)), o_prm = carve(sizeof(KnnParams));
  const size_t o_count = carve(((size_t)cell_cap + 1) * 4), o_start = carve(((size_t)cell_cap + 1) * 4);
  const size_t o_cursor = carve(((size_t)cell_cap + 1) * 4), o_cop = carve((size_t)P * 4);
  const size_t o_sorted = carve((size_t)P * sizeof(float4));
  char* raw = scratch_alloc(scratch_user, off + SFGS_ALIGN);
  if (!raw) return sfgs_set_error(SFGS_E_ALLOC, "dist2_knn3: scratch allocator returned NULL", 0);
  char* base = sfgs_align_ptr(raw);
  int* bbox = (int*)(base + o_bbox);
  KnnParams* prm = (KnnParams*)(base + o_prm);
  uint32_t* cell_count = (uint32_t*)(base + o_count);
  uint32_t* cell_start = (uint32_t*)(base + o_start);
  uint32_t* cell_cursor = (uint32_t*)(base + o_cursor);
  uint32_t* cop = (uint32_t*)(base + o_cop);
  float4* sorted = (float4*)(base + o_sorted);

  if (cudaMemsetAsync(cell_count, 0, ((size_t)cell_cap + 1) * 4, st) != cudaSuccess)
    return sfgs_set_error(SFGS_E_CUDA, "dist2_knn3: memset", (int)cudaGetLastError());
  const int pb = (P + 255) / 256;
  SFGS_COUNT_LAUNCH(); knn_init_kernel<<<1, 32, 0, st>>>(bbox);
  SFGS_COUNT_LAUNCH(); knn_bbox_kernel<<<pb < 592 ? pb : 592, 256, 0, st>>>(P, points, bbox);
  SFGS_COUNT_LAUNCH(); knn_params_kernel<<<1, 32, 0, st>>>(P, bbox, cell_cap, prm);
  SFGS_COUNT_LAUNCH(); knn_count_kernel<<<pb, 256, 0, st>>>(P, points, prm, cell_count, cop);
  SFGS_COUNT_LAUNCH(); knn_scan_kernel<<<1, 1024, 0, st>>>(prm, cell_count, cell_start, cell_cursor);
  SFGS_COUNT_LAUNCH(); knn_fill_kernel<<<pb, 256, 0, st>>>(P, points, cop, cell_start, cell_cursor, sorted);
  SFGS_COUNT_LAUNCH(); knn_query_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, prm, cell_start, sorted, mean_dist2);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SFGS_OK : sfgs_set_error(SFGS_E_CUDA, "dist2_knn3", (int)e);
}
