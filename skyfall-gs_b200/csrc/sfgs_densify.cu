// sfgs_densify.cu — adaptive density control of the training loop as three passes.
//
// Reference: train.py:311-322 (per-iteration statistics, every iteration below densify_until_iter) and
// scene/gaussian_model.py:564-742 (densify_and_prune every densification_interval iterations).
//
//  * statistics  — the reference indexes five [P] tensors with a boolean mask four times per iteration (each a nonzero +
//    gather + scatter, i.e. a host synchronisation and ~20 launches).  Here: one kernel, one thread per Gaussian, in place.
//  * densify_and_prune — the reference builds the clone set, concatenates every parameter AND both Adam moments,
//    builds the split set, concatenates again, filters the split sources out, then filters again by opacity / size:
//    four full rewrites of ~177 floats per Gaussian plus ~60 small launches.  The outcome for every ORIGINAL Gaussian
//    is decided by its own values alone (its mean gradient, largest scale, opacity), so here one pass writes an action
//    byte per Gaussian and per-1024 counts, one block scans the counts, and one pass moves every surviving row —
//    originals with their Adam moments, clones and split children with zero moments — to its final place.  Row order is
//    the reference's: [surviving originals][surviving clones][surviving children, first sample][..., second sample].
//
// Arithmetic follows what the reference's eager torch ops compute on the GPU, operation by operation (every product
// and sum rounded on its own: separate torch kernels cannot contract into FMAs), so masks and copied values are
// bit-identical; only a child's position goes through the reference's cuBLAS bmm and is compared at 1e-6 relative.
#include "sfgs_common.cuh"

namespace {

constexpr int DN_THREADS = 256;
constexpr int DN_SPAN = 1024;        // Gaussians per block: the unit of the count scan
constexpr int DN_COUNTERS = 5;       // kept originals, kept clones, split sources, kept split sources, clone selections
constexpr int DN_MAX_FIELDS = 8;
constexpr int DN_MAX_COLS = 160;

enum : unsigned char { ACT_KEEP = 1, ACT_CLONE_KEEP = 2, ACT_SPLIT_SEL = 4, ACT_SPLIT_KEEP = 8, ACT_CLONE_SEL = 16 };

// ---------------------------------------------------------------------------------------------------------------------
// per-iteration statistics (train.py:314-315, gaussian_model.py:744-749)
__global__ void __launch_bounds__(256)
densification_stats_kernel(int P, const float* __restrict__ grad4, const int* __restrict__ radii,
                           float* __restrict__ max_radii2D, float* __restrict__ accum, float* __restrict__ accum_abs,
                           float* __restrict__ accum_abs_max, float* __restrict__ denom) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const int r = radii[i];
  if (r <= 0) return;                                       // visibility_filter = radii > 0
  const float4 g = reinterpret_cast<const float4*>(grad4)[i];
  // torch.norm over two elements: each square rounded, then the sum, then the root
  const float n_sgn = __fsqrt_rn(__fadd_rn(__fmul_rn(g.x, g.x), __fmul_rn(g.y, g.y)));
  const float n_abs = __fsqrt_rn(__fadd_rn(__fmul_rn(g.z, g.z), __fmul_rn(g.w, g.w)));
  max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
  accum[i] = __fadd_rn(accum[i], n_sgn);
  accum_abs[i] = __fadd_rn(accum_abs[i], n_abs);
  accum_abs_max[i] = fmaxf(accum_abs_max[i], n_abs);
  denom[i] = __fadd_rn(denom[i], 1.f);
}

// ---------------------------------------------------------------------------------------------------------------------
struct PlanArgs {
  int P;
  const float* accum;
  const float* accum_abs;
  const float* denom;
  const float* scaling;      // raw (log) scales [P,3]
  const float* opacity;      // raw (logit) opacity [P]
  float max_grad;
  const float* abs_threshold;   // device scalar: the quantile Q of gaussian_model.py:708
  float split_scale;         // percent_dense * extent
  float min_opacity;
  int screen_test;           // max_screen_size is truthy
  float max_screen;
  float world_scale;         // 0.1 * extent
};

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }
// scale of a split child: exp(log(exp(s) / (0.8 N))), N = 2; torch divides by a python scalar as a product with 1/1.6 = 0.625
__device__ __forceinline__ float child_raw_scale(float s_raw) { return logf(__fmul_rn(expf(s_raw), 0.625f)); }

__device__ __forceinline__ unsigned char plan_action(const PlanArgs& a, int i, float abs_thr) {
  const float d = a.denom[i];
  float g = __fdiv_rn(a.accum[i], d), ga = __fdiv_rn(a.accum_abs[i], d);
  if (g != g) g = 0.f;                                       // grads[grads.isnan()] = 0
  if (ga != ga) ga = 0.f;
  // torch.norm over the single column is sqrt(g*g) = |g| (exact in binary floating point away from under/overflow)
  const bool sel = fabsf(g) >= a.max_grad || fabsf(ga) >= abs_thr;
  const float s0 = a.scaling[3 * (size_t)i], s1 = a.scaling[3 * (size_t)i + 1], s2 = a.scaling[3 * (size_t)i + 2];
  const float smax = fmaxf(fmaxf(expf(s0), expf(s1)), expf(s2));
  const bool small = smax <= a.split_scale;
  const bool clone = sel && small, split = sel && !small;
  const bool faint = sigmoid_ref(a.opacity[i]) < a.min_opacity;
  // max_radii2D is all zeros by the time the reference prunes (densification_postfix resets it, gaussian_model.py:651)
  const bool screen_big = a.screen_test && 0.f > a.max_screen;
  const bool prune_self = faint || (a.screen_test && (screen_big || smax > a.world_scale));
  unsigned char act = 0;
  if (!split && !prune_self) act |= ACT_KEEP;
  if (clone) act |= ACT_CLONE_SEL;
  if (clone && !prune_self) act |= ACT_CLONE_KEEP;
  if (split) {
    act |= ACT_SPLIT_SEL;
    const float cmax = fmaxf(fmaxf(expf(child_raw_scale(s0)), expf(child_raw_scale(s1))), expf(child_raw_scale(s2)));
    const bool prune_child = faint || (a.screen_test && (screen_big || cmax > a.world_scale));
    if (!prune_child) act |= ACT_SPLIT_KEEP;
  }
  return act;
}

__device__ __forceinline__ void count_flags(unsigned char act, int c[DN_COUNTERS]) {
  c[0] = (act & ACT_KEEP) ? 1 : 0;
  c[1] = (act & ACT_CLONE_KEEP) ? 1 : 0;
  c[2] = (act & ACT_SPLIT_SEL) ? 1 : 0;
  c[3] = (act & ACT_SPLIT_KEEP) ? 1 : 0;
  c[4] = (act & ACT_CLONE_SEL) ? 1 : 0;
}

__global__ void __launch_bounds__(DN_THREADS)
densify_plan_kernel(PlanArgs a, unsigned char* __restrict__ action, int* __restrict__ block_counts) {
  __shared__ int s_cnt[DN_COUNTERS];
  if (threadIdx.x < DN_COUNTERS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const float abs_thr = *a.abs_threshold;
  int c[DN_COUNTERS] = {0, 0, 0, 0, 0};
  for (int k = 0; k < DN_SPAN / DN_THREADS; k++) {
    const int i = blockIdx.x * DN_SPAN + k * DN_THREADS + threadIdx.x;
    if (i < a.P) {
      const unsigned char act = plan_action(a, i, abs_thr);
      action[i] = act;
      int f[DN_COUNTERS];
      count_flags(act, f);
#pragma unroll
      for (int q = 0; q < DN_COUNTERS; q++) c[q] += f[q];
    }
  }
#pragma unroll
  for (int q = 0; q < DN_COUNTERS; q++) {
    int v = c[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[q], v);
  }
  __syncthreads();
  if (threadIdx.x < DN_COUNTERS) block_counts[blockIdx.x * DN_COUNTERS + threadIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of the per-block counts, in place; totals[q] = sum.  One block; chunks of 1024 blocks with a carry.
__global__ void __launch_bounds__(1024)
densify_scan_kernel(int nb, int* __restrict__ block_counts, int* __restrict__ totals) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int q = 0; q < DN_COUNTERS; q++) {
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
      const int b = base + t;
      const int v = b < nb ? block_counts[b * DN_COUNTERS + q] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (lane == 31) s_warp[w] = inc;
      __syncthreads();
      if (w == 0) {
        int x = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
        s_warp[lane] = x;               // inclusive over warps
      }
      __syncthreads();
      const int carry = s_carry;
      const int excl = carry + (w ? s_warp[w - 1] : 0) + inc - v;
      if (b < nb) block_counts[b * DN_COUNTERS + q] = excl;
      __syncthreads();
      if (t == 1023) s_carry = carry + s_warp[31];
      __syncthreads();
    }
    if (t == 0) totals[q] = s_carry;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct Field {
  const float* src[3];     // parameter, Adam exp_avg, Adam exp_avg_sq (the moments may be null: no optimizer state yet)
  float* dst[3];
  int width;
};
struct ApplyArgs {
  int P;
  int n_fields;
  int total_cols;
  int kept, kept_clones, split_sel, kept_split;
  const unsigned char* action;
  const int* block_offsets;
  const float* noise;      // [2 * split_sel, 3] standard normal
  Field f[DN_MAX_FIELDS];  // 0 xyz, 1 f_dc, 2 f_rest, 3 opacity, 4 scaling, 5 rotation, 6.. further per-Gaussian parameters
};

// one component of a split child's position: (build_rotation(rot) @ (eps * exp(scale)))[c] + xyz[c]
__device__ __forceinline__ float child_xyz(const ApplyArgs& a, int i, const float* eps, int c) {
  const float* q_ = a.f[5].src[0] + 4 * (size_t)i;
  const float* s_ = a.f[4].src[0] + 3 * (size_t)i;
  const float r0 = q_[0], r1 = q_[1], r2 = q_[2], r3 = q_[3];
  // utils/general_utils.py:78-99, one rounding per torch op
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)), __fmul_rn(r3, r3)));
  const float r = __fdiv_rn(r0, nrm), x = __fdiv_rn(r1, nrm), y = __fdiv_rn(r2, nrm), z = __fdiv_rn(r3, nrm);
  float R0, R1, R2;
  if (c == 0) {
    R0 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
    R1 = __fmul_rn(2.f, __fsub_rn(__fmul_rn(x, y), __fmul_rn(r, z)));
    R2 = __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, z), __fmul_rn(r, y)));
  } else if (c == 1) {
    R0 = __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, y), __fmul_rn(r, z)));
    R1 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z))));
    R2 = __fmul_rn(2.f, __fsub_rn(__fmul_rn(y, z), __fmul_rn(r, x)));
  } else {
    R0 = __fmul_rn(2.f, __fsub_rn(__fmul_rn(x, z), __fmul_rn(r, y)));
    R1 = __fmul_rn(2.f, __fadd_rn(__fmul_rn(y, z), __fmul_rn(r, x)));
    R2 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
  }
  const float v0 = __fmul_rn(eps[0], expf(s_[0])), v1 = __fmul_rn(eps[1], expf(s_[1])), v2 = __fmul_rn(eps[2], expf(s_[2]));
  const float dot = fmaf(R2, v2, fmaf(R1, v1, __fmul_rn(R0, v0)));        // the reference's bmm (cuBLAS): order not specified
  return __fadd_rn(dot, a.f[0].src[0][3 * (size_t)i + c]);
}

__global__ void __launch_bounds__(DN_THREADS)
densify_apply_kernel(ApplyArgs a) {
  __shared__ unsigned char s_act[DN_SPAN];
  __shared__ unsigned long long s_rank[DN_SPAN];    // five 12-bit exclusive in-block ranks per Gaussian
  __shared__ unsigned long long s_warp[DN_THREADS / 32];
  __shared__ unsigned char s_col_field[DN_MAX_COLS];
  __shared__ unsigned char s_col_off[DN_MAX_COLS];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int base = blockIdx.x * DN_SPAN;
  for (int j = t; j < DN_SPAN; j += DN_THREADS) s_act[j] = (base + j < a.P) ? a.action[base + j] : 0;
  if (t == 0) {
    int col = 0;
    for (int f = 0; f < a.n_fields; f++)
      for (int o = 0; o < a.f[f].width; o++, col++) { s_col_field[col] = (unsigned char)f; s_col_off[col] = (unsigned char)o; }
  }
  __syncthreads();
  // ---- in-block exclusive ranks: thread t owns Gaussians 4t .. 4t+3
  unsigned long long mine[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int f[DN_COUNTERS];
    count_flags(s_act[4 * t + k], f);
    unsigned long long p = 0;
#pragma unroll
    for (int q = 0; q < DN_COUNTERS; q++) p |= (unsigned long long)f[q] << (12 * q);
    mine[k] = sum;
    sum += p;
  }
  unsigned long long inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  unsigned long long before = inc - sum;
  for (int k = 0; k < w; k++) before += s_warp[k];
#pragma unroll
  for (int k = 0; k < 4; k++) s_rank[4 * t + k] = before + mine[k];
  __syncthreads();

  int off[DN_COUNTERS];
#pragma unroll
  for (int q = 0; q < DN_COUNTERS; q++) off[q] = a.block_offsets[blockIdx.x * DN_COUNTERS + q];
  const bool moments = a.f[0].src[1] != nullptr;

  // ---- one warp per Gaussian; lanes stride over the concatenated columns of all fields
  for (int j = w; j < DN_SPAN; j += DN_THREADS / 32) {
    const unsigned char act = s_act[j];
    if (!(act & (ACT_KEEP | ACT_CLONE_KEEP | ACT_SPLIT_KEEP))) continue;
    const int i = base + j;
    const unsigned long long rk = s_rank[j];
    const int r_keep = off[0] + (int)(rk & 0xfff), r_clone = off[1] + (int)((rk >> 12) & 0xfff);
    const int r_sel = off[2] + (int)((rk >> 24) & 0xfff), r_split = off[3] + (int)((rk >> 36) & 0xfff);
    for (int col = lane; col < a.total_cols; col += 32) {
      const int f = s_col_field[col], o = s_col_off[col];
      const Field& F = a.f[f];
      const int wd = F.width;
      const float v = F.src[0][(size_t)i * wd + o];
      if (act & ACT_KEEP) {
        const size_t d = (size_t)r_keep * wd + o;
        F.dst[0][d] = v;
        if (moments) { F.dst[1][d] = F.src[1][(size_t)i * wd + o]; F.dst[2][d] = F.src[2][(size_t)i * wd + o]; }
      }
      if (act & ACT_CLONE_KEEP) {
        const size_t d = (size_t)(a.kept + r_clone) * wd + o;
        F.dst[0][d] = v;
        if (moments) { F.dst[1][d] = 0.f; F.dst[2][d] = 0.f; }
      }
      if (act & ACT_SPLIT_KEEP) {
#pragma unroll
        for (int n = 0; n < 2; n++) {
          const size_t row = (size_t)a.kept + a.kept_clones + (size_t)n * a.kept_split + r_split;
          const size_t d = row * wd + o;
          float cv = v;
          if (f == 0) cv = child_xyz(a, i, a.noise + 3 * ((size_t)n * a.split_sel + r_sel), o);
          else if (f == 4) cv = child_raw_scale(v);
          F.dst[0][d] = cv;
          if (moments) { F.dst[1][d] = 0.f; F.dst[2][d] = 0.f; }
        }
      }
    }
  }
}

}  // namespace

extern "C" int sfgs_set_error(int code, const char* what, int cuda_error);   // sfgs_api.cu

static int cuda_fail(const char* what, cudaError_t e) { return sfgs_set_error(SFGS_E_CUDA, what, (int)e); }

extern "C" int sfgs_densification_stats(int P, const float* viewspace_grad, const int* radii, float* max_radii2D,
                                        float* grad_accum, float* grad_accum_abs, float* grad_accum_abs_max, float* denom,
                                        void* stream) {
  if (P < 0) return sfgs_set_error(SFGS_E_BADARG, "densification_stats: negative size", 0);
  if (P == 0) return SFGS_OK;
  if (!viewspace_grad || !radii || !max_radii2D || !grad_accum || !grad_accum_abs || !grad_accum_abs_max || !denom)
    return sfgs_set_error(SFGS_E_BADARG, "densification_stats: null pointer", 0);
  if (((uintptr_t)viewspace_grad & 15) != 0)
    return sfgs_set_error(SFGS_E_BADARG, "densification_stats: viewspace_grad must be 16-byte aligned [P,4]", 0);
  densification_stats_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, viewspace_grad, radii, max_radii2D, grad_accum,
                                                                              grad_accum_abs, grad_accum_abs_max, denom);
  g_sfgs_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SFGS_OK : cuda_fail("densification_stats: launch", e);
}

extern "C" int sfgs_densify_plan_blocks(int P) { return P <= 0 ? 0 : (P + DN_SPAN - 1) / DN_SPAN; }

extern "C" int sfgs_densify_plan(int P, const float* grad_accum, const float* grad_accum_abs, const float* denom,
                                 const float* scaling, const float* opacity, float max_grad, const float* abs_threshold,
                                 float split_scale, float min_opacity, int screen_test, float max_screen_size,
                                 float world_scale, unsigned char* action, int* block_offsets, int* totals_dev,
                                 int totals_host[5], void* stream) {
  if (P < 0) return sfgs_set_error(SFGS_E_BADARG, "densify_plan: negative size", 0);
  if (!totals_host) return sfgs_set_error(SFGS_E_BADARG, "densify_plan: null totals", 0);
  for (int q = 0; q < DN_COUNTERS; q++) totals_host[q] = 0;
  if (P == 0) return SFGS_OK;
  if (!grad_accum || !grad_accum_abs || !denom || !scaling || !opacity || !abs_threshold || !action || !block_offsets || !totals_dev)
    return sfgs_set_error(SFGS_E_BADARG, "densify_plan: null pointer", 0);
  cudaStream_t st = (cudaStream_t)stream;
  PlanArgs a{P, grad_accum, grad_accum_abs, denom, scaling, opacity, max_grad, abs_threshold,
             split_scale, min_opacity, screen_test, max_screen_size, world_scale};
  const int nb = (P + DN_SPAN - 1) / DN_SPAN;
  densify_plan_kernel<<<nb, DN_THREADS, 0, st>>>(a, action, block_offsets);
  densify_scan_kernel<<<1, 1024, 0, st>>>(nb, block_offsets, totals_dev);
  g_sfgs_launches.fetch_add(2, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("densify_plan: launch", e);
  // the caller sizes the new tensors from the totals: the one host synchronisation of densify_and_prune
  e = cudaMemcpyAsync(totals_host, totals_dev, DN_COUNTERS * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return cuda_fail("densify_plan: copy", e);
  e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? SFGS_OK : cuda_fail("densify_plan: synchronize", e);
}

extern "C" int sfgs_densify_apply(int P, const unsigned char* action, const int* block_offsets, const int totals[5],
                                  const float* noise, int n_fields, const int* widths, const float* const* src,
                                  const float* const* src_exp_avg, const float* const* src_exp_avg_sq, float* const* dst,
                                  float* const* dst_exp_avg, float* const* dst_exp_avg_sq, void* stream) {
  if (P < 0) return sfgs_set_error(SFGS_E_BADARG, "densify_apply: negative size", 0);
  if (P == 0) return SFGS_OK;
  if (!action || !block_offsets || !totals || !widths || !src || !dst)
    return sfgs_set_error(SFGS_E_BADARG, "densify_apply: null pointer", 0);
  if (n_fields < 6 || n_fields > DN_MAX_FIELDS)
    return sfgs_set_error(SFGS_E_BADARG, "densify_apply: fields are xyz, f_dc, f_rest, opacity, scaling, rotation and up to two more", 0);
  if (widths[0] != 3 || widths[3] != 1 || widths[4] != 3 || widths[5] != 4)
    return sfgs_set_error(SFGS_E_BADARG, "densify_apply: xyz/opacity/scaling/rotation must be 3/1/3/4 wide", 0);
  const bool moments = src_exp_avg != nullptr;
  if (moments && (!src_exp_avg_sq || !dst_exp_avg || !dst_exp_avg_sq))
    return sfgs_set_error(SFGS_E_BADARG, "densify_apply: Adam moments must be given for all or none", 0);
  ApplyArgs a{};
  a.P = P; a.n_fields = n_fields;
  a.kept = totals[0]; a.kept_clones = totals[1]; a.split_sel = totals[2]; a.kept_split = totals[3];
  a.action = action; a.block_offsets = block_offsets; a.noise = noise;
  if (a.split_sel > 0 && !noise) return sfgs_set_error(SFGS_E_BADARG, "densify_apply: noise is required when Gaussians split", 0);
  int cols = 0;
  for (int f = 0; f < n_fields; f++) {
    if (widths[f] <= 0 || !src[f] || !dst[f]) return sfgs_set_error(SFGS_E_BADARG, "densify_apply: bad field", 0);
    a.f[f].width = widths[f];
    a.f[f].src[0] = src[f]; a.f[f].dst[0] = dst[f];
    a.f[f].src[1] = moments ? src_exp_avg[f] : nullptr;  a.f[f].src[2] = moments ? src_exp_avg_sq[f] : nullptr;
    a.f[f].dst[1] = moments ? dst_exp_avg[f] : nullptr;  a.f[f].dst[2] = moments ? dst_exp_avg_sq[f] : nullptr;
    if (moments && (!a.f[f].src[1] || !a.f[f].src[2] || !a.f[f].dst[1] || !a.f[f].dst[2]))
      return sfgs_set_error(SFGS_E_BADARG, "densify_apply: null Adam moment", 0);
    cols += widths[f];
  }
  if (cols > DN_MAX_COLS) return sfgs_set_error(SFGS_E_BADARG, "densify_apply: more than 160 columns per Gaussian", 0);
  a.total_cols = cols;
  const int nb = (P + DN_SPAN - 1) / DN_SPAN;
  densify_apply_kernel<<<nb, DN_THREADS, 0, (cudaStream_t)stream>>>(a);
  g_sfgs_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SFGS_OK : cuda_fail("densify_apply: launch", e);
}
