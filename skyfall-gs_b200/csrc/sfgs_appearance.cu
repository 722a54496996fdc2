// sfgs_appearance.cu — the appearance path of render() as one tensor-core kernel.
//
// Replaces, for the forward pass, gaussian_renderer/__init__.py:105-118 + EmbeddingModel.forward
// (scene/gaussian_model.py:45-69) + eval_sh (utils/sh_utils.py) — in torch: a [P,59] concat, three Linear layers with
// [P,128] fp32 intermediates, a repeat, a [P,48] tone map, a [P,3,16] transpose, the SH evaluation and four clamps,
// ~25 kernels and ~2.5 KB of HBM traffic per Gaussian — by one kernel that reads each Gaussian's 48 SH features, 24
// Fourier features and position once and writes its 3 colours (324 B per Gaussian):
//
//   inp    = [min(features[:3], 1) | gemb | aemb]                      (59 = 3 + 24 + 32)
//   h      = W3 relu(W2 relu(W1 inp + b1) + b2) + b3 ;  offset, mul = 0.01 h[:3], 0.01 h[3:]
//   toned  = min(min(features, 1) * mul[ch] + [offset/C0 on the DC term], 1)           ([16,3] per Gaussian)
//   colour = max(eval_sh(D, toned, normalize(xyz - campos)) + 0.5, 0)
//
// The per-camera embedding `aemb` is the same vector for every Gaussian, so its 32 columns of W1 fold into the bias
// (b1' = b1 + W1[:, 27:] aemb, computed once per CTA) and layer 1 contracts over K = 27 (padded to 32).
//
// Tensor cores: a CTA owns a tile of 128 Gaussians = the M = 128 rows of three tcgen05.mma (kind::f16, split-bf16
// operands, fp32 accumulate) GEMMs: [128x32]x[32x128], [128x128]x[128x128], [128x128]x[128x16].  Weights are converted to bf16
// once per CTA and stay in shared memory in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 bytes per
// core; SBO = 128 B between 8-row groups, LBO = bytes between the two 16-byte K chunks of one MMA); activations never
// leave the SM: thread m packs row m of the A operand into the same layout, one elected thread issues the MMAs, the
// fp32 accumulator lives in TMEM (128 lanes x 128 columns) and comes back with tcgen05.ld (lane m -> thread m) for the
// bias + ReLU + bf16 pack of the next layer.  CTAs are persistent over tiles and run TWO tiles at a time: two groups
// of 128 threads, each with its own operand buffer, accumulator columns and mbarrier, synchronised only inside the
// group (named barriers), so one group's epilogue arithmetic overlaps the other group's MMAs and loads.
//
// Precision: the multiplier head of a trained model is O(1) (h ~ 100 before the x0.01), so plain bf16 operands
// (2^-9 relative) would leave ~1e-2 in the colours (measured on the golden vectors: 8e-3).  Every operand is therefore
// carried as TWO bf16 values, x = hi + lo with lo = bf16(x - hi) (16 mantissa bits together), and every product is
// three tensor-core MMAs into the same fp32 accumulator:  A W ~= Ah Wh + Ah Wl + Al Wh  (the dropped Al Wl term is
// 2^-18 relative).  The GEMMs stay on tcgen05 at 3x the MMA count — still far below the kernel's HBM time — and the
// colours agree with the reference's float32 modules to 1e-4 (tests/test_appearance.py).  Everything after the MLP is
// float32.
#include <cuda_bf16.h>
#include "sfgs_common.cuh"

namespace {

constexpr int AP_THREADS = 256;       // two groups of 128 threads, each with its own tile in flight
constexpr int AP_GROUPS = 2;
constexpr int AP_M = 128;            // Gaussians per tile
constexpr int AP_H = 128;            // hidden width
constexpr int AP_K1 = 32;            // 3 + 24 padded
constexpr int AP_N3 = 16;            // 6 outputs padded to the smallest N of an M = 128 MMA
constexpr int AP_G = 24, AP_E = 32;  // Fourier features per Gaussian, per-camera embedding size
constexpr int AP_TMEM_COLS = 256;      // 128 accumulator columns per group

struct alignas(128) ApSmem {                       // [2] = {hi, lo} halves of the split operands
  __nv_bfloat16 A[AP_GROUPS][2][AP_H / 8][AP_M / 8][8][8];   // 2 x 64 KB  activations, K-major cores: [k chunk][row group][row][8 k]
  __nv_bfloat16 W1[2][AP_K1 / 8][AP_H / 8][8][8];   // 16 KB  [k chunk][n group][n][8 k]
  __nv_bfloat16 W2[2][AP_H / 8][AP_H / 8][8][8];    // 64 KB
  __nv_bfloat16 W3[2][AP_H / 8][AP_N3 / 8][8][8];   //  8 KB
  float b1[AP_H], b2[AP_H], b3[8];
  unsigned long long bar[AP_GROUPS];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t smem_desc(const void* p, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory matrix descriptor, no swizzle ("interleave"): start >> 4 in [0,14), LBO >> 4 in [16,30),
  // SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64)
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  const uint32_t zero = 0;   // disable-output-lane mask: none
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(zero) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok;
  uint32_t spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();      // a lost completion becomes a launch failure, never a hung GPU
  } while (!ok);
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                    *reinterpret_cast<uint32_t*>(&d));
}
// x = hi + lo: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = v[i] - __bfloat162float(__float2bfloat16(v[i]));
  hi = pack8(v);
  lo = pack8(r);
}
__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* lo, float x) {
  const __nv_bfloat16 h = __float2bfloat16(x);
  *hi = h;
  *lo = __float2bfloat16(x - __bfloat162float(h));
}

__global__ void __launch_bounds__(AP_THREADS, 1)
appearance_fwd_kernel(int P, int D, const float* __restrict__ features /*[P,16,3]*/, const float* __restrict__ gemb /*[P,24]*/,
                      const float* __restrict__ aemb /*[32]*/, const float* __restrict__ W1 /*[128,59]*/,
                      const float* __restrict__ b1, const float* __restrict__ W2 /*[128,128]*/, const float* __restrict__ b2,
                      const float* __restrict__ W3 /*[6,128]*/, const float* __restrict__ b3,
                      const float* __restrict__ means3D, const float* __restrict__ campos, float* __restrict__ colors) {
  extern __shared__ __align__(128) unsigned char ap_smem[];
  ApSmem& S = *reinterpret_cast<ApSmem*>(ap_smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid >> 7, t = tid & 127;       // group and row inside the group's tile
  constexpr int IN = 3 + AP_G + AP_E;     // 59 columns of W1
  auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };

  // ---- once per CTA: TMEM, barrier, weights -> bf16 core-matrix layout, folded bias
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(&S.tmem_base)), "r"(AP_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&S.bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&S.bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < AP_H * AP_K1; i += AP_THREADS) {          // W1[n][k], k < 27 (colour 3 + Fourier 24), zero padded
    const int n = i / AP_K1, k = i - n * AP_K1;
    split_store(&S.W1[0][k >> 3][n >> 3][n & 7][k & 7], &S.W1[1][k >> 3][n >> 3][n & 7][k & 7], k < 3 + AP_G ? W1[n * IN + k] : 0.f);
  }
  for (int i = tid; i < AP_H * AP_H; i += AP_THREADS) {
    const int n = i / AP_H, k = i - n * AP_H;
    split_store(&S.W2[0][k >> 3][n >> 3][n & 7][k & 7], &S.W2[1][k >> 3][n >> 3][n & 7][k & 7], W2[i]);
  }
  for (int i = tid; i < AP_N3 * AP_H; i += AP_THREADS) {
    const int n = i / AP_H, k = i - n * AP_H;
    split_store(&S.W3[0][k >> 3][n >> 3][n & 7][k & 7], &S.W3[1][k >> 3][n >> 3][n & 7][k & 7], n < 6 ? W3[n * AP_H + k] : 0.f);
  }
  if (tid < AP_H) {
    float acc = b1[tid];
#pragma unroll 8
    for (int e = 0; e < AP_E; e++) acc = fmaf(W1[tid * IN + 3 + AP_G + e], aemb[e], acc);
    S.b1[tid] = acc;
    S.b2[tid] = b2[tid];
    if (tid < 8) S.b3[tid] = tid < 6 ? b3[tid] : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_all = S.tmem_base;
  const uint32_t tmem = tmem_all + (uint32_t)(grp * 128);                   // this group's 128 accumulator columns
  const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);    // a warp reaches the lane quarter warp % 4
  auto& A = S.A[grp];
  unsigned long long* bar = &S.bar[grp];
  uint32_t phase = 0;
  const float cx = campos[0], cy = campos[1], cz = campos[2];

  const int ntiles = (P + AP_M - 1) / AP_M;
  // One thread per group asks L2 for the NEXT tile's rows (three contiguous spans) while this tile is computed: the
  // row-per-thread loads below then find their lines in L2 instead of paying the HBM latency in the critical path.
  auto prefetch_tile = [&](int tl) {
    if (tl >= ntiles) return;
    const int r0 = tl * AP_M;
    const unsigned rows = (unsigned)((P - r0) < AP_M ? (P - r0) : AP_M);
    const float* pf = features + (size_t)r0 * 48;
    const float* pg = gemb + (size_t)r0 * AP_G;
    const float* pm = means3D + (size_t)r0 * 3;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf), "r"(rows * 192u) : "memory");
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pg), "r"(rows * (unsigned)(AP_G * 4)) : "memory");
    const unsigned mb = (rows * 12u) & ~15u;
    if (mb && (reinterpret_cast<uintptr_t>(pm) & 15) == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pm), "r"(mb) : "memory");
  };
  for (int tile = blockIdx.x * AP_GROUPS + grp; tile < ntiles; tile += gridDim.x * AP_GROUPS) {
    if (t == 0) prefetch_tile(tile + gridDim.x * AP_GROUPS);
    const int g = tile * AP_M + t;
    const bool valid = g < P;
    const int gi = valid ? g : P - 1;                                   // rows past P shadow the last Gaussian
    // ---- A0: row m = [min(dc, 1) (3) | Fourier features (24) | 0 (5)]
    {
      float in[AP_K1];
      const float* f = features + (size_t)gi * 48;
      in[0] = fminf(f[0], 1.f); in[1] = fminf(f[1], 1.f); in[2] = fminf(f[2], 1.f);
      const float4* g4 = reinterpret_cast<const float4*>(gemb + (size_t)gi * AP_G);
#pragma unroll
      for (int q = 0; q < AP_G / 4; q++) {
        const float4 v = g4[q];
        in[3 + 4 * q] = v.x; in[4 + 4 * q] = v.y; in[5 + 4 * q] = v.z; in[6 + 4 * q] = v.w;
      }
#pragma unroll
      for (int k = 3 + AP_G; k < AP_K1; k++) in[k] = 0.f;
#pragma unroll
      for (int kc = 0; kc < AP_K1 / 8; kc++) {
        uint4 hi, lo;
        split8(in + 8 * kc, hi, lo);
        *reinterpret_cast<uint4*>(&A[0][kc][t >> 3][t & 7][0]) = hi;
        *reinterpret_cast<uint4*>(&A[1][kc][t >> 3][t & 7][0]) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync();
    // ---- layer 1: [128 x 32] x [32 x 128]
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < AP_K1 / 16; ks++) {
        const uint64_t ah = smem_desc(&A[0][2 * ks][0][0][0], 2048, 128), al = smem_desc(&A[1][2 * ks][0][0][0], 2048, 128);
        const uint64_t wh = smem_desc(&S.W1[0][2 * ks][0][0][0], 2048, 128), wl = smem_desc(&S.W1[1][2 * ks][0][0][0], 2048, 128);
        umma_bf16(tmem, ah, wh, instr_desc(AP_M, AP_H), ks > 0);
        umma_bf16(tmem, ah, wl, instr_desc(AP_M, AP_H), 1);
        umma_bf16(tmem, al, wh, instr_desc(AP_M, AP_H), 1);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- h1 = relu(acc + b1') -> A (K = 128)
#pragma unroll
    for (int c0 = 0; c0 < AP_H; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_lane + c0, v);
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i] + S.b1[c0 + i], 0.f);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint4 hi, lo;
        split8(v + 8 * q, hi, lo);
        *reinterpret_cast<uint4*>(&A[0][(c0 >> 3) + q][t >> 3][t & 7][0]) = hi;
        *reinterpret_cast<uint4*>(&A[1][(c0 >> 3) + q][t >> 3][t & 7][0]) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync();
    // ---- layer 2: [128 x 128] x [128 x 128]
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < AP_H / 16; ks++) {
        const uint64_t ah = smem_desc(&A[0][2 * ks][0][0][0], 2048, 128), al = smem_desc(&A[1][2 * ks][0][0][0], 2048, 128);
        const uint64_t wh = smem_desc(&S.W2[0][2 * ks][0][0][0], 2048, 128), wl = smem_desc(&S.W2[1][2 * ks][0][0][0], 2048, 128);
        umma_bf16(tmem, ah, wh, instr_desc(AP_M, AP_H), ks > 0);
        umma_bf16(tmem, ah, wl, instr_desc(AP_M, AP_H), 1);
        umma_bf16(tmem, al, wh, instr_desc(AP_M, AP_H), 1);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < AP_H; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_lane + c0, v);
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i] + S.b2[c0 + i], 0.f);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint4 hi, lo;
        split8(v + 8 * q, hi, lo);
        *reinterpret_cast<uint4*>(&A[0][(c0 >> 3) + q][t >> 3][t & 7][0]) = hi;
        *reinterpret_cast<uint4*>(&A[1][(c0 >> 3) + q][t >> 3][t & 7][0]) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync();
    // ---- layer 3: [128 x 128] x [128 x 16]
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < AP_H / 16; ks++) {
        const uint64_t ah = smem_desc(&A[0][2 * ks][0][0][0], 2048, 128), al = smem_desc(&A[1][2 * ks][0][0][0], 2048, 128);
        const uint64_t wh = smem_desc(&S.W3[0][2 * ks][0][0][0], 256, 128), wl = smem_desc(&S.W3[1][2 * ks][0][0][0], 256, 128);
        umma_bf16(tmem, ah, wh, instr_desc(AP_M, AP_N3), ks > 0);
        umma_bf16(tmem, ah, wl, instr_desc(AP_M, AP_N3), 1);
        umma_bf16(tmem, al, wh, instr_desc(AP_M, AP_N3), 1);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float h[8];
    tmem_ld8(tmem_lane, h);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the next tile's MMAs overwrite these columns

    // ---- tone map + SH evaluation (fp32), scene/gaussian_model.py:60-69, gaussian_renderer/__init__.py:109-117
    float off[3], mul[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { off[c] = (h[c] + S.b3[c]) * 0.01f / SH_C0; mul[c] = (h[3 + c] + S.b3[3 + c]) * 0.01f; }
    const float px = means3D[3 * (size_t)gi], py = means3D[3 * (size_t)gi + 1], pz = means3D[3 * (size_t)gi + 2];
    float dx = px - cx, dy = py - cy, dz = pz - cz;
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx * inv, y = dy * inv, z = dz * inv;
    float bas[16];
    bas[0] = SH_C0;
    bas[1] = -SH_C1 * y; bas[2] = SH_C1 * z; bas[3] = -SH_C1 * x;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    bas[4] = SH_C2_0 * xy; bas[5] = SH_C2_1 * yz; bas[6] = SH_C2_2 * (2.f * zz - xx - yy); bas[7] = SH_C2_3 * xz;
    bas[8] = SH_C2_4 * (xx - yy);
    bas[9] = SH_C3_0 * y * (3.f * xx - yy); bas[10] = SH_C3_1 * xy * z; bas[11] = SH_C3_2 * y * (4.f * zz - xx - yy);
    bas[12] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy); bas[13] = SH_C3_4 * x * (4.f * zz - xx - yy);
    bas[14] = SH_C3_5 * z * (xx - yy); bas[15] = SH_C3_6 * x * (xx - 3.f * yy);
    const int ncoef = (D + 1) * (D + 1);
    float rgb[3] = {0.f, 0.f, 0.f};
    const float4* f4 = reinterpret_cast<const float4*>(features + (size_t)gi * 48);
#pragma unroll
    for (int q = 0; q < 12; q++) {
      const float4 v = f4[q];
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = 4 * q + u, k = j / 3, ch = j - 3 * k;          // features[k][ch]
        float tv = fminf(e[u], 1.f) * mul[ch] + (k == 0 ? off[ch] : 0.f);
        tv = fminf(tv, 1.f);
        if (k < ncoef) rgb[ch] = fmaf(bas[k], tv, rgb[ch]);
      }
    }
    if (valid) {
      colors[3 * (size_t)g] = fmaxf(rgb[0] + 0.5f, 0.f);
      colors[3 * (size_t)g + 1] = fmaxf(rgb[1] + 0.5f, 0.f);
      colors[3 * (size_t)g + 2] = fmaxf(rgb[2] + 0.5f, 0.f);
    }
    group_sync();      // every lane has read its accumulator columns and A before the next tile rewrites them
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "r"(AP_TMEM_COLS) : "memory");
}

}  // namespace

extern "C" int sfgs_set_error(int code, const char* what, int cuda_error);   // sfgs_api.cu

extern "C" int sfgs_appearance_forward(int P, int D, int M, const float* features, const float* gemb, int G, const float* aemb,
                                       int E, const float* W1, const float* b1, const float* W2, const float* b2,
                                       const float* W3, const float* b3, const float* means3D, const float* campos,
                                       float* colors, void* stream) {
  if (P < 0) return sfgs_set_error(SFGS_E_BADARG, "appearance_forward: P < 0", 0);
  if (P == 0) return SFGS_OK;
  if (M != 16 || G != AP_G || E != AP_E || D < 0 || D > 3)
    return sfgs_set_error(SFGS_E_UNSUPPORTED, "appearance_forward: built for 16 SH coefficients, 24 Fourier features, a 32-wide embedding", 0);
  if (!features || !gemb || !aemb || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !means3D || !campos || !colors)
    return sfgs_set_error(SFGS_E_BADARG, "appearance_forward: null pointer", 0);
  if ((reinterpret_cast<uintptr_t>(features) & 15) || (reinterpret_cast<uintptr_t>(gemb) & 15))
    return sfgs_set_error(SFGS_E_BADARG, "appearance_forward: features / gemb must be 16-byte aligned", 0);
  static SfgsPerDeviceOnce once;
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (once.first_use() || (dev >= 0 && dev < 64 && sms[dev] == 0)) {
    cudaFuncSetAttribute(appearance_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ApSmem));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) sms[dev] = n;
  }
  const int nsm = (dev >= 0 && dev < 64 && sms[dev] > 0) ? sms[dev] : 148;
  const int ntiles = (P + AP_M - 1) / AP_M;
  const int pairs = (ntiles + AP_GROUPS - 1) / AP_GROUPS;
  const int grid = pairs < nsm ? pairs : nsm;                // persistent: one CTA (two tile groups) per SM
  SFGS_COUNT_LAUNCH();
  appearance_fwd_kernel<<<grid, AP_THREADS, sizeof(ApSmem), (cudaStream_t)stream>>>(P, D, features, gemb, aemb, W1, b1, W2, b2,
                                                                                  W3, b3, means3D, campos, colors);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SFGS_OK : sfgs_set_error(SFGS_E_CUDA, "appearance_forward", (int)e);
}
