// sfgs_render_bwd.cu — adjoint of the alpha compositing, one CTA per 16x16 tile.
//
// Replaces BACKWARD::render / renderCUDA (RAST/cuda_rasterizer/backward.cu:509-754).
// The reference issues >= 14 global float atomicAdd per (pixel, Gaussian) pair.
// Here every warp (an 8x4 pixel block) reduces its 14 partial gradients with a
// 16-shuffle butterfly, the 8 warps of a tile are combined in a fixed order
// through shared-memory slots, and one 16-byte vector reduction per (tile,
// Gaussian, quarter) reaches global memory: ~1000x fewer global atomics, and
// everything inside a tile is deterministic.
//
// The per-pixel recursion uses the scalar form of the reference's accumulators:
//   dL/dalpha_j = T_j * ( g_j - A_j ),  g_j = <dL/dpix, attr_j>,
//   A_j = alpha_{j+1} g_{j+1} + (1 - alpha_{j+1}) A_{j+1}
// which is the channel-sum of accum_rec/accum_red/accum_ren/accum_rea
// (backward.cu:660-711) and needs 3 registers instead of 16.
#include "sfgs_common.cuh"

namespace {

constexpr int BWD_THREADS = 256;
constexpr int BWD_WARPS = BWD_THREADS / 32;
constexpr int BWD_BATCH = 64;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// 16 values per lane -> component (lane >> 1) summed over the warp, valid in every lane
__device__ __forceinline__ float butterfly16(const float v[16], int lane) {
  float w[8], x[4], y[2], z;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float send = hi ? v[i] : v[i + 8];
      const float keep = hi ? v[i + 8] : v[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float send = hi ? w[i] : w[i + 4];
      const float keep = hi ? w[i + 4] : w[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const float send = hi ? x[i] : x[i + 2];
      const float keep = hi ? x[i + 2] : x[i];
      y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    const float send = hi ? y[0] : y[1];
    const float keep = hi ? y[1] : y[0];
    z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  z += __shfl_xor_sync(0xffffffffu, z, 1);
  return z;
}

template <bool HAS_EXTRA>
__global__ void __launch_bounds__(BWD_THREADS)
render_bwd_kernel(const uint2* __restrict__ ranges, const char* __restrict__ binning_base,
                  const uint32_t* __restrict__ hdr, int W, int H, int ED, int band0,
                  const float* __restrict__ bg_color, const float* __restrict__ rec,
                  const float* __restrict__ extras, const float* __restrict__ accum_alphas,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                  const float* __restrict__ dL_dpixel_depths, const float* __restrict__ dL_dpixel_norms,
                  const float* __restrict__ dL_dpixel_alphas, const float* __restrict__ dL_dpixel_extras,
                  float* __restrict__ acc /* [P,16] zero-initialised */, float* __restrict__ dL_dextras) {
  __shared__ __align__(16) float4 s_rec[2][BWD_BATCH][4];                 // 8 KB
  __shared__ uint32_t s_id[2][BWD_BATCH];
  __shared__ __align__(16) float s_part[BWD_WARPS][BWD_BATCH][16];        // 32 KB
  __shared__ unsigned long long s_mask[BWD_WARPS];
  __shared__ uint32_t s_maxc[BWD_WARPS];
  __shared__ uint32_t s_bits[2][BWD_WARPS][BWD_BATCH / 32];

  // the binning buffer layout depends on the capacity the forward used; it is recorded in the image header
  const unsigned long long cap = ((unsigned long long)hdr[HDR_CAP_HI] << 32) | hdr[HDR_CAP_LO];
  const BinningLayout bl(const_cast<char*>(binning_base), (size_t)cap);
  const uint32_t* __restrict__ point_list = bl.point_list;
  const unsigned char* __restrict__ inst_mask = bl.inst_mask;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tiles_x = (W + SFGS_TILE - 1) / SFGS_TILE;
  const int tile_y = blockIdx.y + band0;
  const int tile = tile_y * tiles_x + blockIdx.x;
  const int px = blockIdx.x * SFGS_TILE + (wid & 1) * 8 + (lane & 7);
  const int py = tile_y * SFGS_TILE + (wid >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float pixfx = (float)px, pixfy = (float)py;
  const size_t HW = (size_t)H * W;

  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  if (total == 0) return;

  const float T_final = inside ? (1 - accum_alphas[pix_id]) : 0;
  float T = T_final;
  const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;

  float dLc0 = 0, dLc1 = 0, dLc2 = 0, dLd = 0, dLn0 = 0, dLn1 = 0, dLn2 = 0, dLa = 0;
  if (inside) {
    dLc0 = dL_dpixels[0 * HW + pix_id]; dLc1 = dL_dpixels[1 * HW + pix_id]; dLc2 = dL_dpixels[2 * HW + pix_id];
    dLd = dL_dpixel_depths[pix_id];
    dLn0 = dL_dpixel_norms[0 * HW + pix_id]; dLn1 = dL_dpixel_norms[1 * HW + pix_id]; dLn2 = dL_dpixel_norms[2 * HW + pix_id];
    dLa = dL_dpixel_alphas[pix_id];
  }
  float bg_dot = 0;
  bg_dot += bg_color[0] * dLc0; bg_dot += bg_color[1] * dLc1; bg_dot += bg_color[2] * dLc2;

  // extra-attribute slow path state (generic fallback, Skyfall-GS never uses it)
  float accum_ree[HAS_EXTRA ? SFGS_MAX_EXTRA : 1];
  float last_extra[HAS_EXTRA ? SFGS_MAX_EXTRA : 1];
  float dL_dpixel_extra[HAS_EXTRA ? SFGS_MAX_EXTRA : 1];
  if (HAS_EXTRA) {
    for (int i = 0; i < SFGS_MAX_EXTRA; i++) { accum_ree[i] = 0; last_extra[i] = 0; dL_dpixel_extra[i] = 0; }
    if (inside) for (int i = 0; i < ED; i++) dL_dpixel_extra[i] = dL_dpixel_extras[i * HW + pix_id];
  }

  float A = 0.f, g_last = 0.f, last_alpha = 0.f;
  const float ddelx_dx = 0.5 * W;
  const float ddely_dy = 0.5 * H;

  // entries at list position >= max(last_contributor) over the tile are never used: skip them entirely
  uint32_t wmax = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) s_maxc[wid] = wmax;
  __syncthreads();
  uint32_t bmax = 0;
#pragma unroll
  for (int w = 0; w < BWD_WARPS; w++) bmax = max(bmax, s_maxc[w]);
  const int used = min((int)bmax, total);   // positions [0, used) matter
  if (used == 0) return;
  const int nbatches = (used + BWD_BATCH - 1) / BWD_BATCH;

  // batch b holds positions used-1-b*BATCH-e, e = 0..BATCH-1 (back to front); thread e stages record e
  auto issue = [&](int batch, int stage) {
    unsigned m = 0;
    if (tid < BWD_BATCH) {
      const int pos = used - 1 - batch * BWD_BATCH - tid;
      if (pos >= 0) {
        m = inst_mask[range.x + pos];
        if (m) {
          const uint32_t id = point_list[range.x + pos];
          const float* src = rec + (size_t)id * REC_FLOATS;
#pragma unroll
          for (int q = 0; q < 4; q++) cp_async16(&s_rec[stage][tid][q], src + q * 4);
          s_id[stage][tid] = id;
        }
      }
    }
    cp_async_commit();
    if (tid < BWD_BATCH) {
#pragma unroll
      for (int blk = 0; blk < BWD_WARPS; blk++) {
        const unsigned word = __ballot_sync(0xffffffffu, (m >> blk) & 1u);
        if (lane == 0) s_bits[stage][blk][wid] = word;
      }
    }
  };

  issue(0, 0);
  for (int b = 0; b < nbatches; b++) {
    const int stage = b & 1;
    cp_async_wait_all();
    __syncthreads();   // stage visible; previous flush finished reading s_part / s_mask
    if (b + 1 < nbatches) issue(b + 1, stage ^ 1);
    const int first_pos = used - 1 - b * BWD_BATCH;
    const int cnt = min(BWD_BATCH, first_pos + 1);
    unsigned long long mymask = 0ull;

    // entries this warp can use: reach bit set and pos < wmax  <=>  e >= first_pos - wmax + 1
    int e0 = first_pos - (int)wmax + 1;
    if (e0 < 0) e0 = 0;
    for (int word = e0 >> 5; word < BWD_BATCH / 32; word++) {
     unsigned bits = s_bits[stage][wid][word];
     if (word == (e0 >> 5)) bits &= 0xffffffffu << (e0 & 31);
     while (bits) {
      const int e = word * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      const int pos = first_pos - e;
      const float4 ra = s_rec[stage][e][0];   // mx, my, con.x, con.y
      const float4 rb = s_rec[stage][e][1];   // con.z, opac, depth
      const float dx = ra.x - pixfx, dy = ra.y - pixfy;
      const float power = -0.5f * (ra.z * dx * dx + rb.x * dy * dy) - ra.w * dx * dy;
      const float G = exp(power);
      const float alpha = min(0.99f, rb.y * G);
      const bool active = ((uint32_t)pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
      if (!__any_sync(0xffffffffu, active)) continue;

      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = 0.f;
      if (active) {
        const float4 rc = s_rec[stage][e][2];   // r, g, b, nx
        const float4 rd = s_rec[stage][e][3];   // ny, nz
        const float inv_1ma = __frcp_rn(1.f - alpha);   // shared by the T recovery and the background term
        T = T * inv_1ma;
        const float weight = alpha * T;
        float g = rc.x * dLc0;
        g += rc.y * dLc1; g += rc.z * dLc2; g += rb.z * dLd;
        g += rc.w * dLn0; g += rd.x * dLn1; g += rd.y * dLn2; g += dLa;
        A = last_alpha * g_last + (1.f - last_alpha) * A;
        g_last = g;
        float dL_dalpha = g - A;
        if (HAS_EXTRA) {
          const uint32_t gid = s_id[stage][e];
          for (int ch = 0; ch < ED; ch++) {
            const float ex = extras[(size_t)gid * ED + ch];
            accum_ree[ch] = last_alpha * last_extra[ch] + (1.f - last_alpha) * accum_ree[ch];
            last_extra[ch] = ex;
            dL_dalpha += (ex - accum_ree[ch]) * dL_dpixel_extra[ch];
            atomicAdd(&dL_dextras[(size_t)gid * ED + ch], weight * dL_dpixel_extra[ch]);
          }
        }
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final * inv_1ma) * bg_dot;

        const float dL_dG = rb.y * dL_dalpha;
        const float gdx = G * dx, gdy = G * dy;
        const float dG_ddelx = -gdx * ra.z - gdy * ra.w;
        const float dG_ddely = -gdy * rb.x - gdx * ra.w;
        v[0] = weight * dLc0; v[1] = weight * dLc1; v[2] = weight * dLc2;
        v[3] = weight * dLd;
        v[4] = weight * dLn0; v[5] = weight * dLn1; v[6] = weight * dLn2;
        v[7] = dL_dG * dG_ddelx * ddelx_dx;
        v[8] = dL_dG * dG_ddely * ddely_dy;
        v[9] = fabsf(v[7]) + fabsf(v[8]);
        v[10] = -0.5f * gdx * dx * dL_dG;
        v[11] = -0.5f * gdx * dy * dL_dG;
        v[12] = -0.5f * gdy * dy * dL_dG;
        v[13] = G * dL_dalpha;
      }
      const float z = butterfly16(v, lane);
      if ((lane & 1) == 0) s_part[wid][e][lane >> 1] = z;
      mymask |= (1ull << e);
     }
    }
    if (lane == 0) s_mask[wid] = mymask;
    __syncthreads();
    // flush: thread -> (entry, quarter); fixed warp order => deterministic per tile
    {
      const int e = tid >> 2, q = tid & 3;
      if (e < cnt) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        bool any = false;
#pragma unroll
        for (int w = 0; w < BWD_WARPS; w++) {
          if ((s_mask[w] >> e) & 1ull) {
            const float4 p = *reinterpret_cast<const float4*>(&s_part[w][e][q * 4]);
            s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            any = true;
          }
        }
        if (any) red_add_v4(acc + (size_t)s_id[stage][e] * 16 + q * 4, s);
      }
    }
  }
}

}  // namespace

void sfgs_launch_render_bwd(const sfgs_backward_args* a, const GeomLayout& g, const ImageLayout& im,
                            const BinningLayout& b, float* acc, cudaStream_t st) {
  const int band0 = a->tile_row_end > a->tile_row_begin ? a->tile_row_begin : 0;
  const int band1 = a->tile_row_end > a->tile_row_begin ? (a->tile_row_end < im.tiles_y ? a->tile_row_end : im.tiles_y) : im.tiles_y;
  if (band1 <= band0) return;
  dim3 grid(im.tiles_x, band1 - band0, 1);
  SFGS_COUNT_LAUNCH();
  if (a->ED > 0)
    render_bwd_kernel<true><<<grid, BWD_THREADS, 0, st>>>(
        im.ranges, (const char*)b.point_list, im.hdr, a->width, a->height, a->ED, band0, a->background, g.rec, a->extra_attrs,
        a->accum_alphas,
        im.n_contrib, a->dL_dpix, a->dL_dpix_depth, a->dL_dpix_norm, a->dL_dpix_alpha, a->dL_dpix_extra, acc,
        a->dL_dextra);
  else
    render_bwd_kernel<false><<<grid, BWD_THREADS, 0, st>>>(
        im.ranges, (const char*)b.point_list, im.hdr, a->width, a->height, 0, band0, a->background, g.rec, nullptr,
        a->accum_alphas,
        im.n_contrib, a->dL_dpix, a->dL_dpix_depth, a->dL_dpix_norm, a->dL_dpix_alpha, nullptr, acc, nullptr);
}
