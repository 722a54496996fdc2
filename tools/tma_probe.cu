// tools/tma_probe.cu — which TMA box shapes does cp.async.bulk.tensor accept on this GPU?  (debug aid)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
using u64 = unsigned long long;
struct Map1 { CUtensorMap m; };
template <int RANK>
__global__ void probe(const __grid_constant__ Map1 tm, int bw, int bh, int x, int y, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* box = reinterpret_cast<float*>(sm);
  u64* bar = reinterpret_cast<u64*>(sm + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bw * bh * 4));
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"((unsigned)__cvta_generic_to_shared(box)), "l"(&tm.m), "r"(x), "r"(y), "r"(0), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"((unsigned)__cvta_generic_to_shared(box)), "l"(&tm.m), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
  }
  __syncthreads();
  unsigned ok = 0; const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(a) : "memory");
  if (threadIdx.x == 0) { out[0] = box[0]; out[1] = box[bw * 5 + 5]; out[2] = box[bw * bh - 1]; }
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  const int W = 1920, H = 1080, PL = 3;
  std::vector<float> h((size_t)W * H * PL);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 1000) + 1.f;
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 16); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792);
  struct Cfg { int rank, bw, bh, x, y; } cfgs[] = {{2, 64, 32, 0, -5}, {2, 64, 32, -8, -5}, {2, 80, 42, -8, -5}, {3, 80, 42, -8, -5},
                                                   {3, 80, 42, 1912, 1070}, {3, 80, 42, -4, -5}, {3, 76, 42, -4, -5}, {3, 80, 42, -5, -5}};
  for (auto c : cfgs) {
    Map1 tm = {};
    cuuint64_t dims[3] = {W, H, PL}; cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank %d box %dx%d at (%d,%d): encode failed %d\n", c.rank, c.bw, c.bh, c.x, c.y, (int)r); continue; }
    if (c.rank == 3) probe<3><<<1, 128, 33792>>>(tm, c.bw, c.bh, c.x, c.y, o); else probe<2><<<1, 128, 33792>>>(tm, c.bw, c.bh, c.x, c.y, o);
    cudaError_t e = cudaDeviceSynchronize();
    float ho[3] = {0, 0, 0};
    if (e == cudaSuccess) cudaMemcpy(ho, o, 12, cudaMemcpyDeviceToHost);
    printf("rank %d box %dx%d at (%d,%d): %s  box[0]=%g box[5][5]=%g last=%g (expect h[0]=%g)\n", c.rank, c.bw, c.bh, c.x, c.y,
           cudaGetErrorString(e), ho[0], ho[1], ho[2], h[0]);
    if (e != cudaSuccess) { printf("context lost, stopping\n"); return 1; }
  }
  return 0;
}
