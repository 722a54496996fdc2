// oracle/ref_shim.cu — TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" face over the UNMODIFIED reference rasterizer
// (CudaRasterizer::Rasterizer, RAST/cuda_rasterizer/rasterizer.h:20-103) and
// SimpleKNN::knn (KNN/simple_knn.h), compiled together with the reference's own
// .cu files *where they lie* under /root/reference by oracle/build_ref.py into
// oracle/_ref/libref_rasterizer.so.  No reference source is copied into this
// repository; this file only calls the reference's public C++ API and uses its
// own state structs (rasterizer_impl.h) to expose intermediate buffers to the
// parity tests.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
#include <cstdint>
#include <functional>
#include <cuda_runtime.h>
#include "rasterizer.h"
#include "rasterizer_impl.h"
#include "simple_knn.h"

typedef char* (*ref_alloc_fn)(void* user, size_t bytes);

extern "C" {

int ref_forward(ref_alloc_fn geom, void* geom_user, ref_alloc_fn binning, void* binning_user, ref_alloc_fn img,
                void* img_user, int P, int D, int M, int ED, const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                const float* norm3D_precomp, const float* extra_attrs, const float* viewmatrix,
                const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, float kernel_size,
                int prefiltered, float* out_color, float* out_depth, float* out_norm, float* out_alpha,
                float* out_extra, int* radii, int debug) {
  std::function<char*(size_t)> g = [=](size_t n) { return geom(geom_user, n); };
  std::function<char*(size_t)> b = [=](size_t n) { return binning(binning_user, n); };
  std::function<char*(size_t)> i = [=](size_t n) { return img(img_user, n); };
  try {
    return CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, ED, background, W, H, means3D, shs, colors_precomp,
                                               opacities, scales, scale_modifier, rotations, cov3D_precomp,
                                               norm3D_precomp, extra_attrs, viewmatrix, projmatrix, cam_pos, tan_fovx,
                                               tan_fovy, kernel_size, prefiltered != 0, out_color, out_depth, out_norm,
                                               out_alpha, out_extra, radii, debug != 0);
  } catch (...) {
    return -1;
  }
}

int ref_backward(int P, int D, int M, int R, int ED, const float* background, int W, int H, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* norm3D_precomp,
                 const float* extra_attrs, const float* viewmatrix, const float* projmatrix, const float* campos,
                 float tan_fovx, float tan_fovy, float kernel_size, const int* radii, char* geom_buffer,
                 char* binning_buffer, char* img_buffer, const float* accum_alphas, const float* dL_dpix,
                 const float* dL_dpix_depth, const float* dL_dpix_norm, const float* dL_dpix_alpha,
                 const float* dL_dpix_extra, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dnorm3D,
                 float* dL_dsh, float* dL_dscale, float* dL_drot, float* dL_dextra, int debug) {
  try {
    CudaRasterizer::Rasterizer::backward(P, D, M, R, ED, background, W, H, means3D, shs, colors_precomp, scales,
                                         scale_modifier, rotations, cov3D_precomp, norm3D_precomp, extra_attrs,
                                         viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, kernel_size, radii,
                                         geom_buffer, binning_buffer, img_buffer, accum_alphas, dL_dpix, dL_dpix_depth,
                                         dL_dpix_norm, dL_dpix_alpha, dL_dpix_extra, dL_dmean2D, dL_dconic,
                                         dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dnorm3D, dL_dsh,
                                         dL_dscale, dL_drot, dL_dextra, debug != 0);
    return 0;
  } catch (...) {
    return -1;
  }
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, unsigned char* present) {
  CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
}

// Sub-array pointers of the reference's opaque state buffers (same carving as the reference itself).
struct ref_geom_view {
  float* depths; unsigned char* clamped; float* means2D; float* cov3D; float* conic_opacity; float* rgb;
  float* norm3D; uint32_t* tiles_touched; uint32_t* point_offsets;
};
struct ref_binning_view { uint64_t* keys_unsorted; uint64_t* keys; uint32_t* list_unsorted; uint32_t* list; };
struct ref_image_view { uint32_t* n_contrib; uint32_t* ranges; };

void ref_geom_layout(char* buf, int P, ref_geom_view* v) {
  CudaRasterizer::GeometryState s = CudaRasterizer::GeometryState::fromChunk(buf, P);
  v->depths = s.depths; v->clamped = (unsigned char*)s.clamped; v->means2D = (float*)s.means2D; v->cov3D = s.cov3D;
  v->conic_opacity = (float*)s.conic_opacity; v->rgb = s.rgb; v->norm3D = s.norm3D;
  v->tiles_touched = s.tiles_touched; v->point_offsets = s.point_offsets;
}
void ref_binning_layout(char* buf, int R, ref_binning_view* v) {
  CudaRasterizer::BinningState s = CudaRasterizer::BinningState::fromChunk(buf, R);
  v->keys_unsorted = s.point_list_keys_unsorted; v->keys = s.point_list_keys;
  v->list_unsorted = s.point_list_unsorted; v->list = s.point_list;
}
void ref_image_layout(char* buf, int N, ref_image_view* v) {
  CudaRasterizer::ImageState s = CudaRasterizer::ImageState::fromChunk(buf, N);
  v->n_contrib = s.n_contrib; v->ranges = (uint32_t*)s.ranges;
}

int ref_copy_d2d(void* dst, const void* src, size_t bytes) { return (int)cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice); }

void ref_knn(int P, float* points, float* mean_dists) { SimpleKNN::knn(P, (float3*)points, mean_dists); }

}  // extern "C"
