// temporary stubs (replaced by sfgs_ssim.cu / sfgs_knn.cu)
#include "sfgs_common.cuh"
extern "C" {
int sfgs_fusedssim_forward(float, float, int, int, int, int, const float*, const float*, int, float*, float*, float*, float*, void*) { return SFGS_E_UNSUPPORTED; }
int sfgs_fusedssim_backward(float, float, int, int, int, int, const float*, const float*, const float*, const float*, const float*, const float*, float*, void*) { return SFGS_E_UNSUPPORTED; }
int sfgs_dist2_knn3(int, const float*, float*, sfgs_alloc_fn, void*, void*) { return SFGS_E_UNSUPPORTED; }
}
