// temporary stub (replaced by sfgs_knn.cu)
#include "sfgs_common.cuh"
extern "C" {
int sfgs_dist2_knn3(int, const float*, float*, sfgs_alloc_fn, void*, void*) { return SFGS_E_UNSUPPORTED; }
}
