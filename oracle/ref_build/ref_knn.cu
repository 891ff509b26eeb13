// C ABI around the reference's SimpleKNN::knn (submodules/simple-knn/simple_knn.h:16-20), compiled from
// simple_knn.cu WHERE IT LIES under /root/reference (nothing copied) so that tests and the golden
// generator can drive the unmodified reference kernels through ctypes without the torch binding
// (spatial.cu / ext.cpp).  TEST INFRASTRUCTURE: nothing under eogs2_b200/ links or loads this.
// <cstdint> first: simple_knn.cu uses uint32_t without including it (gcc 13).
#include <cstdint>
#include "simple_knn.cu"

extern "C" int eogs_ref_knn(int P, float* points, float* mean_dists)
{
    // spatial.cu:22-24: means = full({P}, 0.0); SimpleKNN::knn(P, points, means)
    SimpleKNN::knn(P, (float3*)points, mean_dists);
    return (int)cudaDeviceSynchronize();
}
