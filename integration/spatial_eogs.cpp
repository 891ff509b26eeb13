// spatial_eogs.cpp — replaces submodules/simple-knn/spatial.cu (distCUDA2, :15-26) and simple_knn.cu
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include "eogs_raster.h"

torch::Tensor distCUDA2(const torch::Tensor& points) {
    const int P = points.size(0);
    auto means = torch::full({P}, 0.0, points.options().dtype(torch::kFloat32));
    auto scratch = torch::empty({(long long)eogs_knn_bytes(P)}, points.options().dtype(torch::kByte));
    int rc = eogs_knn_dist2(c10::cuda::getCurrentCUDAStream().stream(), P, points.contiguous().data_ptr<float>(),
                            scratch.data_ptr(), scratch.numel(), means.data_ptr<float>());
    TORCH_CHECK(rc == 0, eogs_last_error());
    return means;
}
