// rasterize_points_eogs.cpp — what the reference's pybind module `_C` would bind instead of
// DGR/rasterize_points.cu + DGR/ext.cpp (DGR = submodules/diff-gaussian-rasterization of gardiens/EOGS2), for a
// maintainer who keeps DGR/diff_gaussian_rasterization/__init__.py byte for byte and only swaps the native side.
// Compiled by g++ (no nvcc), linked against libeogs_raster.so.  tests/test_integration_stub.py syntax-checks it
// against the torch headers; INTEGRATION.md section 2 explains the mapping.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include "eogs_raster.h"

static void* alloc_cb(void* user, int which, size_t bytes) {          // resizeFunctional, rasterize_points.cu:27-33
    auto* bufs = static_cast<torch::Tensor*>(user);                    // [geom, binning, image, point_list]
    bufs[which].resize_({(long long)bytes});
    return bufs[which].data_ptr();
}

static const float* fptr(const torch::Tensor& t) { return t.numel() ? t.data_ptr<float>() : nullptr; }

// RasterizeGaussiansCUDA, rasterize_points.cu:35-124 (same arguments, same 7-tuple)
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered, const bool antialiasing, const bool debug) {
    TORCH_CHECK(means3D.ndimension() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    const int P = means3D.size(0), H = image_height, W = image_width;
    auto f = means3D.options().dtype(torch::kFloat32);
    auto out_color = torch::zeros({5, H, W}, f), out_invdepth = torch::zeros({1, H, W}, f);
    auto radii = torch::zeros({P}, means3D.options().dtype(torch::kInt32));
    auto bytes = torch::TensorOptions(torch::kByte).device(means3D.device());
    torch::Tensor bufs[4] = {torch::empty({0}, bytes), torch::empty({0}, bytes), torch::empty({0}, bytes), torch::empty({0}, bytes)};
    void *geom = nullptr, *image = nullptr; uint32_t* plist = nullptr; uint32_t rendered = 0;
    if (P != 0) {
        const auto m = means3D.contiguous(), s = scales.contiguous(), r = rotations.contiguous(), c3 = cov3D_precomp.contiguous(),
                   o = opacity.contiguous(), c = colors.contiguous(), v = viewmatrix.contiguous(), bg = background.contiguous();
        int rc = eogs_rasterize_forward(c10::cuda::getCurrentCUDAStream().stream(), P, W, H, 5, fptr(m), fptr(s), fptr(r),
                                        fptr(c3), fptr(o), fptr(c), fptr(v), scale_modifier, antialiasing, fptr(bg),
                                        alloc_cb, bufs, radii.data_ptr<int>(), out_color.data_ptr<float>(),
                                        out_invdepth.data_ptr<float>(), &geom, &plist, &image, &rendered);
        TORCH_CHECK(rc == 0, eogs_last_error());
    }
    (void)projmatrix; (void)tan_fovx; (void)tan_fovy; (void)sh; (void)degree; (void)campos; (void)prefiltered; (void)debug;
    // geomBuffer, binningBuffer (= point_list), imgBuffer keep the reference's slots in the returned tuple
    return std::make_tuple((int)rendered, out_color, radii, bufs[0], bufs[3], bufs[2], out_invdepth);
}

// RasterizeGaussiansBackwardCUDA, rasterize_points.cu:126-224.  Same arguments; the returned tuple has the reference's nine
// tensors.  dL_dT is returned as a [1, 6] tensor holding sum_p dL_dT[p] — DGR __init__.py:180-192 only ever sums it
// over p — so __init__.py runs unchanged on it.
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& opacities, const torch::Tensor& scales,
                               const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                               const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_invdepth,
                               const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                               const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer,
                               const torch::Tensor& imageBuffer, const bool antialiasing, const bool debug) {
    const int P = means3D.size(0), H = dL_dout_color.size(1), W = dL_dout_color.size(2);
    auto f = means3D.options().dtype(torch::kFloat32);
    auto dL_dmeans3D = torch::zeros({P, 3}, f), dL_dmeans2D = torch::zeros({P, 3}, f), dL_dcolors = torch::zeros({P, 5}, f);
    auto dL_dopacity = torch::zeros({P, 1}, f), dL_dcov3D = torch::zeros({P, 6}, f), dL_dsh = torch::zeros({P, 0, 3}, f);
    auto dL_dscales = torch::zeros({P, 3}, f), dL_drotations = torch::zeros({P, 4}, f), cam_sums = torch::zeros({16}, f);
    if (P != 0) {
        auto scratch = torch::empty({(long long)eogs_grad_scratch_floats(P)}, f);
        const bool precomp = cov3D_precomp.numel() != 0;
        const auto m = means3D.contiguous(), s = scales.contiguous(), r = rotations.contiguous(), c3 = cov3D_precomp.contiguous(),
                   o = opacities.contiguous(), c = colors.contiguous(), v = viewmatrix.contiguous(), pj = projmatrix.contiguous(),
                   bg = background.contiguous(), g = dL_dout_color.contiguous(), gi = dL_dout_invdepth.contiguous();
        int rc = eogs_backward(c10::cuda::getCurrentCUDAStream().stream(), P, W, H, 5, (uint32_t)R, fptr(m), fptr(s), fptr(r),
                               fptr(c3), fptr(o), fptr(c), fptr(v), fptr(pj), scale_modifier, antialiasing, fptr(bg),
                               radii.data_ptr<int>(), geomBuffer.data_ptr(), reinterpret_cast<const uint32_t*>(binningBuffer.data_ptr()),
                               imageBuffer.data_ptr(), fptr(g), fptr(gi), scratch.data_ptr<float>(),
                               dL_dmeans2D.data_ptr<float>(), dL_dcolors.data_ptr<float>(), dL_dopacity.data_ptr<float>(),
                               dL_dmeans3D.data_ptr<float>(), precomp ? dL_dcov3D.data_ptr<float>() : nullptr,
                               precomp ? nullptr : dL_dscales.data_ptr<float>(), precomp ? nullptr : dL_drotations.data_ptr<float>(),
                               cam_sums.data_ptr<float>());
        TORCH_CHECK(rc == 0, eogs_last_error());
    }
    (void)tan_fovx; (void)tan_fovy; (void)sh; (void)degree; (void)campos; (void)debug;
    auto dL_dT = cam_sums.slice(0, 0, 6).reshape({1, 6}).clone();
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dT);
}

// markVisible, rasterize_points.cu:226-245
torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    const int P = means3D.size(0);
    auto present = torch::full({P}, false, means3D.options().dtype(torch::kBool));
    if (P > 0) {
        int rc = eogs_mark_visible(c10::cuda::getCurrentCUDAStream().stream(), P, means3D.contiguous().data_ptr<float>(),
                                   viewmatrix.contiguous().data_ptr<float>(), projmatrix.contiguous().data_ptr<float>(),
                                   reinterpret_cast<uint8_t*>(present.data_ptr<bool>()));
        TORCH_CHECK(rc == 0, eogs_last_error());
    }
    return present;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {                             // DGR/ext.cpp:15-18
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
}
