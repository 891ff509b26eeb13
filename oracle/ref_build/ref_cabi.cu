// C ABI around the reference's CudaRasterizer::Rasterizer (DGR/cuda_rasterizer/rasterizer.h:24-92),
// so that tests and bench.py --impl reference can drive the UNMODIFIED reference kernels through
// ctypes without compiling DGR/rasterize_points.cu against torch headers (5 minutes of nvcc).
// TEST INFRASTRUCTURE: nothing under eogs2_b200/ links or loads this.
//
// The three growable byte buffers of DGR/rasterize_points.cu:78-85 are provided by the caller
// through `alloc(which, bytes)` (which: 0 geometry, 1 binning, 2 image), like the reference's
// resizeFunctional lambdas.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

typedef char* (*ref_alloc_fn)(int which, size_t bytes);

static void copy_err(char* buf, int len, const char* msg) {
    if (buf && len > 0) { strncpy(buf, msg, len - 1); buf[len - 1] = 0; }
}

extern "C" {

int eogs_ref_forward(int P, const float* bg, int W, int H, const float* means3D, const float* colors,
                     const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                     const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                     int prefiltered, float* out_color, float* out_invdepth, int antialiasing,
                     int* radii, int debug, ref_alloc_fn alloc, char* errbuf, int errlen)
{
    try {
        std::function<char*(size_t)> g = [alloc](size_t n) { return alloc(0, n); };
        std::function<char*(size_t)> b = [alloc](size_t n) { return alloc(1, n); };
        std::function<char*(size_t)> i = [alloc](size_t n) { return alloc(2, n); };
        return CudaRasterizer::Rasterizer::forward(
            g, b, i, P, /*D=*/0, /*M=*/0, bg, W, H, means3D, /*shs=*/nullptr, colors, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
            prefiltered != 0, out_color, out_invdepth, antialiasing != 0, radii, debug != 0);
    } catch (const std::exception& e) {
        copy_err(errbuf, errlen, e.what());
        return -1;
    }
}

int eogs_ref_backward(int P, int R, const float* bg, int W, int H, const float* means3D,
                      const float* colors, const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* campos,
                      float tanfovx, float tanfovy, const int* radii, char* geom, char* binning,
                      char* image, const float* dL_dpix, const float* dL_dinvdepths,
                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                      float* dL_dscale, float* dL_drot, int antialiasing, int debug, float* dL_dT,
                      char* errbuf, int errlen)
{
    try {
        CudaRasterizer::Rasterizer::backward(
            P, 0, 0, R, bg, W, H, means3D, nullptr, colors, opacities, scales, scale_modifier, rotations,
            cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy, radii, geom, binning, image,
            dL_dpix, dL_dinvdepths, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dinvdepth,
            dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, antialiasing != 0, debug != 0, dL_dT);
        return 0;
    } catch (const std::exception& e) {
        copy_err(errbuf, errlen, e.what());
        return -1;
    }
}

void eogs_ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

// Byte offsets (relative to `base`) of the fields inside the reference's opaque buffers,
// obtained by running its own fromChunk on the pointer (rasterizer_impl.cu:155-194).
void eogs_ref_geom_layout(char* base, size_t P, uint64_t* out /*[7]*/) {
    char* p = base;
    CudaRasterizer::GeometryState s = CudaRasterizer::GeometryState::fromChunk(p, P);
    out[0] = (char*)s.depths - base;         out[1] = (char*)s.internal_radii - base;
    out[2] = (char*)s.means2D - base;        out[3] = (char*)s.cov3D - base;
    out[4] = (char*)s.conic_opacity - base;  out[5] = (char*)s.tiles_touched - base;
    out[6] = (char*)s.point_offsets - base;
}
void eogs_ref_binning_layout(char* base, size_t R, uint64_t* out /*[4]*/) {
    char* p = base;
    CudaRasterizer::BinningState s = CudaRasterizer::BinningState::fromChunk(p, R);
    out[0] = (char*)s.point_list - base;       out[1] = (char*)s.point_list_unsorted - base;
    out[2] = (char*)s.point_list_keys - base;  out[3] = (char*)s.point_list_keys_unsorted - base;
}
void eogs_ref_image_layout(char* base, size_t N, uint64_t* out /*[3]*/) {
    char* p = base;
    CudaRasterizer::ImageState s = CudaRasterizer::ImageState::fromChunk(p, N);
    out[0] = (char*)s.accum_alpha - base;  out[1] = (char*)s.n_contrib - base;
    out[2] = (char*)s.ranges - base;
}

}  // extern "C"
