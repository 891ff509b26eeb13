// DSM splat: the `plyflatten` step of compute_dsm_from_view (utils/dsm_utils.py:27-37, tsdf.py:578-580) on the
// GPU.  The reference calls the third-party package `plyflatten`
// (requirements.txt:18, unpinned, absent from /root/reference and from this image), whose published
// algorithm (plyflatten.c `rasterize_cloud`) is restated in oracle/eogs_oracle.c:oracle_plyflatten:
//   for every point (x, y, v): cell i = (int)(w (x - xoff) / (w res)), j = (int)(h (-y + yoff) / (h res));
//   skipped when outside [0,w) x [0,h); for every cell (i+k1, j+k2), |k1|,|k2| <= radius, inside the raster:
//   weight = exp(-dist^2 / (2 sigma^2)), dist = hypot to the cell centre (fp32); the cell keeps the
//   weighted running average of v; cells nobody reached are NaN.
// The running average is sequential in the reference; here a cell accumulates sum(w v) and sum(w) with
// fp64 atomics (order-independent to ~1e-16) and one finalising pass divides: same value up to the fp32
// rounding of the reference's running mean.  HBM/L2-atomic bound: 24 B read per point, (2r+1)^2 x 2 atomics.
#include "common.cuh"
#include <math.h>

namespace eogs {

__device__ __forceinline__ int dsm_rescale(double x, double mn, double mx, int w, bool& inside) {
    const int r = (int)(w * (x - mn) / (mx - mn));        // C conversion: truncation toward zero
    inside = r >= 0 && r < w;
    return r;
}

__global__ void __launch_bounds__(256)
dsm_splat_kernel(long long N, const double* __restrict__ cloud, double xoff, double yoff, double resolution,
                 int w, int h, int radius, float sigma, double* __restrict__ accum)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double xx = cloud[3 * k], yy = cloud[3 * k + 1];
    const float v = (float)cloud[3 * k + 2];
    bool in_x, in_y;
    const int i = dsm_rescale(xx, xoff, xoff + w * resolution, w, in_x);
    const int j = dsm_rescale(-yy, -yoff, -yoff + h * resolution, h, in_y);
    if (!in_x || !in_y) return;
    const double sigma2mult2 = 2.0 * (double)sigma * (double)sigma;
    for (int k1 = -radius; k1 <= radius; k1++)
        for (int k2 = -radius; k2 <= radius; k2++) {
            const int ii = i + k1, jj = j + k2;
            if (ii < 0 || ii >= w || jj < 0 || jj >= h) continue;
            const float dist_x = (float)(xx - (xoff + resolution * (0.5 + ii)));
            const float dist_y = (float)(yy - (yoff - resolution * (0.5 + jj)));
            const float dist = hypotf(dist_x, dist_y);
            const float weight = (float)exp(-(double)(dist * dist) / sigma2mult2);
            const size_t cell = (size_t)jj * w + ii;
            atomicAdd(&accum[2 * cell], (double)weight * (double)v);
            atomicAdd(&accum[2 * cell + 1], (double)weight);
        }
}

__global__ void __launch_bounds__(256)
dsm_finalize_kernel(size_t cells, const double* __restrict__ accum, float* __restrict__ raster)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const double sw = accum[2 * c + 1];
    raster[c] = sw != 0.0 ? (float)(accum[2 * c] / sw) : __int_as_float(0x7fc00000);
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_dsm_splat(eogs_stream_t stream, long long N, const double* cloud, double xoff, double yoff,
                            double resolution, int xsize, int ysize, int radius, float sigma,
                            double* accum, float* raster)
{
    if (N < 0 || xsize <= 0 || ysize <= 0 || radius < 0 || !(resolution > 0.0)) { set_error("bad sizes"); return -1; }
    if (!accum || !raster || (N > 0 && !cloud)) { set_error("null argument"); return -4; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t cells = (size_t)xsize * ysize;
    EOGS_CUDA(cudaMemsetAsync(accum, 0, cells * 2 * sizeof(double), s));
    if (N > 0) {
        dsm_splat_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(N, cloud, xoff, yoff, resolution, xsize, ysize,
                                                                     radius, sigma, accum);
        EOGS_LAUNCH_CHECK("dsm_splat_kernel");
    }
    dsm_finalize_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(cells, accum, raster);
    EOGS_LAUNCH_CHECK("dsm_finalize_kernel");
    return 0;
}

}  // extern "C"
