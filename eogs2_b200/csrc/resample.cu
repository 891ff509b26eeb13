// Sun-view / virtual-camera resample: reproject the camera's (u, v, altitude) grid into a virtual
// camera's render and sample it bilinearly — ONE kernel forward, ONE backward.
// Replaces, in render_resample_virtual_camera (gaussian_renderer/renderer_cc_shadow.py:28-46):
//     virtual_uv = einsum("...ij,...j->...i", cam2virt, rendered_uva)[..., :2]
//     sample     = F.grid_sample(virtual_render[None], virtual_uv[None], align_corners=True)[0]
//     rgb, altitude = sample[:3], sample[3];  altitude[(virtual_uv.abs() > 1).any(-1)] = -100
// (three torch ops, a 5-channel sample of which one channel is dropped, a boolean mask and an
// indexed write) and their autograd.  Semantics follow ATen's grid_sampler_2d (bilinear, zeros
// padding, align_corners=True): x = (u + 1)/2 (Wv - 1), taps at floor(x), floor(x)+1, out-of-bounds
// taps contribute nothing.
//
// HBM / L2-bound gather.  Algorithmic bytes per output pixel: 12 B uva in, 4 taps x 4 channels x 4 B
// gathered (neighbouring pixels share sectors), 16 + 8 B out; backward the same plus 16 float
// reductions into the virtual image's gradient (red.global.add.f32) and 12 B of d_uva.
#include "common.cuh"

namespace eogs {

constexpr int RS_THREADS = 256;
constexpr int RS_CH = 4;                 // rgb + altitude are sampled; the opacity channel is not used

struct Taps {
    int x0, y0;                          // north-west tap
    float wx1, wy1;                      // x - x0, y - y0 (weights of the +1 taps)
    bool in_x0, in_x1, in_y0, in_y1;
    float u, v;
};

__device__ __forceinline__ Taps make_taps(const float* __restrict__ M, float a, float b, float c, int Wv, int Hv) {
    Taps t;
    t.u = fmaf(M[2], c, fmaf(M[1], b, M[0] * a));
    t.v = fmaf(M[5], c, fmaf(M[4], b, M[3] * a));
    const float x = (t.u + 1.f) * 0.5f * (float)(Wv - 1);
    const float y = (t.v + 1.f) * 0.5f * (float)(Hv - 1);
    const float fx = floorf(x), fy = floorf(y);
    t.x0 = (int)fx; t.y0 = (int)fy;
    t.wx1 = x - fx; t.wy1 = y - fy;
    // NaN / huge coordinates: the float -> int conversion saturates and every tap is out of bounds
    t.in_x0 = t.x0 >= 0 && t.x0 < Wv; t.in_x1 = t.x0 + 1 >= 0 && t.x0 + 1 < Wv;
    t.in_y0 = t.y0 >= 0 && t.y0 < Hv; t.in_y1 = t.y0 + 1 >= 0 && t.y0 + 1 < Hv;
    return t;
}

__global__ void __launch_bounds__(RS_THREADS)
resample_fwd_kernel(int Hv, int Wv, int npix, const float* __restrict__ virt, const float* __restrict__ cam2virt,
                    const float* __restrict__ uva, float* __restrict__ out_rgb, float* __restrict__ out_alt,
                    float* __restrict__ out_uv)
{
    __shared__ float M[9];
    if (threadIdx.x < 9) M[threadIdx.x] = __ldg(cam2virt + threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * RS_THREADS + threadIdx.x;
    if (i >= npix) return;
    const Taps t = make_taps(M, __ldg(uva + 3 * (size_t)i), __ldg(uva + 3 * (size_t)i + 1), __ldg(uva + 3 * (size_t)i + 2), Wv, Hv);
    const float w00 = (1.f - t.wx1) * (1.f - t.wy1), w10 = t.wx1 * (1.f - t.wy1);
    const float w01 = (1.f - t.wx1) * t.wy1, w11 = t.wx1 * t.wy1;
    const size_t plane = (size_t)Hv * Wv;
    const size_t o00 = (size_t)t.y0 * Wv + t.x0;
    float s[RS_CH];
#pragma unroll
    for (int c = 0; c < RS_CH; c++) {
        const float* p = virt + c * plane;
        float acc = 0.f;
        if (t.in_y0 && t.in_x0) acc = fmaf(__ldg(p + o00), w00, acc);
        if (t.in_y0 && t.in_x1) acc = fmaf(__ldg(p + o00 + 1), w10, acc);
        if (t.in_y1 && t.in_x0) acc = fmaf(__ldg(p + o00 + Wv), w01, acc);
        if (t.in_y1 && t.in_x1) acc = fmaf(__ldg(p + o00 + Wv + 1), w11, acc);
        s[c] = acc;
    }
    const bool outside = fabsf(t.u) > 1.f || fabsf(t.v) > 1.f;
    out_rgb[i] = s[0]; out_rgb[(size_t)npix + i] = s[1]; out_rgb[2 * (size_t)npix + i] = s[2];
    out_alt[i] = outside ? -100.f : s[3];
    out_uv[2 * (size_t)i] = t.u; out_uv[2 * (size_t)i + 1] = t.v;
}

__global__ void __launch_bounds__(RS_THREADS)
resample_bwd_kernel(int Hv, int Wv, int npix, const float* __restrict__ virt, const float* __restrict__ cam2virt,
                    const float* __restrict__ uva, const float* __restrict__ d_rgb, const float* __restrict__ d_alt,
                    const float* __restrict__ d_uv_in, float* __restrict__ d_virt, float* __restrict__ d_uva,
                    float* __restrict__ d_cam2virt)
{
    __shared__ float M[9];
    __shared__ float s_part[RS_THREADS / 32][6];
    if (threadIdx.x < 9) M[threadIdx.x] = __ldg(cam2virt + threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * RS_THREADS + threadIdx.x;
    float sums[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (i < npix) {
        const float a = __ldg(uva + 3 * (size_t)i), b = __ldg(uva + 3 * (size_t)i + 1), c3 = __ldg(uva + 3 * (size_t)i + 2);
        const Taps t = make_taps(M, a, b, c3, Wv, Hv);
        const bool outside = fabsf(t.u) > 1.f || fabsf(t.v) > 1.f;
        float g[RS_CH];
        g[0] = d_rgb ? __ldg(d_rgb + i) : 0.f;
        g[1] = d_rgb ? __ldg(d_rgb + (size_t)npix + i) : 0.f;
        g[2] = d_rgb ? __ldg(d_rgb + 2 * (size_t)npix + i) : 0.f;
        g[3] = (d_alt && !outside) ? __ldg(d_alt + i) : 0.f;      // the -100 overwrite cuts the gradient
        const float w00 = (1.f - t.wx1) * (1.f - t.wy1), w10 = t.wx1 * (1.f - t.wy1);
        const float w01 = (1.f - t.wx1) * t.wy1, w11 = t.wx1 * t.wy1;
        const size_t plane = (size_t)Hv * Wv;
        const size_t o00 = (size_t)t.y0 * Wv + t.x0;
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int c = 0; c < RS_CH; c++) {
            const float* p = virt + c * plane;
            float* dp = d_virt + c * plane;
            const float gc = g[c];
            if (t.in_y0 && t.in_x0) { const float v = __ldg(p + o00);          atomicAdd(dp + o00, w00 * gc);          gix -= v * (1.f - t.wy1) * gc; giy -= v * (1.f - t.wx1) * gc; }
            if (t.in_y0 && t.in_x1) { const float v = __ldg(p + o00 + 1);      atomicAdd(dp + o00 + 1, w10 * gc);      gix += v * (1.f - t.wy1) * gc; giy -= v * t.wx1 * gc; }
            if (t.in_y1 && t.in_x0) { const float v = __ldg(p + o00 + Wv);     atomicAdd(dp + o00 + Wv, w01 * gc);     gix -= v * t.wy1 * gc;         giy += v * (1.f - t.wx1) * gc; }
            if (t.in_y1 && t.in_x1) { const float v = __ldg(p + o00 + Wv + 1); atomicAdd(dp + o00 + Wv + 1, w11 * gc); gix += v * t.wy1 * gc;         giy += v * t.wx1 * gc; }
        }
        float du = gix * 0.5f * (float)(Wv - 1), dv = giy * 0.5f * (float)(Hv - 1);
        if (d_uv_in) { du += __ldg(d_uv_in + 2 * (size_t)i); dv += __ldg(d_uv_in + 2 * (size_t)i + 1); }
        d_uva[3 * (size_t)i] = M[0] * du + M[3] * dv;
        d_uva[3 * (size_t)i + 1] = M[1] * du + M[4] * dv;
        d_uva[3 * (size_t)i + 2] = M[2] * du + M[5] * dv;
        sums[0] = du * a; sums[1] = du * b; sums[2] = du * c3;
        sums[3] = dv * a; sums[4] = dv * b; sums[5] = dv * c3;
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float v = sums[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) v += s_part[w][threadIdx.x];
        if (v != 0.f) atomicAdd(d_cam2virt + threadIdx.x, v);
    }
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_resample_forward(eogs_stream_t stream, int Cv, int Hv, int Wv, int H, int W,
                                   const float* virtual_render, const float* cam2virt, const float* rendered_uva,
                                   float* out_rgb, float* out_altitude, float* out_uv)
{
    if (Cv < RS_CH || Hv <= 0 || Wv <= 0 || H <= 0 || W <= 0) { set_error("bad sizes Cv=%d Hv=%d Wv=%d H=%d W=%d", Cv, Hv, Wv, H, W); return -1; }
    if (!virtual_render || !cam2virt || !rendered_uva || !out_rgb || !out_altitude || !out_uv) { set_error("null argument"); return -4; }
    const int npix = H * W;
    resample_fwd_kernel<<<(npix + RS_THREADS - 1) / RS_THREADS, RS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        Hv, Wv, npix, virtual_render, cam2virt, rendered_uva, out_rgb, out_altitude, out_uv);
    EOGS_LAUNCH_CHECK("resample_fwd_kernel");
    return 0;
}

EOGS_API int eogs_resample_backward(eogs_stream_t stream, int Cv, int Hv, int Wv, int H, int W,
                                    const float* virtual_render, const float* cam2virt, const float* rendered_uva,
                                    const float* dL_drgb, const float* dL_daltitude, const float* dL_duv,
                                    float* dL_dvirtual, float* dL_duva, float* dL_dcam2virt)
{
    if (Cv < RS_CH || Hv <= 0 || Wv <= 0 || H <= 0 || W <= 0) { set_error("bad sizes Cv=%d Hv=%d Wv=%d H=%d W=%d", Cv, Hv, Wv, H, W); return -1; }
    if (!virtual_render || !cam2virt || !rendered_uva || !dL_dvirtual || !dL_duva || !dL_dcam2virt) { set_error("null argument"); return -4; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    EOGS_CUDA(cudaMemsetAsync(dL_dvirtual, 0, (size_t)Cv * Hv * Wv * sizeof(float), s));
    EOGS_CUDA(cudaMemsetAsync(dL_dcam2virt, 0, 9 * sizeof(float), s));
    const int npix = H * W;
    resample_bwd_kernel<<<(npix + RS_THREADS - 1) / RS_THREADS, RS_THREADS, 0, s>>>(
        Hv, Wv, npix, virtual_render, cam2virt, rendered_uva, dL_drgb, dL_daltitude, dL_duv, dL_dvirtual, dL_duva,
        dL_dcam2virt);
    EOGS_LAUNCH_CHECK("resample_bwd_kernel");
    return 0;
}

}  // extern "C"
