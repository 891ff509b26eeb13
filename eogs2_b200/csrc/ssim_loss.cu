// Photometric loss of EOGS++ as ONE forward and ONE backward kernel:
//     L = (1 - lambda) * mean|img - gt| + lambda * (1 - SSIM(img, gt))          (loss/shadow.py:21-29,
//         utils/loss_utils.py:18-85: 11x11 Gaussian window sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2)
// The reference evaluates it with five depthwise F.conv2d calls, ~15 elementwise kernels and their
// autograd.  Here a 16x16-pixel block stages the 26x26 neighbourhood of
// both images in shared memory, runs the separable window (horizontal then vertical, 5 moments at
// once), forms the SSIM map and reduces both sums to two device scalars; it also stores the three
// per-pixel partial derivatives of the map (w.r.t. mu1, E[x^2], E[xy]) so that the backward is again a
// single separable convolution of three maps, finished per pixel as
//     dL/dimg = g * [ (1 - lambda)/N * sign(img - gt) - lambda/N * (W*d_mu1 + 2 img W*d_s11 + gt W*d_s12) ]
// written straight in the rasterizer's planar [C,H,W] layout (it is the dL_dpix the blend backward reads).
//
// HBM-bound: forward reads 8 B and writes 12 B per pixel-channel, backward reads 20 B and writes 4 B;
// ~150 FMA per pixel-channel from shared memory.
#include "common.cuh"

namespace eogs {

constexpr int SS_T = 16;                 // output tile
constexpr int SS_R = 5;                  // window radius (window_size 11)
constexpr int SS_P = SS_T + 2 * SS_R;    // 26: staged neighbourhood
constexpr float SS_C1 = 0.01f * 0.01f, SS_C2 = 0.03f * 0.03f;

struct Window { float w[2 * SS_R + 1]; };

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) t = s_red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    }
    __syncthreads();
    return t;                            // valid in thread 0
}

__global__ void __launch_bounds__(SS_T * SS_T)
photometric_fwd_kernel(int H, int W, Window win, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ maps, float* __restrict__ sums)
{
    __shared__ float s_a[SS_P][SS_P + 1], s_b[SS_P][SS_P + 1];
    __shared__ float s_h[5][SS_P][SS_T + 1];
    __shared__ float s_red[8];
    const int c = blockIdx.z, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W;
    const float* pa = img + c * plane;
    const float* pb = gt + c * plane;
    for (int i = threadIdx.x; i < SS_P * SS_P; i += SS_T * SS_T) {
        const int ly = i / SS_P, lx = i - ly * SS_P;
        const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;          // zero padding (conv2d padding=5)
        s_a[ly][lx] = in ? __ldg(pa + (size_t)gy * W + gx) : 0.f;
        s_b[ly][lx] = in ? __ldg(pb + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_P * SS_T; i += SS_T * SS_T) {       // horizontal pass, 26 rows x 16 columns
        const int ly = i / SS_T, lx = i - ly * SS_T;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; k++) {
            const float a = s_a[ly][lx + k], b = s_b[ly][lx + k], w = win.w[k];
            m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
            s11 = fmaf(w, a * a, s11); s22 = fmaf(w, b * b, s22); s12 = fmaf(w, a * b, s12);
        }
        s_h[0][ly][lx] = m1; s_h[1][ly][lx] = m2; s_h[2][ly][lx] = s11; s_h[3][ly][lx] = s22; s_h[4][ly][lx] = s12;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * SS_R; k++) {                                // vertical pass
        const float w = win.w[k];
        mu1 = fmaf(w, s_h[0][ty + k][tx], mu1); mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
        e11 = fmaf(w, s_h[2][ty + k][tx], e11); e22 = fmaf(w, s_h[3][ty + k][tx], e22);
        e12 = fmaf(w, s_h[4][ty + k][tx], e12);
    }
    const int gx = x0 + tx, gy = y0 + ty;
    const bool inside = gx < W && gy < H;
    float ssim_v = 0.f, l1_v = 0.f;
    if (inside) {
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float sig1 = e11 - mu1_sq, sig2 = e22 - mu2_sq, sig12 = e12 - mu12;
        const float A = mu1_sq + mu2_sq + SS_C1, B = sig1 + sig2 + SS_C2;
        const float Cn = 2.f * mu12 + SS_C1, D = 2.f * sig12 + SS_C2;
        const float inv_AB = 1.f / (A * B);
        const float m = Cn * D * inv_AB;
        ssim_v = m;
        // partial derivatives of the map w.r.t. (mu1, E[x^2], E[xy]) at this pixel (mu2, E[y^2] fixed)
        const float d_mu1 = 2.f * mu2 * (D - Cn) * inv_AB - m * 2.f * mu1 * (B - A) * inv_AB;
        const float d_s11 = -m / B;
        const float d_s12 = 2.f * Cn * inv_AB;
        const size_t o = c * plane + (size_t)gy * W + gx;
        const size_t stride = (size_t)gridDim.z * plane;
        maps[o] = d_mu1; maps[stride + o] = d_s11; maps[2 * stride + o] = d_s12;
        l1_v = fabsf(s_a[ty + SS_R][tx + SS_R] - s_b[ty + SS_R][tx + SS_R]);
    }
    const float bs = block_sum_256(ssim_v, s_red);
    const float bl = block_sum_256(l1_v, s_red);
    if (threadIdx.x == 0) { atomicAdd(sums, bs); atomicAdd(sums + 1, bl); }
}

__global__ void photometric_finish_kernel(float inv_n, float lambda, const float* __restrict__ sums, float* __restrict__ out) {
    const float ssim_mean = sums[0] * inv_n, l1_mean = sums[1] * inv_n;
    out[0] = (1.f - lambda) * l1_mean + lambda * (1.f - ssim_mean);
    out[1] = ssim_mean;
    out[2] = l1_mean;
}

__global__ void __launch_bounds__(SS_T * SS_T)
photometric_bwd_kernel(int H, int W, Window win, float inv_n, float lambda, const float* __restrict__ img,
                       const float* __restrict__ gt, const float* __restrict__ maps,
                       const float* __restrict__ dL_dloss, float* __restrict__ dL_dimg)
{
    __shared__ float s_m[3][SS_P][SS_P + 1];
    __shared__ float s_h[3][SS_P][SS_T + 1];
    const int c = blockIdx.z, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W, stride = (size_t)gridDim.z * plane;
    for (int i = threadIdx.x; i < SS_P * SS_P; i += SS_T * SS_T) {
        const int ly = i / SS_P, lx = i - ly * SS_P;
        const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = c * plane + (size_t)gy * W + gx;
#pragma unroll
        for (int k = 0; k < 3; k++) s_m[k][ly][lx] = in ? __ldg(maps + k * stride + o) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_P * SS_T; i += SS_T * SS_T) {
        const int ly = i / SS_T, lx = i - ly * SS_T;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; k++) {
            const float w = win.w[k];
            a0 = fmaf(w, s_m[0][ly][lx + k], a0); a1 = fmaf(w, s_m[1][ly][lx + k], a1); a2 = fmaf(w, s_m[2][ly][lx + k], a2);
        }
        s_h[0][ly][lx] = a0; s_h[1][ly][lx] = a1; s_h[2][ly][lx] = a2;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * SS_R; k++) {
        const float w = win.w[k];
        c0 = fmaf(w, s_h[0][ty + k][tx], c0); c1 = fmaf(w, s_h[1][ty + k][tx], c1); c2 = fmaf(w, s_h[2][ty + k][tx], c2);
    }
    const size_t o = c * plane + (size_t)gy * W + gx;
    const float a = __ldg(img + o), b = __ldg(gt + o);
    const float d_ssim = c0 + 2.f * a * c1 + b * c2;
    const float diff = a - b;
    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);       // torch.abs backward: sign(0) = 0
    const float g = dL_dloss ? __ldg(dL_dloss) : 1.f;
    dL_dimg[o] = g * inv_n * ((1.f - lambda) * sgn - lambda * d_ssim);
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_photometric_forward(eogs_stream_t stream, int C, int H, int W, const float* window11,
                                      const float* image, const float* gt, float lambda_dssim,
                                      float* maps, float* sums2, float* out3)
{
    if (C <= 0 || H <= 0 || W <= 0) { set_error("bad sizes C=%d H=%d W=%d", C, H, W); return -1; }
    if (!window11 || !image || !gt || !maps || !sums2 || !out3) { set_error("null argument"); return -4; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Window win;
    for (int k = 0; k < 2 * SS_R + 1; k++) win.w[k] = window11[k];      // host array
    EOGS_CUDA(cudaMemsetAsync(sums2, 0, 2 * sizeof(float), s));
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C);
    photometric_fwd_kernel<<<grid, SS_T * SS_T, 0, s>>>(H, W, win, image, gt, maps, sums2);
    EOGS_LAUNCH_CHECK("photometric_fwd_kernel");
    photometric_finish_kernel<<<1, 1, 0, s>>>(1.f / ((float)C * (float)H * (float)W), lambda_dssim, sums2, out3);
    EOGS_LAUNCH_CHECK("photometric_finish_kernel");
    return 0;
}

EOGS_API int eogs_photometric_backward(eogs_stream_t stream, int C, int H, int W, const float* window11,
                                       const float* image, const float* gt, float lambda_dssim,
                                       const float* maps, const float* dL_dloss, float* dL_dimage)
{
    if (C <= 0 || H <= 0 || W <= 0) { set_error("bad sizes C=%d H=%d W=%d", C, H, W); return -1; }
    if (!window11 || !image || !gt || !maps || !dL_dimage) { set_error("null argument"); return -4; }
    Window win;
    for (int k = 0; k < 2 * SS_R + 1; k++) win.w[k] = window11[k];
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C);
    photometric_bwd_kernel<<<grid, SS_T * SS_T, 0, static_cast<cudaStream_t>(stream)>>>(
        H, W, win, 1.f / ((float)C * (float)H * (float)W), lambda_dssim, image, gt, maps, dL_dloss, dL_dimage);
    EOGS_LAUNCH_CHECK("photometric_bwd_kernel");
    return 0;
}

}  // extern "C"
