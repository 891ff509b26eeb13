// Photometric loss of EOGS++ as ONE forward and ONE backward kernel:
//     L = (1 - lambda) * mean|img - gt| + lambda * (1 - SSIM(img, gt))          (loss/shadow.py:21-29,
//         utils/loss_utils.py:18-85: 11x11 Gaussian window sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2)
// The reference evaluates it with five depthwise F.conv2d calls, ~15 elementwise kernels and their
// autograd.  Here a 32x32-pixel block stages the 42x42 neighbourhood of
// both images in shared memory, runs the separable window (horizontal then vertical, 5 moments at
// once), forms the SSIM map and reduces both sums to two device scalars; it also stores the three
// per-pixel partial derivatives of the map (w.r.t. mu1, E[x^2], E[xy]) so that the backward is again a
// single separable convolution of three maps, finished per pixel as
//     dL/dimg = g * [ (1 - lambda)/N * sign(img - gt) - lambda/N * (W*d_mu1 + 2 img W*d_s11 + gt W*d_s12) ]
// written straight in the rasterizer's planar [C,H,W] layout (it is the dL_dpix the blend backward reads).
//
// Bound: shared-memory instruction issue, not HBM (ncu of the first version: MIO throttle the top stall, DRAM 7-11 %).
// Both passes are therefore REGISTER-TILED: a thread produces 4 adjacent outputs from a sliding window it holds in
// registers — the horizontal pass reads its 14 + 2 inputs per image as four LDS.128, the vertical pass reads 14
// values per map for 4 outputs — 21 shared-memory instructions per output pixel instead of 77, and the halo
// amplification of the staged tile drops from 2.6x (16x16) to 1.7x (32x32).
// HBM: forward reads 8 B and writes 12 B per pixel-channel, backward reads 20 B and writes 4 B.
#include "common.cuh"

namespace eogs {

constexpr int SS_T = 32;                 // output tile (square)
constexpr int SS_R = 5;                  // window radius (window_size 11)
constexpr int SS_P = SS_T + 2 * SS_R;    // 42: staged neighbourhood
constexpr int SS_PITCH = 44;             // staged row pitch in floats: 16-byte aligned rows, columns 42..43 are zero
constexpr int SS_THREADS = 256;
constexpr int SS_Q = 4;                  // outputs per thread and pass
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

// Stage the 42x42 neighbourhood of one plane (zero padding outside the image, conv2d padding=5) into rows of
// SS_PITCH floats; the two pad columns are zero.
__device__ __forceinline__ void stage_plane(const float* __restrict__ src, int H, int W, int x0, int y0,
                                            float (*dst)[SS_PITCH]) {
    for (int i = threadIdx.x; i < SS_P * SS_PITCH; i += SS_THREADS) {
        const int ly = i / SS_PITCH, lx = i - ly * SS_PITCH;
        const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
        const bool in = lx < SS_P && gy >= 0 && gy < H && gx >= 0 && gx < W;
        dst[ly][lx] = in ? __ldg(src + (size_t)gy * W + gx) : 0.f;
    }
}

// 16 consecutive floats of a staged row, starting at a multiple of 4: four conflict-free LDS.128
__device__ __forceinline__ void load16(const float* row, float* v) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 q = *reinterpret_cast<const float4*>(row + 4 * k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}

__global__ void __launch_bounds__(SS_THREADS)
photometric_fwd_kernel(int H, int W, Window win, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ maps, float* __restrict__ sums)
{
    __shared__ __align__(16) float s_a[SS_P][SS_PITCH], s_b[SS_P][SS_PITCH];
    __shared__ __align__(16) float s_h[5][SS_P][SS_T];
    __shared__ float s_red[8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W;
    stage_plane(img + c * plane, H, W, x0, y0, s_a);
    stage_plane(gt + c * plane, H, W, x0, y0, s_b);
    __syncthreads();
    // horizontal pass: 42 rows x 8 items of 4 adjacent outputs
    for (int it = threadIdx.x; it < SS_P * (SS_T / SS_Q); it += SS_THREADS) {
        const int ly = it / (SS_T / SS_Q), x = (it - ly * (SS_T / SS_Q)) * SS_Q;
        float a[16], b[16];
        load16(&s_a[ly][x], a);
        load16(&s_b[ly][x], b);
        float m1[SS_Q], m2[SS_Q], s11[SS_Q], s22[SS_Q], s12[SS_Q];
#pragma unroll
        for (int o = 0; o < SS_Q; o++) { m1[o] = m2[o] = s11[o] = s22[o] = s12[o] = 0.f; }
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) {                      // input j feeds output o with tap k = j - o
            const float aa = a[j] * a[j], bb = b[j] * b[j], ab = a[j] * b[j];
#pragma unroll
            for (int o = 0; o < SS_Q; o++) {
                const int k = j - o;
                if (k < 0 || k > 2 * SS_R) continue;
                const float w = win.w[k];
                m1[o] = fmaf(w, a[j], m1[o]); m2[o] = fmaf(w, b[j], m2[o]);
                s11[o] = fmaf(w, aa, s11[o]); s22[o] = fmaf(w, bb, s22[o]); s12[o] = fmaf(w, ab, s12[o]);
            }
        }
        *reinterpret_cast<float4*>(&s_h[0][ly][x]) = make_float4(m1[0], m1[1], m1[2], m1[3]);
        *reinterpret_cast<float4*>(&s_h[1][ly][x]) = make_float4(m2[0], m2[1], m2[2], m2[3]);
        *reinterpret_cast<float4*>(&s_h[2][ly][x]) = make_float4(s11[0], s11[1], s11[2], s11[3]);
        *reinterpret_cast<float4*>(&s_h[3][ly][x]) = make_float4(s22[0], s22[1], s22[2], s22[3]);
        *reinterpret_cast<float4*>(&s_h[4][ly][x]) = make_float4(s12[0], s12[1], s12[2], s12[3]);
    }
    __syncthreads();
    // vertical pass: thread = column tx, output rows 4 ty .. 4 ty + 3
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float r[5][SS_Q];
#pragma unroll
    for (int m = 0; m < 5; m++) {
        float col[SS_Q + 2 * SS_R];
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) col[j] = s_h[m][SS_Q * ty + j][tx];
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) acc = fmaf(win.w[k], col[o + k], acc);
            r[m][o] = acc;
        }
    }
    const int gx = x0 + tx;
    float ssim_v = 0.f, l1_v = 0.f;
#pragma unroll
    for (int o = 0; o < SS_Q; o++) {
        const int gy = y0 + SS_Q * ty + o;
        if (gx >= W || gy >= H) continue;
        const float mu1 = r[0][o], mu2 = r[1][o], e11 = r[2][o], e22 = r[3][o], e12 = r[4][o];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float sig1 = e11 - mu1_sq, sig2 = e22 - mu2_sq, sig12 = e12 - mu12;
        const float A = mu1_sq + mu2_sq + SS_C1, B = sig1 + sig2 + SS_C2;
        const float Cn = 2.f * mu12 + SS_C1, D = 2.f * sig12 + SS_C2;
        const float inv_AB = 1.f / (A * B);
        const float mval = Cn * D * inv_AB;
        ssim_v += mval;
        // partial derivatives of the map w.r.t. (mu1, E[x^2], E[xy]) at this pixel (mu2, E[y^2] fixed)
        const float d_mu1 = 2.f * mu2 * (D - Cn) * inv_AB - mval * 2.f * mu1 * (B - A) * inv_AB;
        const float d_s11 = -mval / B;
        const float d_s12 = 2.f * Cn * inv_AB;
        const size_t ofs = c * plane + (size_t)gy * W + gx;
        const size_t stride = (size_t)gridDim.z * plane;
        maps[ofs] = d_mu1; maps[stride + ofs] = d_s11; maps[2 * stride + ofs] = d_s12;
        l1_v += fabsf(s_a[SS_Q * ty + o + SS_R][tx + SS_R] - s_b[SS_Q * ty + o + SS_R][tx + SS_R]);
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

__global__ void __launch_bounds__(SS_THREADS)
photometric_bwd_kernel(int H, int W, Window win, float inv_n, float lambda, const float* __restrict__ img,
                       const float* __restrict__ gt, const float* __restrict__ maps,
                       const float* __restrict__ dL_dloss, float* __restrict__ dL_dimg)
{
    __shared__ __align__(16) float s_m[3][SS_P][SS_PITCH];
    __shared__ __align__(16) float s_h[3][SS_P][SS_T];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W, stride = (size_t)gridDim.z * plane;
#pragma unroll
    for (int k = 0; k < 3; k++) stage_plane(maps + k * stride + c * plane, H, W, x0, y0, s_m[k]);
    __syncthreads();
    for (int it = threadIdx.x; it < SS_P * (SS_T / SS_Q); it += SS_THREADS) {
        const int ly = it / (SS_T / SS_Q), x = (it - ly * (SS_T / SS_Q)) * SS_Q;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float v[16], acc[SS_Q] = {0.f, 0.f, 0.f, 0.f};
            load16(&s_m[m][ly][x], v);
#pragma unroll
            for (int j = 0; j < SS_Q + 2 * SS_R; j++)
#pragma unroll
                for (int o = 0; o < SS_Q; o++) {
                    const int k = j - o;
                    if (k < 0 || k > 2 * SS_R) continue;
                    acc[o] = fmaf(win.w[k], v[j], acc[o]);
                }
            *reinterpret_cast<float4*>(&s_h[m][ly][x]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float r[3][SS_Q];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        float col[SS_Q + 2 * SS_R];
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) col[j] = s_h[m][SS_Q * ty + j][tx];
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) acc = fmaf(win.w[k], col[o + k], acc);
            r[m][o] = acc;
        }
    }
    const int gx = x0 + tx;
    const float g = dL_dloss ? __ldg(dL_dloss) : 1.f;
#pragma unroll
    for (int o = 0; o < SS_Q; o++) {
        const int gy = y0 + SS_Q * ty + o;
        if (gx >= W || gy >= H) continue;
        const size_t ofs = c * plane + (size_t)gy * W + gx;
        const float a = __ldg(img + ofs), b = __ldg(gt + ofs);
        const float d_ssim = r[0][o] + 2.f * a * r[1][o] + b * r[2][o];
        const float diff = a - b;
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);   // torch.abs backward: sign(0) = 0
        dL_dimg[ofs] = g * inv_n * ((1.f - lambda) * sgn - lambda * d_ssim);
    }
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
    photometric_fwd_kernel<<<grid, SS_THREADS, 0, s>>>(H, W, win, image, gt, maps, sums2);
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
    photometric_bwd_kernel<<<grid, SS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        H, W, win, 1.f / ((float)C * (float)H * (float)W), lambda_dssim, image, gt, maps, dL_dloss, dL_dimage);
    EOGS_LAUNCH_CHECK("photometric_bwd_kernel");
    return 0;
}

}  // extern "C"
