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
// registers (the horizontal pass reads its 14 + 2 inputs as LDS.128, the vertical pass reads 14 values per map for 4
// outputs), and PACKED: the two images travel as float2 pairs, so (mu1, mu2) and (E[x^2], E[y^2]) are one FFMA2 per tap
// each (f32x2.cuh) — 3 FMA instructions per tap instead of 5, 15 shared-memory instructions per output pixel instead
// of 77; the halo amplification of the staged tile drops from 2.6x (16x16) to 1.7x (32x32).
// HBM: forward reads 8 B and writes 12 B per pixel-channel, backward reads 20 B and writes 4 B.
#include "common.cuh"
#include "f32x2.cuh"

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

// Stage the 42x42 neighbourhood of TWO planes interleaved as float2 {p, q} (zero padding outside the image, conv2d
// padding=5) into rows of SS_PITCH pairs; the two pad columns are zero.  A pair is the operand of one packed FFMA2.
__device__ __forceinline__ void stage_pair(const float* __restrict__ p, const float* __restrict__ q, int H, int W,
                                           int x0, int y0, float2 (*dst)[SS_PITCH]) {
    for (int i = threadIdx.x; i < SS_P * SS_PITCH; i += SS_THREADS) {
        const int ly = i / SS_PITCH, lx = i - ly * SS_PITCH;
        const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
        const bool in = lx < SS_P && gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = (size_t)gy * W + gx;
        dst[ly][lx] = in ? make_float2(__ldg(p + o), __ldg(q + o)) : make_float2(0.f, 0.f);
    }
}
__device__ __forceinline__ void stage_plane(const float* __restrict__ src, int H, int W, int x0, int y0,
                                            float (*dst)[SS_PITCH]) {
    for (int i = threadIdx.x; i < SS_P * SS_PITCH; i += SS_THREADS) {
        const int ly = i / SS_PITCH, lx = i - ly * SS_PITCH;
        const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
        const bool in = lx < SS_P && gy >= 0 && gy < H && gx >= 0 && gx < W;
        dst[ly][lx] = in ? __ldg(src + (size_t)gy * W + gx) : 0.f;
    }
}

// 16 consecutive pairs / floats of a staged row, starting at a multiple of 4: conflict-free LDS.128
__device__ __forceinline__ void load16(const float2* row, f2* v) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 q = *reinterpret_cast<const float4*>(row + 2 * k);
        v[2 * k] = mk2(q.x, q.y); v[2 * k + 1] = mk2(q.z, q.w);
    }
}
__device__ __forceinline__ void load16(const float* row, float* v) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 q = *reinterpret_cast<const float4*>(row + 4 * k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}
__device__ __forceinline__ void store4(float2* dst, const f2* v) {
    *reinterpret_cast<float4*>(dst) = make_float4(lo2(v[0]), hi2(v[0]), lo2(v[1]), hi2(v[1]));
    *reinterpret_cast<float4*>(dst + 2) = make_float4(lo2(v[2]), hi2(v[2]), lo2(v[3]), hi2(v[3]));
}

__global__ void __launch_bounds__(SS_THREADS)
photometric_fwd_kernel(int H, int W, Window win, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ maps, float* __restrict__ sums)
{
    // the two images travel as pairs {img, gt}: (mu1, mu2) and (E[x^2], E[y^2]) are each ONE packed accumulator
    __shared__ __align__(16) float2 s_ab[SS_P][SS_PITCH];
    __shared__ __align__(16) float2 s_hm[SS_P][SS_T], s_hs[SS_P][SS_T];      // horizontal results {m1, m2}, {s11, s22}
    __shared__ __align__(16) float s_hx[SS_P][SS_T];                          // and s12
    __shared__ float s_red[8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W;
    stage_pair(img + c * plane, gt + c * plane, H, W, x0, y0, s_ab);
    __syncthreads();
    // horizontal pass: 42 rows x 8 items of 4 adjacent outputs
    for (int it = threadIdx.x; it < SS_P * (SS_T / SS_Q); it += SS_THREADS) {
        const int ly = it / (SS_T / SS_Q), x = (it - ly * (SS_T / SS_Q)) * SS_Q;
        f2 ab[16];
        load16(&s_ab[ly][x], ab);
        f2 m[SS_Q], sq[SS_Q];
        float s12[SS_Q];
#pragma unroll
        for (int o = 0; o < SS_Q; o++) { m[o] = bc2(0.f); sq[o] = bc2(0.f); s12[o] = 0.f; }
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) {                      // input j feeds output o with tap k = j - o
            const f2 p2 = mul2(ab[j], ab[j]);
            const float pq = lo2(ab[j]) * hi2(ab[j]);
#pragma unroll
            for (int o = 0; o < SS_Q; o++) {
                const int k = j - o;
                if (k < 0 || k > 2 * SS_R) continue;
                const float w = win.w[k];
                fma2_acc(m[o], bc2(w), ab[j]);
                fma2_acc(sq[o], bc2(w), p2);
                s12[o] = fmaf(w, pq, s12[o]);
            }
        }
        store4(&s_hm[ly][x], m);
        store4(&s_hs[ly][x], sq);
        *reinterpret_cast<float4*>(&s_hx[ly][x]) = make_float4(s12[0], s12[1], s12[2], s12[3]);
    }
    __syncthreads();
    // vertical pass: thread = column tx, output rows 4 ty .. 4 ty + 3
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    f2 rm[SS_Q], rs[SS_Q];
    float rx[SS_Q];
    {
        f2 col[SS_Q + 2 * SS_R];
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) { const float2 q = s_hm[SS_Q * ty + j][tx]; col[j] = mk2(q.x, q.y); }
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            f2 acc = bc2(0.f);
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) fma2_acc(acc, bc2(win.w[k]), col[o + k]);
            rm[o] = acc;
        }
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) { const float2 q = s_hs[SS_Q * ty + j][tx]; col[j] = mk2(q.x, q.y); }
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            f2 acc = bc2(0.f);
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) fma2_acc(acc, bc2(win.w[k]), col[o + k]);
            rs[o] = acc;
        }
        float cx[SS_Q + 2 * SS_R];
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) cx[j] = s_hx[SS_Q * ty + j][tx];
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) acc = fmaf(win.w[k], cx[o + k], acc);
            rx[o] = acc;
        }
    }
    const int gx = x0 + tx;
    float ssim_v = 0.f, l1_v = 0.f;
#pragma unroll
    for (int o = 0; o < SS_Q; o++) {
        const int gy = y0 + SS_Q * ty + o;
        if (gx >= W || gy >= H) continue;
        const float mu1 = lo2(rm[o]), mu2 = hi2(rm[o]), e11 = lo2(rs[o]), e22 = hi2(rs[o]), e12 = rx[o];
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
        const float2 px = s_ab[SS_Q * ty + o + SS_R][tx + SS_R];
        l1_v += fabsf(px.x - px.y);
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
    // maps 0 and 1 (d_mu1, d_s11) travel as pairs, map 2 (d_s12) alone
    __shared__ __align__(16) float2 s_m01[SS_P][SS_PITCH];
    __shared__ __align__(16) float s_m2[SS_P][SS_PITCH];
    __shared__ __align__(16) float2 s_h01[SS_P][SS_T];
    __shared__ __align__(16) float s_h2[SS_P][SS_T];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)H * W, stride = (size_t)gridDim.z * plane;
    stage_pair(maps + c * plane, maps + stride + c * plane, H, W, x0, y0, s_m01);
    stage_plane(maps + 2 * stride + c * plane, H, W, x0, y0, s_m2);
    __syncthreads();
    for (int it = threadIdx.x; it < SS_P * (SS_T / SS_Q); it += SS_THREADS) {
        const int ly = it / (SS_T / SS_Q), x = (it - ly * (SS_T / SS_Q)) * SS_Q;
        f2 v01[16], a01[SS_Q];
        float v2[16], a2[SS_Q];
        load16(&s_m01[ly][x], v01);
        load16(&s_m2[ly][x], v2);
#pragma unroll
        for (int o = 0; o < SS_Q; o++) { a01[o] = bc2(0.f); a2[o] = 0.f; }
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++)
#pragma unroll
            for (int o = 0; o < SS_Q; o++) {
                const int k = j - o;
                if (k < 0 || k > 2 * SS_R) continue;
                fma2_acc(a01[o], bc2(win.w[k]), v01[j]);
                a2[o] = fmaf(win.w[k], v2[j], a2[o]);
            }
        store4(&s_h01[ly][x], a01);
        *reinterpret_cast<float4*>(&s_h2[ly][x]) = make_float4(a2[0], a2[1], a2[2], a2[3]);
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    f2 r01[SS_Q];
    float r2[SS_Q];
    {
        f2 col[SS_Q + 2 * SS_R];
        float cx[SS_Q + 2 * SS_R];
#pragma unroll
        for (int j = 0; j < SS_Q + 2 * SS_R; j++) {
            const float2 q = s_h01[SS_Q * ty + j][tx];
            col[j] = mk2(q.x, q.y);
            cx[j] = s_h2[SS_Q * ty + j][tx];
        }
#pragma unroll
        for (int o = 0; o < SS_Q; o++) {
            f2 acc = bc2(0.f);
            float ax = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * SS_R; k++) { fma2_acc(acc, bc2(win.w[k]), col[o + k]); ax = fmaf(win.w[k], cx[o + k], ax); }
            r01[o] = acc; r2[o] = ax;
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
        const float d_ssim = lo2(r01[o]) + 2.f * a * hi2(r01[o]) + b * r2[o];
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
