// Backward of the geometry stage, one thread per Gaussian, fused into ONE kernel:
//   computeCov2DCUDA          (DGR/cuda_rasterizer/backward.cu:147-327)
//   preprocessCUDA<5> backward (backward.cu:400-454) with computeCov3D backward (:331-394)
//   the three torch reductions of _RasterizeGaussians.backward that build grad_viewmatrix
//   (DGR/diff_gaussian_rasterization/__init__.py:174-202), as in-kernel block reductions:
//     cam_sums[0..5]   = sum_p dL_dT[p][0..5]
//     cam_sums[6..11]  = means3D^T @ dL_dmeans2D[:, :2]   (3x2 row-major)
//     cam_sums[12..13] = sum_p dL_dmeans2D[p][0..1]
// so no P x 6 dL_dT tensor, no P x 2 x 2 dL_dconic tensor and no dL_dcov3D tensor (unless the
// caller passed precomputed covariances) ever reach HBM.
//
// dL_dT uses the intended per-Gaussian stride (6*idx + k).  The reference writes dL_dT[idx + k]
// (backward.cu:320-325), a data race that leaves race-dependent garbage in that term; see
// DESIGN.md "carve-outs".
//
// HBM-bound streaming kernel: 64 B gradient record + 44 B parameters in, 4*(3+C+1+3+3+4) B out.
#include "common.cuh"
#include "geom_math.cuh"

namespace eogs {

constexpr int PBW_THREADS = 256;
constexpr int NSUMS = 14;

// RAW = backward of the fused render path: the inputs are the raw parameters (log-scales,
// un-normalised quaternions, opacity logits) and the gradients are chained through the activations and
// through colors_precomp = [SH2RGB(f_dc), altitude(xyz), 1] in this kernel:
//   dL_dcolors    -> dL_df_dc [P,3]  = C0 * dL_drgb            (utils/sh_utils.py:125-126)
//   dL_dmeans3D  += a * dL_daltitude                           (ECEF_to_UVA, affine_cameras.py:432-438)
//   alt_sums[4]   = sum_p dL_daltitude * (x, y, z, 1)          (gradient of the altitude affine row)
//   dL_dopacity   -> dL_dlogit      = dL_dop * op * (1 - op)
//   dL_dscales    -> dL_dlog_scale  = dL_ds * s
//   dL_drotations -> dL_draw_quat   = (dL_dq - q (q . dL_dq)) / max(|r|, 1e-12)
template <int C, bool RAW>
__global__ void __launch_bounds__(PBW_THREADS, RAW ? 2 : 3)
preprocess_bwd_kernel(int P, int W, int H, const float* __restrict__ alt_affine, float* __restrict__ alt_sums,
                      const float* __restrict__ means3D, const float* __restrict__ scales,
                      const float4* __restrict__ rotations, const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ opacities, const float* __restrict__ view,
                      const float* __restrict__ proj, float scale_modifier, bool antialiasing,
                      const int32_t* __restrict__ radii, const float4* __restrict__ grad_rec,
                      float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dcolors,
                      float* __restrict__ dL_dopacity, float* __restrict__ dL_dmeans3D,
                      float* __restrict__ dL_dcov3D, float* __restrict__ dL_dscales,
                      float4* __restrict__ dL_drotations, float* __restrict__ cam_sums)
{
    __shared__ float s_view[16], s_proj[16];
    constexpr int NS = NSUMS + (RAW ? 4 : 0);
    __shared__ float s_part[PBW_THREADS / 32][NS];
    if (threadIdx.x < 16) {
        s_view[threadIdx.x] = __ldg(view + threadIdx.x);
        s_proj[threadIdx.x] = __ldg(proj + threadIdx.x);
    }
    __syncthreads();

    const int idx = blockIdx.x * PBW_THREADS + threadIdx.x;
    float sums[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) sums[k] = 0.f;

    if (idx < P) {
        float g_mean2D[2] = {0.f, 0.f}, g_col[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, g_op = 0.f;
        float g_mean3D[3] = {0.f, 0.f, 0.f}, g_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float g_scale[3] = {0.f, 0.f, 0.f};
        float4 g_rot = make_float4(0.f, 0.f, 0.f, 0.f);

        if (__ldg(radii + idx) > 0) {
            const float4 ga = __ldg(grad_rec + (size_t)idx * (GRAD_STRIDE / 4));
            const float4 gb = __ldg(grad_rec + (size_t)idx * (GRAD_STRIDE / 4) + 1);
            const float4 gc = __ldg(grad_rec + (size_t)idx * (GRAD_STRIDE / 4) + 2);
            // The blend backward accumulated plain sums over the Gaussian's pixels (u = G dL/dalpha):
            //   ga.x, ga.y = sum u dx, sum u dy;   ga.z, ga.w, gb.x = sum u (dx dx, dx dy, dy dy);   gb.y = sum u
            // The per-Gaussian constants are applied below, once the conic and the opacity are recomputed:
            //   dL_dconic = -1/2 opacity * (ga.z, ga.w, gb.x)                                (backward.cu:634-640)
            //   dL_dmean2D = -(W/2, H/2) * opacity * conic . (sum u dx, sum u dy)            (backward.cu:631-632)
            const float mom_x = ga.x, mom_y = ga.y;
            float dcon_x = ga.z, dcon_y = ga.w, dcon_w = gb.x;
            g_op = gb.y;
            g_col[0] = gb.z; g_col[1] = gb.w; g_col[2] = gc.x; g_col[3] = gc.y; g_col[4] = gc.z;

            const float mx = __ldg(means3D + 3 * (size_t)idx), my = __ldg(means3D + 3 * (size_t)idx + 1),
                        mz = __ldg(means3D + 3 * (size_t)idx + 2);

            // ---- recompute Sigma3D, T, Sigma2D exactly as the forward did ----
            float c3[6];
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            float sx = 0.f, sy = 0.f, sz = 0.f, raw_norm = 1.f;
            float act_s[3] = {0.f, 0.f, 0.f};
            float op_act = 0.f;
            if (RAW) op_act = act_sigmoid(__ldg(opacities + idx));
            Rot3 R;
            Mat3 M;
            if (cov3D_precomp) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = __ldg(cov3D_precomp + 6 * (size_t)idx + k);
            } else {
                q = __ldg(rotations + idx);
                sx = __ldg(scales + 3 * (size_t)idx);
                sy = __ldg(scales + 3 * (size_t)idx + 1);
                sz = __ldg(scales + 3 * (size_t)idx + 2);
                if (RAW) {
                    sx = expf(sx); sy = expf(sy); sz = expf(sz);
                    act_s[0] = sx; act_s[1] = sy; act_s[2] = sz;
                    raw_norm = quat_norm_clamped(q);
                    q = act_normalize(q, raw_norm);
                }
                sx = __fmul_rn(scale_modifier, sx);
                sy = __fmul_rn(scale_modifier, sy);
                sz = __fmul_rn(scale_modifier, sz);
                R = quat_to_R(q.x, q.y, q.z, q.w);
                M = scale_rot(sx, sy, sz, R);
                cov3d_from_M(M, c3);
            }
            const Affine2x3 T = make_T(s_view, W, H);
            float c_xx, c_xy, c_yy;
            cov2d_from_cov3d(T, c3, c_xx, c_xy, c_yy);

            // ---- computeCov2DCUDA (backward.cu:198-251) ----
            constexpr float h_var = 0.3f;
            float dL_dc_xx = 0.f, dL_dc_xy = 0.f, dL_dc_yy = 0.f;
            float aa_scale = 1.f;                          // the forward's opacity factor (forward.cu:229-235)
            if (antialiasing) {
                const float det_cov = c_xx * c_yy - c_xy * c_xy;
                c_xx += h_var; c_yy += h_var;
                const float det_plus = c_xx * c_yy - c_xy * c_xy;
                const float ratio = det_cov / det_plus;
                const float h_scale = sqrtf(fmaxf(0.000025f, ratio));
                aa_scale = h_scale;
                const float d_h = g_op * (RAW ? op_act : __ldg(opacities + idx));
                g_op = g_op * h_scale;
                const float d_inside_root = ratio <= 0.000025f ? 0.f : d_h / (2.f * h_scale);
                const float x = c_xx, y = c_yy, z = c_xy, w = h_var;
                const float den = w * w + w * (x + y) + x * y - z * z;
                const float denom_f = d_inside_root / (den * den);
                dL_dc_xx = w * (w * y + y * y + z * z) * denom_f;
                dL_dc_yy = w * (w * x + x * x + z * z) * denom_f;
                dL_dc_xy = -2.f * w * z * (w + x + y) * denom_f;
            } else {
                c_xx += h_var; c_yy += h_var;
            }
            const float denom = c_xx * c_yy - c_xy * c_xy;
            {
                const float op_fwd = (RAW ? op_act : __ldg(opacities + idx)) * aa_scale;      // opacity as the forward stored it
                const float det_inv = __fdiv_rn(1.f, denom);                 // conic = (c_yy, -c_xy, c_xx) / det (forward.cu:239-241)
                const float con_x = c_yy * det_inv, con_y = -c_xy * det_inv, con_z = c_xx * det_inv;
                g_mean2D[0] = -0.5f * (float)W * op_fwd * (con_x * mom_x + con_y * mom_y);
                g_mean2D[1] = -0.5f * (float)H * op_fwd * (con_z * mom_y + con_y * mom_x);
                const float hop = -0.5f * op_fwd;
                dcon_x *= hop; dcon_y *= hop; dcon_w *= hop;
            }
            const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
            if (denom2inv != 0.f) {
                dL_dc_xx += denom2inv * (-c_yy * c_yy * dcon_x + 2.f * c_xy * c_yy * dcon_y + (denom - c_xx * c_yy) * dcon_w);
                dL_dc_yy += denom2inv * (-c_xx * c_xx * dcon_w + 2.f * c_xx * c_xy * dcon_y + (denom - c_xx * c_yy) * dcon_x);
                dL_dc_xy += denom2inv * 2.f * (c_xy * c_yy * dcon_x - (denom + 2.f * c_xy * c_xy) * dcon_y + c_xx * c_xy * dcon_w);
                g_cov[0] = T.t00 * T.t00 * dL_dc_xx + T.t00 * T.t10 * dL_dc_xy + T.t10 * T.t10 * dL_dc_yy;
                g_cov[3] = T.t01 * T.t01 * dL_dc_xx + T.t01 * T.t11 * dL_dc_xy + T.t11 * T.t11 * dL_dc_yy;
                g_cov[5] = T.t02 * T.t02 * dL_dc_xx + T.t02 * T.t12 * dL_dc_xy + T.t12 * T.t12 * dL_dc_yy;
                g_cov[1] = 2.f * T.t00 * T.t01 * dL_dc_xx + (T.t00 * T.t11 + T.t01 * T.t10) * dL_dc_xy + 2.f * T.t10 * T.t11 * dL_dc_yy;
                g_cov[2] = 2.f * T.t00 * T.t02 * dL_dc_xx + (T.t00 * T.t12 + T.t02 * T.t10) * dL_dc_xy + 2.f * T.t10 * T.t12 * dL_dc_yy;
                g_cov[4] = 2.f * T.t02 * T.t01 * dL_dc_xx + (T.t01 * T.t12 + T.t02 * T.t11) * dL_dc_xy + 2.f * T.t11 * T.t12 * dL_dc_yy;
            }

            // ---- dL_dT, upper 2x3 of T (backward.cu:270-281), reduced over Gaussians ----
            const float tv0 = T.t00 * c3[0] + T.t01 * c3[1] + T.t02 * c3[2];   // row 0 of T times Vrk columns
            const float tv1 = T.t00 * c3[1] + T.t01 * c3[3] + T.t02 * c3[4];
            const float tv2 = T.t00 * c3[2] + T.t01 * c3[4] + T.t02 * c3[5];
            const float uv0 = T.t10 * c3[0] + T.t11 * c3[1] + T.t12 * c3[2];   // row 1
            const float uv1 = T.t10 * c3[1] + T.t11 * c3[3] + T.t12 * c3[4];
            const float uv2 = T.t10 * c3[2] + T.t11 * c3[4] + T.t12 * c3[5];
            sums[0] = 2.f * tv0 * dL_dc_xx + uv0 * dL_dc_xy;
            sums[1] = 2.f * tv1 * dL_dc_xx + uv1 * dL_dc_xy;
            sums[2] = 2.f * tv2 * dL_dc_xx + uv2 * dL_dc_xy;
            sums[3] = 2.f * uv0 * dL_dc_yy + tv0 * dL_dc_xy;
            sums[4] = 2.f * uv1 * dL_dc_yy + tv1 * dL_dc_xy;
            sums[5] = 2.f * uv2 * dL_dc_yy + tv2 * dL_dc_xy;

            // ---- mean: dL_dmean3D = A_2x3^T dL_dmean2D, with projmatrix (backward.cu:439-445) ----
            const float gx = g_mean2D[0], gy = g_mean2D[1];
            g_mean3D[0] = s_proj[0] * gx + s_proj[1] * gy;
            g_mean3D[1] = s_proj[4] * gx + s_proj[5] * gy;
            g_mean3D[2] = s_proj[8] * gx + s_proj[9] * gy;
            sums[6] = mx * gx;  sums[7] = mx * gy;
            sums[8] = my * gx;  sums[9] = my * gy;
            sums[10] = mz * gx; sums[11] = mz * gy;
            sums[12] = gx;      sums[13] = gy;

            // ---- computeCov3D backward (backward.cu:331-394) ----
            if (!cov3D_precomp) {
                const float dS[3][3] = {{g_cov[0], 0.5f * g_cov[1], 0.5f * g_cov[2]},
                                        {0.5f * g_cov[1], g_cov[3], 0.5f * g_cov[4]},
                                        {0.5f * g_cov[2], 0.5f * g_cov[4], g_cov[5]}};
                const float Mm[3][3] = {{M.m00, M.m01, M.m02}, {M.m10, M.m11, M.m12}, {M.m20, M.m21, M.m22}};   // [c][k]
                const float Rm[3][3] = {{R.r00, R.r01, R.r02}, {R.r10, R.r11, R.r12}, {R.r20, R.r21, R.r22}};
                const float s[3] = {sx, sy, sz};
                // dL_dM[c][k] = 2 * sum_m M[m][k] * dS[c][m];   D[k][c] = s_k * dL_dM[c][k] = dL/dR[c][k]
                float D[3][3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float gs = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float dM = 2.f * (Mm[0][k] * dS[c][0] + Mm[1][k] * dS[c][1] + Mm[2][k] * dS[c][2]);
                        gs += Rm[c][k] * dM;
                        D[k][c] = s[k] * dM;
                    }
                    g_scale[k] = gs;
                }
                const float r = q.x, x = q.y, y = q.z, z = q.w;
                g_rot.x = 2.f * z * (D[0][1] - D[1][0]) + 2.f * y * (D[2][0] - D[0][2]) + 2.f * x * (D[1][2] - D[2][1]);
                g_rot.y = 2.f * y * (D[1][0] + D[0][1]) + 2.f * z * (D[2][0] + D[0][2]) + 2.f * r * (D[1][2] - D[2][1]) - 4.f * x * (D[2][2] + D[1][1]);
                g_rot.z = 2.f * x * (D[1][0] + D[0][1]) + 2.f * r * (D[2][0] - D[0][2]) + 2.f * z * (D[1][2] + D[2][1]) - 4.f * y * (D[2][2] + D[0][0]);
                g_rot.w = 2.f * r * (D[0][1] - D[1][0]) + 2.f * x * (D[2][0] + D[0][2]) + 2.f * y * (D[1][2] + D[2][1]) - 4.f * z * (D[1][1] + D[0][0]);
            }
            if (RAW) {
                // chain through the activations exactly as autograd does behind the unfused call: the kernel's
                // dL_dscales (which, like the reference's, is taken w.r.t. mod * s) times d exp(x)/dx = s
                g_scale[0] *= act_s[0]; g_scale[1] *= act_s[1]; g_scale[2] *= act_s[2];
                const float qd = q.x * g_rot.x + q.y * g_rot.y + q.z * g_rot.z + q.w * g_rot.w;
                const float inv_n = 1.f / raw_norm;
                g_rot = make_float4((g_rot.x - q.x * qd) * inv_n, (g_rot.y - q.y * qd) * inv_n,
                                    (g_rot.z - q.z * qd) * inv_n, (g_rot.w - q.w * qd) * inv_n);
                g_op = g_op * op_act * (1.f - op_act);
                const float g_alt = g_col[3];
                g_mean3D[0] += __ldg(alt_affine) * g_alt;
                g_mean3D[1] += __ldg(alt_affine + 1) * g_alt;
                g_mean3D[2] += __ldg(alt_affine + 2) * g_alt;
                sums[NSUMS] = mx * g_alt; sums[NSUMS + 1] = my * g_alt; sums[NSUMS + 2] = mz * g_alt; sums[NSUMS + 3] = g_alt;
                g_col[0] *= SH_C0; g_col[1] *= SH_C0; g_col[2] *= SH_C0;
            }
        }

        dL_dmeans2D[3 * (size_t)idx] = g_mean2D[0];
        dL_dmeans2D[3 * (size_t)idx + 1] = g_mean2D[1];
        dL_dmeans2D[3 * (size_t)idx + 2] = 0.f;
        constexpr int COUT = RAW ? 3 : C;            // RAW: dL_df_dc [P,3]
#pragma unroll
        for (int ch = 0; ch < COUT; ch++) dL_dcolors[COUT * (size_t)idx + ch] = g_col[ch];
        dL_dopacity[idx] = g_op;
#pragma unroll
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * (size_t)idx + k] = g_mean3D[k];
        if (dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)idx + k] = g_cov[k];
        }
        if (dL_dscales) {
#pragma unroll
            for (int k = 0; k < 3; k++) dL_dscales[3 * (size_t)idx + k] = g_scale[k];
        }
        if (dL_drotations) dL_drotations[idx] = g_rot;
    }

    // ---- block reduction of the 14 camera sums: shuffle tree, then one atomic per block ----
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; k++) {
        float v = sums[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < PBW_THREADS / 32; w++) v += s_part[w][threadIdx.x];
        if (v != 0.f) atomicAdd(threadIdx.x < NSUMS ? cam_sums + threadIdx.x : alt_sums + (threadIdx.x - NSUMS), v);
    }
}

int launch_preprocess_bwd(cudaStream_t s, int P, int W, int H, int channels, bool raw_params,
                          const float* alt_affine, float* alt_sums,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities,
                          const float* view, const float* proj, float scale_modifier,
                          bool antialiasing, const int32_t* radii, const float* grad_rec,
                          float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                          float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                          float* dL_drotations, float* cam_sums)
{
    // quaternions travel as float4 (one LDG.128 in, one STG.128 out): a view with a storage offset of an odd number
    // of floats would fault with "misaligned address" and poison the context — refuse it here instead
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if ((rotations && !a16(rotations)) || (dL_drotations && !a16(dL_drotations))) {
        set_error("rotations and dL_drotations must be 16-byte aligned (got %p, %p)", (const void*)rotations, (void*)dL_drotations);
        return -2;
    }
    EOGS_CUDA(cudaMemsetAsync(cam_sums, 0, 16 * sizeof(float), s));
    if (raw_params) {
        if (channels != 5 || cov3D_precomp || !alt_affine || !alt_sums) { set_error("the fused-parameter path renders 5 channels from scales+rotations"); return -1; }
        EOGS_CUDA(cudaMemsetAsync(alt_sums, 0, 4 * sizeof(float), s));
    }
    const int blocks = (P + PBW_THREADS - 1) / PBW_THREADS;
    auto run = [&](auto kernel) {
        kernel<<<blocks, PBW_THREADS, 0, s>>>(
            P, W, H, alt_affine, alt_sums, means3D, scales, reinterpret_cast<const float4*>(rotations), cov3D_precomp,
            opacities, view, proj, scale_modifier, antialiasing, radii,
            reinterpret_cast<const float4*>(grad_rec), dL_dmeans2D, dL_dcolors, dL_dopacity,
            dL_dmeans3D, dL_dcov3D, dL_dscales, reinterpret_cast<float4*>(dL_drotations), cam_sums);
    };
    if (raw_params) run(preprocess_bwd_kernel<5, true>);
    else if (channels == 5) run(preprocess_bwd_kernel<5, false>);
    else if (channels == 3) run(preprocess_bwd_kernel<3, false>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    EOGS_LAUNCH_CHECK("preprocess_bwd_kernel");
    return 0;
}

}  // namespace eogs
