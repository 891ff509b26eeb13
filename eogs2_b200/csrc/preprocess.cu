// Forward geometry stage: one thread per Gaussian.
// Replaces preprocessCUDA<5> (DGR/cuda_rasterizer/forward.cu:154-283).
//
// HBM-bound streaming kernel.  Algorithmic bytes per Gaussian: 44 B of parameters
// (means 12, scales 12, rotations 16, opacity 4) + 4*C B of colours in; 48 B splat record
// + 4 depth + 8 rect + 4 tiles + 4 radius + 8 sort key/id out  = 140 B at C = 5.
// AoS float3 / float[C] inputs are staged through shared memory with 16-byte coalesced
// loads; each thread then reads its own row (stride 3 or 5 floats: conflict-free).
#include "common.cuh"
#include "geom_math.cuh"

namespace eogs {

constexpr int PRE_THREADS = 256;

// Cooperative load of rows [base, base+PRE_THREADS) x K floats into smem, float4-vectorised.
template <int K>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, float* __restrict__ dst,
                                           int base, int P, bool aligned16) {
    const int rows = min(PRE_THREADS, P - base);
    const int nfloat = rows * K;
    const float* s = src + (size_t)base * K;       // base*K*4 bytes: multiple of 16 (base % 256 == 0)
    const int nvec = aligned16 ? (nfloat >> 2) : 0; // a torch view with a storage offset may be only 4-byte aligned
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < nvec; i += PRE_THREADS) d4[i] = __ldg(s4 + i);
    for (int i = (nvec << 2) + threadIdx.x; i < nfloat; i += PRE_THREADS) dst[i] = __ldg(s + i);
}

// RAW = the fused render path (eogs_forward_geometry_params): `scales` holds log-scales, `rotations`
// un-normalised quaternions, `opacities` logits and `colors` the SH DC coefficients [P,3]; the
// activations and colors_precomp = [SH2RGB(f_dc), altitude, 1] (renderer.py:84-107) are computed here
// instead of by ~10 torch kernels and 5 P-sized temporaries.  alt_affine[4]: altitude = a . xyz + b
// (AffineCamera.ECEF_to_UVA, scene/cameras/affine_cameras.py:432-438, third component).
template <int C, bool RAW>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_fwd_kernel(int P, int W, int H, int grid_x, int grid_y, int band_y0, int band_y1,
                      const float* __restrict__ means3D, const float* __restrict__ scales,
                      const float4* __restrict__ rotations, const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ opacities, const float* __restrict__ colors,
                      const float* __restrict__ view, const float* __restrict__ alt_affine,
                      float scale_modifier, bool antialiasing,
                      int align_mask, int32_t* __restrict__ radii, float4* __restrict__ splat,
                      float* __restrict__ alpha_cut, float* __restrict__ depth, uint2* __restrict__ rect,
                      uint32_t* __restrict__ tiles, uint32_t* __restrict__ key_in,
                      uint32_t* __restrict__ id_in, eogs_forward_info* __restrict__ info,
                      volatile eogs_forward_info* host_info)
{
    __shared__ __align__(16) float s_mean[PRE_THREADS * 3];
    __shared__ __align__(16) float s_scale[PRE_THREADS * 3];
    constexpr int CIN = RAW ? 3 : C;                 // floats per Gaussian in `colors`
    __shared__ __align__(16) float s_color[PRE_THREADS * CIN];
    __shared__ float s_view[16];

    const int base = blockIdx.x * PRE_THREADS;
    stage_rows<3>(means3D, s_mean, base, P, align_mask & 1);
    if (scales) stage_rows<3>(scales, s_scale, base, P, align_mask & 2);
    stage_rows<CIN>(colors, s_color, base, P, align_mask & 4);
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(view + threadIdx.x);
    __syncthreads();

    const int idx = base + threadIdx.x;
    if (idx >= P) return;

    // Defaults for a culled Gaussian (reference: radii = tiles_touched = 0, forward.cu:189-190).
    int32_t out_radius = 0;
    uint32_t out_tiles = 0;
    uint2 out_rect = make_uint2(0u, 0u);
    uint32_t out_key = 0xFFFFFFFFu;
    float out_depth = __int_as_float(0x7f800000);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;

    const float px = s_mean[3 * threadIdx.x], py = s_mean[3 * threadIdx.x + 1], pz = s_mean[3 * threadIdx.x + 2];
    const float tx = affine_row(s_view, 0, px, py, pz);
    const float ty = affine_row(s_view, 1, px, py, pz);
    const float tz = affine_row(s_view, 2, px, py, pz);   // altitude (p_view.z)

    float c3[6];
    if (cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = __ldg(cov3D_precomp + 6 * (size_t)idx + k);
    } else {
        float4 q = __ldg(rotations + idx);
        float sx = s_scale[3 * threadIdx.x], sy = s_scale[3 * threadIdx.x + 1], sz = s_scale[3 * threadIdx.x + 2];
        if (RAW) {
            sx = expf(sx); sy = expf(sy); sz = expf(sz);
            q = act_normalize(q, quat_norm_clamped(q));
        }
        cov3d_from_scale_rot(sx, sy, sz, scale_modifier, q, c3);
    }

    const Affine2x3 T = make_T(s_view, W, H);
    float cxx, cxy, cyy;
    cov2d_from_cov3d(T, c3, cxx, cxy, cyy);

    // forward.cu:216-235
    const float b2 = __fmul_rn(cxy, cxy);
    const float det_cov = __fmaf_rn(cxx, cyy, -b2);
    const float a = __fadd_rn(cxx, 0.3f), c = __fadd_rn(cyy, 0.3f);
    const float det = __fmaf_rn(a, c, -b2);
    float aa_scale = 1.0f;
    if (antialiasing) aa_scale = __fsqrt_rn(fmaxf(0.000025f, __fdiv_rn(det_cov, det)));

    if (det != 0.0f) {
        const float det_inv = __fdiv_rn(1.f, det);
        const float conic_x = __fmul_rn(c, det_inv);
        const float conic_y = __fmul_rn(-cxy, det_inv);
        const float conic_z = __fmul_rn(a, det_inv);

        // forward.cu:242-250
        const float mid = __fmul_rn(__fadd_rn(a, c), 0.5f);
        const float root = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));
        const float lam = fmaxf(__fadd_rn(mid, root), __fsub_rn(mid, root));
        const int radius = __float2int_ru(__fmul_rn(__fsqrt_rn(lam), 3.f));
        const float mx = ndc_to_pix(tx, W), my = ndc_to_pix(ty, H);

        // getRect (auxiliary.h:45-55): the radius is passed as int and converted back
        const float rf = (float)radius;
        const int x0 = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fsub_rn(mx, rf), 0.0625f))));
        const int y0 = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fsub_rn(my, rf), 0.0625f))));
        const int x1 = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(mx, rf), 16.f), -1.f), 0.0625f))));
        const int y1 = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(my, rf), 16.f), -1.f), 0.0625f))));
        const uint32_t area = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        // Tile sharding: only the rows of this call's band are emitted; radii keep the whole-image
        // meaning (identical on every rank), tiles / rect are the band's.
        const int y0b = min(band_y1, max(band_y0, y0)), y1b = min(band_y1, max(band_y0, y1));
        const uint32_t area_band = (uint32_t)(x1 - x0) * (uint32_t)(y1b - y0b);

        if (area != 0u) {
            // depth = 200.0 - altitude; the reference traps when it is negative (forward.cu:267-272).
            // We record the condition instead of killing the context, and cull the Gaussian.
            const float d = __fsub_rn(200.0f, tz);
            if (d < 0.f) {
                atomicOr(&info->error, EOGS_ERR_ALTITUDE_ABOVE_200);
            } else {
                out_radius = radius;
                out_tiles = area_band;
                out_rect = make_uint2((uint32_t)x0 | ((uint32_t)y0b << 16), (uint32_t)x1 | ((uint32_t)y1b << 16));
                out_depth = d;
                // sort key: depth bits; a Gaussian without a tile in THIS band sorts behind every contributing one, like the
                // culled ones, so that the first min(P, I) entries of the depth order hold all the work of the binning
                out_key = area_band != 0u ? __float_as_uint(d) : 0xFFFFFFFFu;
                float op_in = __ldg(opacities + idx);
                if (RAW) op_in = act_sigmoid(op_in);
                const float op = __fmul_rn(op_in, aa_scale);
                const float* col = s_color + CIN * threadIdx.x;
                float cc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                if (RAW) {
                    cc[0] = act_sh2rgb(col[0]); cc[1] = act_sh2rgb(col[1]); cc[2] = act_sh2rgb(col[2]);
                    cc[3] = __fadd_rn(__fmaf_rn(pz, __ldg(alt_affine + 2), __fmaf_rn(py, __ldg(alt_affine + 1),
                                      __fmul_rn(px, __ldg(alt_affine)))), __ldg(alt_affine + 3));
                    cc[4] = 1.f;
                } else {
#pragma unroll
                    for (int k = 0; k < C; k++) cc[k] = col[k];
                }
                r0 = make_float4(mx, my, conic_x, conic_y);
                r1 = make_float4(conic_z, op, cc[0], cc[1]);
                r2 = make_float4(cc[2], cc[3], cc[4], __fdiv_rn(1.f, d));
            }
        }
    }

    radii[idx] = out_radius;
    tiles[idx] = out_tiles;
    rect[idx] = out_rect;
    depth[idx] = out_depth;
    key_in[idx] = out_key;
    id_in[idx] = (uint32_t)idx;
    // I = sum of tiles touched is published here, by one red.global per warp, instead of after the depth sort and
    // scan: the host needs it to size the binning buffers, and its device -> host copy can then overlap those
    // two library calls (cabi.cu: forward_geometry_impl).  The reference's 32-bit scan wraps silently past 2^32
    // instances; here a wrap raises EOGS_ERR_TOO_MANY_INSTANCES (the host then asks for tile bands) instead of sizing
    // point_list from a small wrapped count.
    {
        const uint32_t active = __activemask();
        const bool huge = out_tiles > 0x03FFFFFFu;                       // keeps the warp sum below 2^31
        const uint32_t wsum = __reduce_add_sync(active, huge ? 0u : out_tiles);
        if (huge) atomicOr(&info->error, EOGS_ERR_TOO_MANY_INSTANCES);
        if (lane_id() == (uint32_t)(__ffs(active) - 1) && wsum) {
            const uint32_t old = atomicAdd(&info->num_instances, wsum);
            if (old + wsum < old) atomicOr(&info->error, EOGS_ERR_TOO_MANY_INSTANCES);
        }
        // Small scenes (host_info != nullptr): the LAST warp of the grid publishes the finished words straight into the
        // host's pinned struct, which is mapped into the device address space (payload, system fence, `ready`) — no
        // stream-ordered copies stand between this kernel and the depth sort, whose launch latency a 50 k-Gaussian
        // scene feels (BASELINE configs[0]: -15 us of 254).  One device-scope fence per warp orders its contribution
        // before its arrival; at 10^6 Gaussians those 31 k fences cost more than the copies (+21 us), so the caller
        // only asks for this path below 2^18 Gaussians.
        if (host_info) {
            __syncwarp(active);                                          // the lanes' error bits are ordered before the arrival below
            if (lane_id() == (uint32_t)(__ffs(active) - 1)) {
                __threadfence();
                const uint32_t warps = (uint32_t)(P + 31) >> 5;          // warps with at least one Gaussian
                if (atomicAdd(&info->reserved, 1u) + 1u == warps) {
                    __threadfence();
                    const uint32_t n = atomicAdd(&info->num_instances, 0u), e = atomicOr(&info->error, 0u);
                    host_info->num_instances = n;
                    host_info->error = e;
                    __threadfence_system();
                    host_info->ready = 1u;
                    info->ready = 1u;
                }
            }
        }
    }
    float4* rec = splat + (size_t)idx * REC_F4;
    rec[0] = r0; rec[1] = r1; rec[2] = r2;
    alpha_cut[idx] = out_radius > 0 ? alpha_cut_of(r1.y) : __int_as_float(0x7f800000);
}

// Inspection (tests): alpha_cut_of on an array of opacities, plus the two facts that define it —
// flags bit 0: the forward's test accepts power = cut; bit 1: it still accepts the next float below cut.
__global__ void alpha_cut_debug_kernel(int n, const float* __restrict__ op, float* __restrict__ cut,
                                       uint32_t* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float o = op[i], c = alpha_cut_of(o);
    cut[i] = c;
    uint32_t f = 0u;
    if (c <= 0.f && c > -3.0e38f) {
        const uint32_t k = __float_as_uint(c) | 0x80000000u;
        if (!(__fmul_rn(o, expf_as_forward(__uint_as_float(k))) < 1.0f / 255.0f)) f |= 1u;
        if (!(__fmul_rn(o, expf_as_forward(__uint_as_float(k + 1u))) < 1.0f / 255.0f)) f |= 2u;
    }
    flags[i] = f;
}

int launch_alpha_cut_debug(cudaStream_t s, int n, const float* op, float* cut, uint32_t* flags) {
    if (n <= 0) return 0;
    alpha_cut_debug_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, op, cut, flags);
    EOGS_LAUNCH_CHECK("alpha_cut_debug_kernel");
    return 0;
}

int launch_preprocess_fwd(cudaStream_t s, int P, int W, int H, Band band, int channels, bool raw_params,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* view, const float* alt_affine, float scale_modifier, bool antialiasing,
                          int32_t* radii, char* geom, const GeomLayout& L, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host_mapped)
{
    const int grid_x = (W + TILE - 1) / TILE, grid_y = (H + TILE - 1) / TILE;
    const int blocks = (P + PRE_THREADS - 1) / PRE_THREADS;
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int align_mask = (a16(means3D) ? 1 : 0) | (a16(scales) ? 2 : 0) | (a16(colors) ? 4 : 0);
    if (rotations && !a16(rotations)) { set_error("rotations must be 16-byte aligned"); return -2; }
    auto args = [&](auto kernel) {
        kernel<<<blocks, PRE_THREADS, 0, s>>>(
            P, W, H, grid_x, grid_y, band.row_begin, band.row_end, means3D, scales, reinterpret_cast<const float4*>(rotations),
            cov3D_precomp, opacities, colors, view, alt_affine, scale_modifier, antialiasing, align_mask, radii,
            reinterpret_cast<float4*>(geom + L.splat), reinterpret_cast<float*>(geom + L.cut),
            reinterpret_cast<float*>(geom + L.depth),
            reinterpret_cast<uint2*>(geom + L.rect), reinterpret_cast<uint32_t*>(geom + L.tiles),
            reinterpret_cast<uint32_t*>(geom + L.key_in), reinterpret_cast<uint32_t*>(geom + L.order),
            info_dev, info_host_mapped);
    };
    if (raw_params) {
        if (channels != 5 || cov3D_precomp || !alt_affine) { set_error("the fused-parameter path renders 5 channels from scales+rotations"); return -1; }
        args(preprocess_fwd_kernel<5, true>);
    }
    else if (channels == 5) args(preprocess_fwd_kernel<5, false>);
    else if (channels == 3) args(preprocess_fwd_kernel<3, false>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    EOGS_LAUNCH_CHECK("preprocess_fwd_kernel");
    return 0;
}

}  // namespace eogs
