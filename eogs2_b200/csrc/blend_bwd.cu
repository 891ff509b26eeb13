// Backward blend, tile-per-warp: ONE WARP replays one 16x16 tile back to front, each lane owning
// 8 pixels (one in each of the tile's eight 8x4 patches).  Replaces renderCUDA<5> backward
// (DGR/cuda_rasterizer/backward.cu:458-643).
//
// Why a warp per tile.  The reference issues 12 atomicAdd(float) per contributing (pixel,
// Gaussian) pair (backward.cu:598,608,631-640).  The 6+C per-pair terms have to be summed over
// the pixels a Gaussian touches; with one pixel per thread that is a 32-lane transposing
// butterfly (13 shuffles + 26 selects + 13 adds) plus one red.global per (8x4 patch, Gaussian),
// and ncu showed that reduction to be 35 % of all issued instructions of the previous
// 256-thread kernel (profiles/r01a_*).  Here a lane first accumulates the terms of its 8 pixels
// in registers (the per-pair products fold into FFMAs: free), so the butterfly and the
// red.global.add.f32 run once per (tile, Gaussian) — 3.5x fewer on the bench scene — and there is
// no block barrier and no cross-warp staging at all: everything is warp-synchronous.
//
//   staging   lane l fetches list entry (first - l) (3 x LDG.128, prefetched one batch ahead),
//             computes its exact 8-bit patch mask (blend_common.cuh: which 8x4 patches can see
//             alpha >= 1/255; patches whose pixels all stopped earlier are masked too), and
//             stores the record to the warp's shared-memory stage; a ballot of non-empty masks
//             is then walked bit by bit (back to front).
//   replay    per surviving entry and per set patch bit: recompute G and alpha with the
//             forward's exact expression, vote, and for accepted pixels update T, the running
//             "colour behind" dot product and the 11 accumulators.  Per-pixel constants
//             (dL_dpixel, T_final * bg.dL_dpixel) live in shared memory, lane-contiguous
//             (conflict-free LDS.128), the dynamic state (T, accum, n_contrib) in registers.
//   flush     transposing butterfly over the 32 lanes -> lane with slot k holds total k ->
//             one red.global.add.f32 warp instruction into the Gaussian's 64-byte record.
//
// Other differences from the reference kernel (unchanged from the first version)
//   - the replay starts at the tile's max(n_contrib): entries nobody blended are never fetched;
//   - the per-channel accum_rec / last_color recurrences (backward.cu:584-596) collapse into one
//     scalar recurrence on the dot product with dL_dpixel (same linear map);
//   - dL_dinvdepth per Gaussian is not accumulated: the reference computes it and drops it
//     (backward.cu:305-307 commented out; it never reaches a returned gradient).
//
// Bound: FP32 issue.  Algorithmic HBM bytes: 4 B id + 48 B record gathered per instance up to the
// tile's max(n_contrib), 44 B reduced per surviving (tile, Gaussian), 4*(C+1) + 8 B per pixel in.
#include "blend_common.cuh"
#include "f32x2.cuh"

namespace eogs {

#ifndef EOGS_BWD_WARPS
#define EOGS_BWD_WARPS 4                     // tiles (= warps) per CTA: 4 = a 2x2 block of tiles, 2 = 2x1, 1 = one tile
#endif
constexpr int BWD_WARPS = EOGS_BWD_WARPS;
constexpr int BWD_WX = BWD_WARPS >= 2 ? 2 : 1, BWD_WY = BWD_WARPS >= 4 ? 2 : 1;
static_assert(BWD_WX * BWD_WY == BWD_WARPS, "EOGS_BWD_WARPS must be 1, 2 or 4");
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr int NPATCH = (TILE / PATCH_W) * (TILE / PATCH_H);   // 8
constexpr int NSTRIP = NPATCH / 2;           // 4 strips of 16x4 pixels = a left and a right patch

// Accuracy switches for A/B builds (tools/fuzz_check.py): accurate expf / IEEE division instead of
// ex2.approx / rcp.approx for the VALUES (the accept decision never depends on them, see alpha_cut).
#ifndef EOGS_BWD_EXACT_EXP
#define EOGS_BWD_EXACT_EXP 0
#endif
#ifndef EOGS_BWD_EXACT_DIV
#define EOGS_BWD_EXACT_DIV 0
#endif
// Tuning switches (A/B measured on B200): where the per-pixel dynamic state lives.
#ifndef EOGS_BWD_STATE_SMEM
#define EOGS_BWD_STATE_SMEM 0                // 1: T / accum in shared memory (-16 registers, +2 LDS/STS per strip)
#endif

// Per-pixel constants of a lane's pixel PAIR in strip r (left patch 2r, right patch 2r+1), laid out
// as the f32x2 operands the replay consumes: an LDS.128 lands two ready-made register pairs.
struct BwdWarpSmem {
    float4 rec[2][32][REC_F4];        // two stages of 32 packed records (cp.async destinations, 48 B each)
    uint32_t rid[2][32];              // Gaussian id of each staged record
    float cut[2][32];                 // alpha_cut of each staged record: accept iff power >= cut
    float4 pix[NSTRIP][4][32];        // [0] = {g0.L, g0.R, g1.L, g1.R}   [1] = {g2.L, g2.R, g3.L, g3.R}
                                      // [2] = {g4.L, g4.R, ginv.L, ginv.R}
                                      // [3] = {-T_final (bg . g).L, same .R, n_contrib.L, n_contrib.R (int bits)}
#if EOGS_BWD_STATE_SMEM
    float4 state[NSTRIP][32];         // dynamic per-pixel-pair state {T.L, T.R, accum.L, accum.R}
#endif
};                                    // g = dL_dpixel

// 4 CTAs (16 warps) per SM measured best: 3 (142 regs) and 5 (96 regs) are both 14 % slower.
template <int C>
__global__ void __launch_bounds__(BWD_THREADS, 16 / BWD_WARPS)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ splat, const float* __restrict__ alpha_cut,
                 const float* __restrict__ bg, int W, int H,
                 int tiles_x, int tiles_y, int band_row0, int band_h,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinvdepth,
                 float* __restrict__ grad_rec)
{
    constexpr int NV = 6 + C;   // mean2D.xy, conic.xyw, opacity, colours
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char s_dyn[];       // BWD_WARPS x BwdWarpSmem (> 48 KB: opt-in)
    BwdWarpSmem* s_warp = reinterpret_cast<BwdWarpSmem*>(s_dyn);

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // tiles_y = tile rows of the band; brow = row inside the band, tile_y = row in the image
    const int tile_x = (int)blockIdx.x * BWD_WX + (int)(warp % BWD_WX), brow = (int)blockIdx.y * BWD_WY + (int)(warp / BWD_WX);
    if (tile_x >= tiles_x || brow >= tiles_y) return;            // whole warp leaves; no block barriers below
    const int tile_y = brow + band_row0;
    BwdWarpSmem& sm = s_warp[warp];

    const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);
    const float img_x1 = (float)(W - 1), img_y1 = (float)(H - 1);
    const uint2 range = __ldg(ranges + (size_t)brow * tiles_x + tile_x);
    const uint32_t* list = point_list + range.x;

    // ---- per-pixel state: pixel (patch p, lane) = (tile_x*16 + 8*(p&1) + (lane&7), tile_y*16 + 4*(p>>1) + (lane>>3))
    const float pxl = tx0 + (float)(lane & 7u);
    const f2 neg_px = mk2(-pxl, -(pxl + (float)PATCH_W));       // dx = mean.x - px as an add
    const float py_lane = ty0 + (float)(lane >> 3);              // + 4r = the strip's pixel row (exact)

    float bgv[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) bgv[ch] = __ldg(bg + ch);

    int pmax[NPATCH], ncon[NPATCH];
#if !EOGS_BWD_STATE_SMEM
    f2 T2[NSTRIP], accum2[NSTRIP];
#endif
    int nmax = 0;
#pragma unroll
    for (int r = 0; r < NSTRIP; r++) {
        float g[2][5], g_inv[2], Tf[2], nbg[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = 2 * r + h;
            const int px = tile_x * TILE + PATCH_W * h + (int)(lane & 7u);
            const int py = tile_y * TILE + PATCH_H * r + (int)(lane >> 3);
            const bool inside = px < W && py < H;
            const size_t pix_id = (size_t)(py - band_row0 * TILE) * W + px;     // band-compact buffers
            ncon[p] = inside ? (int)__ldg(n_contrib + pix_id) : 0;
            Tf[h] = inside ? __ldg(final_T + pix_id) : 0.f;
            float bg_dot_g = 0.f;
#pragma unroll
            for (int ch = 0; ch < 5; ch++) {
                g[h][ch] = (ch < C && inside) ? __ldg(dL_dpix + (size_t)ch * band_h * W + pix_id) : 0.f;
                if (ch < C) bg_dot_g = fmaf(bgv[ch], g[h][ch], bg_dot_g);
            }
            g_inv[h] = (inside && dL_dinvdepth) ? __ldg(dL_dinvdepth + pix_id) : 0.f;
            nbg[h] = -Tf[h] * bg_dot_g;
            pmax[p] = __reduce_max_sync(FULL, ncon[p]);
            nmax = max(nmax, pmax[p]);
        }
        sm.pix[r][0][lane] = make_float4(g[0][0], g[1][0], g[0][1], g[1][1]);
        sm.pix[r][1][lane] = make_float4(g[0][2], g[1][2], g[0][3], g[1][3]);
        sm.pix[r][2][lane] = make_float4(g[0][4], g[1][4], g_inv[0], g_inv[1]);
        sm.pix[r][3][lane] = make_float4(nbg[0], nbg[1], 0.f, 0.f);
#if EOGS_BWD_STATE_SMEM
        sm.state[r][lane] = make_float4(Tf[0], Tf[1], 0.f, 0.f);
#else
        T2[r] = mk2(Tf[0], Tf[1]);
        accum2[r] = bc2(0.f);
#endif
    }
    if (nmax == 0) return;                     // entries [nmax, n) were blended by no pixel of this tile
    const int rounds = (nmax + 31) >> 5;

    // Batch b, lane l holds list position nmax-1 - (32 b + l): back to front.  Records travel
    // global -> shared with cp.async (no staging registers), one batch ahead of the replay.
    uint32_t id_next = 0;
    {
        const int p0 = nmax - 1 - (int)lane;
        if (p0 >= 0) {
            const uint32_t id = __ldg(list + p0);
            sm.rid[0][lane] = id;
            const float4* src = splat + (size_t)id * REC_F4;
#pragma unroll
            for (int k = 0; k < REC_F4; k++) cp_async16(&sm.rec[0][lane][k], src + k);
            cp_async4(&sm.cut[0][lane], alpha_cut + id);
        }
        cp_async_commit();
        if (p0 - 32 >= 0) id_next = __ldg(list + p0 - 32);
    }
    const int my_slot = fold_slot<NV>(lane);
    // The accumulators are kept un-scaled and un-signed; the constant factors of each gradient slot
    // are applied once per flush: mean2D gets -d(pixel)/d(ndc), the conic terms -1/2.
    const float slot_scale = my_slot == 0 ? -0.5f * (float)W : my_slot == 1 ? -0.5f * (float)H :
                             (my_slot >= 2 && my_slot <= 4) ? -0.5f : 1.f;

    int first = nmax - 1;                      // list position held by lane 0 in this batch
    for (int i = 0; i < rounds; i++, first -= 32) {
        const int pos = first - (int)lane;
        const int stage = i & 1;
        cp_async_wait<0>();                    // this lane's record of batch i has landed
        uint32_t m = 0u;
        if (pos >= 0) m = patch_mask(sm.rec[stage][lane][0], sm.rec[stage][lane][1], tx0, ty0, img_x1, img_y1);
#pragma unroll
        for (int p = 0; p < NPATCH; p++)
            if (pos >= pmax[p]) m &= ~(1u << p);               // every pixel of patch p stopped before pos
        uint32_t todo = __ballot_sync(FULL, m != 0u);
        __syncwarp();                          // all lanes' records visible; previous batch's stage is free

        // next batch's records are in flight while this one is replayed
        if (pos - 32 >= 0) {
            sm.rid[stage ^ 1][lane] = id_next;
            const float4* src = splat + (size_t)id_next * REC_F4;
#pragma unroll
            for (int k = 0; k < REC_F4; k++) cp_async16(&sm.rec[stage ^ 1][lane][k], src + k);
            cp_async4(&sm.cut[stage ^ 1][lane], alpha_cut + id_next);
        }
        cp_async_commit();
        if (pos - 64 >= 0) id_next = __ldg(list + pos - 64);

        while (todo) {
            const int e = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t me = __shfl_sync(FULL, m, e);
            const int pos_e = first - e;
            const float4 ra = sm.rec[stage][e][0];     // mean.x, mean.y, conic.x, conic.y
            const float4 rb = sm.rec[stage][e][1];     // conic.z, opacity, c0, c1
            const float4 rc = sm.rec[stage][e][2];     // c2, c3, c4, 1/depth
            const float col[5] = {rb.z, rb.w, rc.x, rc.y, rc.z};
            const float cut_e = sm.cut[stage][e];      // the forward accepted a pixel of this entry iff power >= cut_e

            f2 v2[NV];
#pragma unroll
            for (int k = 0; k < NV; k++) v2[k] = bc2(0.f);
            bool any;

            // the forward's exponent in its op order (pair_power), on the lane's (left, right) pixel pair
            const f2 dx2 = add2(bc2(ra.x), neg_px);
            const f2 zdx2 = mul2(bc2(ra.z), dx2);
            const f2 wdx2 = mul2(bc2(ra.w), dx2);

            // One 16x4 strip per step = one pixel PAIR per lane, all arithmetic packed (FFMA2).  Written
            // branch-free: a rejected pixel runs the same arithmetic with alpha = G = 0, which leaves T,
            // accum and every accumulator unchanged.  A patch whose mask bit is clear cannot be accepted
            // (the mask is conservative), so the bits only decide whether the strip is visited at all.
            // Stage A, all four strips up front and branch-free: G, alpha and the accept test of the lane's
            // 8 pixels depend only on geometry, never on the replay state, so the four chains
            // (FFMA2 -> ex2 -> min -> compare) are independent and overlap each other's latency.
            f2 a2[NSTRIP], Gv2[NSTRIP];
            uint32_t live = 0u;                                   // bit r: some pixel of strip r accepts this entry
#pragma unroll
            for (int r = 0; r < NSTRIP; r++) {
                const float dy = __fsub_rn(ra.y, py_lane + (float)(PATCH_H * r));
                const float cyy = __fmul_rn(__fmul_rn(rb.x, dy), dy);
                const f2 quad2 = fma2(dx2, zdx2, bc2(cyy));
                const f2 power2 = fma2(quad2, bc2(-0.5f), mul2(wdx2, bc2(-dy)));
                // exp through ex2.approx (relative error ~2^-22): the VALUES carry a 1e-3 bar.  The accept
                // DECISION alpha >= 1/255 must be the forward's, or a pixel on that contour flips and a whole
                // term appears / disappears (1e-3..1e-2 in small scenes, tools/fuzz_parity.py).  It is taken on
                // the exponent instead: power >= alpha_cut, the per-Gaussian threshold the preprocess derived
                // from the forward's own expf (geom_math.cuh: alpha_cut_of) — same decision, same instruction count.
#if EOGS_BWD_EXACT_EXP
                const float G0 = expf(lo2(power2)), G1 = expf(hi2(power2));
#else
                const f2 pl2 = mul2(power2, bc2(1.4426950408889634f));
                const float G0 = ex2_approx(lo2(pl2)), G1 = ex2_approx(hi2(pl2));
#endif
                const f2 og2 = mul2(bc2(rb.y), mk2(G0, G1));
                const float al0 = fminf(0.99f, lo2(og2)), al1 = fminf(0.99f, hi2(og2));
                // entry at list position pos_e is blended by a pixel iff pos_e < n_contrib (backward.cu:556-558)
                const bool v0 = pos_e < ncon[2 * r] && !(lo2(power2) > 0.0f) && !(lo2(power2) < cut_e);
                const bool v1 = pos_e < ncon[2 * r + 1] && !(hi2(power2) > 0.0f) && !(hi2(power2) < cut_e);
                a2[r] = mk2(v0 ? al0 : 0.f, v1 ? al1 : 0.f);
                Gv2[r] = mk2(v0 ? G0 : 0.f, v1 ? G1 : 0.f);
                if (__any_sync(FULL, v0 || v1)) live |= 1u << r;
            }
            (void)me;
            any = live != 0u;

            // Stage B: the sequential part (T, accum recurrences).  One 16x4 strip per step = one pixel
            // PAIR per lane, all arithmetic packed (FFMA2).  Branch-free inside a strip: a rejected
            // pixel runs the same arithmetic with alpha = G = 0, which leaves T, accum and every
            // accumulator unchanged.
#pragma unroll
            for (int r = 0; r < NSTRIP; r++) {
                if (!((live >> r) & 1u)) continue;               // warp-uniform
                const float dy = __fsub_rn(ra.y, py_lane + (float)(PATCH_H * r));
                const float4 pa = sm.pix[r][0][lane], pb = sm.pix[r][1][lane];
                const float4 pc = sm.pix[r][2][lane], pd = sm.pix[r][3][lane];
                const f2 g2[5] = {mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(pb.x, pb.y), mk2(pb.z, pb.w), mk2(pc.x, pc.y)};
                const f2 ginv2 = mk2(pc.z, pc.w), nbg2 = mk2(pd.x, pd.y);
#if EOGS_BWD_STATE_SMEM
                const float4 st = sm.state[r][lane];
                const f2 Told2 = mk2(st.x, st.y), accum_old2 = mk2(st.z, st.w);
#else
                const f2 Told2 = T2[r], accum_old2 = accum2[r];
#endif
                const f2 om2 = fma2(a2[r], bc2(-1.f), bc2(1.f));                // 1 - alpha
#if EOGS_BWD_EXACT_DIV
                const f2 inv2 = mk2(__fdiv_rn(1.f, lo2(om2)), __fdiv_rn(1.f, hi2(om2)));
#else
                const f2 inv2 = mk2(fast_rcp(lo2(om2)), fast_rcp(hi2(om2)));    // exactly 1 for a rejected pixel
#endif
                const f2 Tn2 = mul2(Told2, inv2);
                const f2 w2 = mul2(a2[r], Tn2);
                f2 cg2 = mul2(bc2(rc.w), ginv2);
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    fma2_acc(v2[6 + ch], w2, g2[ch]);
                    fma2_acc(cg2, bc2(col[ch]), g2[ch]);
                }
                // accum = (colour blended behind this entry) . dL_dpixel
                const f2 behind2 = fma2(accum_old2, bc2(-1.f), cg2);
                const f2 dLa2 = fma2(nbg2, inv2, mul2(behind2, Tn2));           // dL/dalpha (finite; x0 below if rejected)
                const f2 acc_new2 = fma2(a2[r], behind2, accum_old2);
#if EOGS_BWD_STATE_SMEM
                sm.state[r][lane] = make_float4(lo2(Tn2), hi2(Tn2), lo2(acc_new2), hi2(acc_new2));
#else
                T2[r] = Tn2; accum2[r] = acc_new2;
#endif
                const f2 dG2 = mul2(bc2(rb.y), dLa2);                            // dL/dG
                const f2 gdx2 = mul2(Gv2[r], dx2), gdy2 = mul2(Gv2[r], bc2(dy));
                fma2_acc(v2[0], dG2, fma2(gdx2, bc2(ra.z), mul2(gdy2, bc2(ra.w))));
                fma2_acc(v2[1], dG2, fma2(gdy2, bc2(rb.x), mul2(gdx2, bc2(ra.w))));
                const f2 hgx2 = mul2(dG2, gdx2);
                fma2_acc(v2[2], hgx2, dx2);
                fma2_acc(v2[3], hgx2, bc2(dy));
                fma2_acc(v2[4], mul2(dG2, gdy2), bc2(dy));
                fma2_acc(v2[5], Gv2[r], dLa2);
            }
            if (any) {
                float v[NV];
#pragma unroll
                for (int k = 0; k < NV; k++) v[k] = lo2(v2[k]) + hi2(v2[k]);
                warp_transpose_reduce<NV>(v, lane);
                const uint32_t gid = sm.rid[stage][e];
                if (my_slot >= 0) atomicAdd(grad_rec + (size_t)gid * GRAD_STRIDE + my_slot, v[0] * slot_scale);
            }
        }
        __syncwarp();
    }
}

int launch_blend_bwd(cudaStream_t s, int W, int H, Band band, int channels, const char* geom,
                     const GeomLayout& GL, const uint32_t* point_list, const char* image,
                     const ImageLayout& IL, const float* bg, const float* dL_dpix,
                     const float* dL_dinvdepth, float* grad_rec)
{
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = band.rows();
    const dim3 grid((tiles_x + BWD_WX - 1) / BWD_WX, (tiles_y + BWD_WY - 1) / BWD_WY, 1);
    constexpr size_t smem = sizeof(BwdWarpSmem) * BWD_WARPS;
    cudaError_t attr_err = cudaSuccess;
    auto run = [&](auto kernel) {
        attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr_err != cudaSuccess) return;
        kernel<<<grid, BWD_THREADS, smem, s>>>(
            reinterpret_cast<const uint2*>(image + IL.ranges), point_list,
            reinterpret_cast<const float4*>(geom + GL.splat), reinterpret_cast<const float*>(geom + GL.cut),
            bg, W, H, tiles_x, tiles_y,
            band.row_begin, band.height(H), reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), dL_dpix, dL_dinvdepth, grad_rec);
    };
    if (channels == 5) run(blend_bwd_kernel<5>);
    else if (channels == 3) run(blend_bwd_kernel<3>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(blend_bwd_kernel)");
    EOGS_LAUNCH_CHECK("blend_bwd_kernel");
    return 0;
}

}  // namespace eogs
