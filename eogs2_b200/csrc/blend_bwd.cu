// Backward blend, tile-per-warp: ONE WARP replays one 16x16 tile back to front, each lane owning
// 8 pixels: the pixel PAIR (x, y), (x, y+4) — x = lane&7, y = lane>>3 — of each of the tile's four
// 8x8 REGIONS, the forward's own pixel-to-lane map (blend_fwd.cu), so the forward's per-region
// culling masks select exactly the units this kernel evaluates.  Replaces renderCUDA<5> backward
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
//   staging   lane l owns list entry (first - l): it reads the entry's culling mask — one byte per instance,
//             written by the forward blend (which 8x8 regions of the tile the Gaussian can reach with
//             alpha >= 1/255; regions whose pixels all stopped earlier are masked too) — and only if the
//             mask is non-empty gathers the record into the warp's shared-memory stage (cp.async, one batch
//             ahead); a ballot of non-empty masks is then walked bit by bit (back to front).
//   replay    per surviving entry and per set region bit: recompute G and alpha with the
//             forward's exact expression, vote, and for accepted pixels update T, the running
//             "colour behind" dot product and the 11 accumulators.  Per-pixel constants
//             (dL_dpixel, T_final * bg.dL_dpixel, n_contrib) live in shared memory, lane-contiguous
//             (conflict-free LDS.128), the dynamic state (T, accum) in registers.
//   flush     the 6+C per-lane partial sums (plain sums: every per-Gaussian constant factor is applied later, once per
//             Gaussian, by preprocess_bwd_kernel) cross the warp through shared memory: 11 conflict-free STS
//             ([value][lane]), then lane (k, h) adds half a row (4 x LDS.128, packed adds), one shuffle
//             joins the halves and lanes 0..10 issue ONE red.global.add.f32 warp instruction into the
//             Gaussian's 64-byte record.  (The transposing shuffle butterfly it replaces — 13 SHFL + 22
//             FSEL + 24 FADD, five dependent levels — was 24 % of the kernel's stall samples, ncu r2a.)
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

#if EOGS_COUNT_PAIRS
__device__ unsigned long long g_counters_bwd[CNT_COUNT];
#endif
int read_counters_bwd(unsigned long long* out, bool reset) {
#if EOGS_COUNT_PAIRS
    unsigned long long tmp[CNT_COUNT];
    EOGS_CUDA(cudaMemcpyFromSymbol(tmp, g_counters_bwd, sizeof(tmp)));
    for (int i = CNT_BWD_EVAL; i < CNT_BWD_EVAL + 5; i++) out[i] = tmp[i];
    if (reset) {
        unsigned long long z[CNT_COUNT] = {};
        EOGS_CUDA(cudaMemcpyToSymbol(g_counters_bwd, z, sizeof(z)));
    }
#else
    (void)out; (void)reset;
#endif
    return 0;
}

constexpr int BWD_WARPS = 4;                 // warps per CTA; every warp is an independent worker (no block barrier)
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr int NREGION = 4;                   // 8x8 regions of a tile: bit q of a culling mask = region (8*(q&1), 8*(q>>1))
constexpr int RED_STRIDE = 36;

// Accuracy switches for A/B builds (tools/fuzz_check.py): accurate expf / IEEE division instead of
// ex2.approx / rcp.approx for the VALUES (the accept decision never depends on them, see alpha_cut).
#ifndef EOGS_BWD_EXACT_EXP
#define EOGS_BWD_EXACT_EXP 0
#endif
#ifndef EOGS_BWD_EXACT_DIV
#define EOGS_BWD_EXACT_DIV 0
#endif
// 1: the per-pixel n_contrib of the lane's 8 pixels are read from shared memory (the free half of pix[q][3]) instead
// of living in 8 registers — the kernel sits at the 128-register cap, where every live value less counts (DESIGN.md
// section 4 on what ptxas does at the cap).  Measured 1.124 (1) vs 1.129 ms (0) on config 2.
#ifndef EOGS_BWD_NCON_SMEM
#define EOGS_BWD_NCON_SMEM 1
#endif

// Scheduling.  The warps are persistent: every warp pulls the next tile from a global queue (one atomicAdd per
// tile) that the forward ordered longest-first (tile_order_kernel below: descending max(n_contrib), tiles nobody
// blended left out).  A warp replays ~7 tiles per launch at the bench size and a heavy tile costs several average
// ones; a static tile -> CTA map left 15 % of the warp slots idle (ncu sm__warps_active 13.6 of 16).

// Per-pixel constants of a lane's pixel PAIR in region q (rows y and y+4: halves .a and .b), laid out
// as the f32x2 operands the replay consumes: an LDS.128 lands two ready-made register pairs.
struct BwdWarpSmem {
    float4 rec[2][32][REC_F4];        // two stages of 32 packed records (cp.async destinations, 48 B each)
    uint32_t rid[2][32];              // Gaussian id of each staged record
    float cut[2][32];                 // alpha_cut of each staged record: accept iff power >= cut
    int rmax[4];                      // max(n_contrib) over each 8x8 region (warp-uniform; read once per batch)
    float red[6 + EOGS_MAX_CHANNELS][RED_STRIDE];   // flush: [value][lane], rows padded to 36 floats (LDS.128 of 8 lanes hit 8 bank groups)
    float4 pix[NREGION][4][32];       // [0] = {g0.a, g0.b, g1.a, g1.b}   [1] = {g2.a, g2.b, g3.a, g3.b}
                                      // [2] = {g4.a, g4.b, ginv.a, ginv.b}
                                      // [3] = {-T_final (bg . g).a, same .b, n_contrib.a, n_contrib.b (int bits)}
};                                    // g = dL_dpixel

// 4 CTAs (16 warps) per SM measured best: 3 (142 regs) and 5 (96 regs) are both 14 % slower.
template <int C>
__global__ void __launch_bounds__(BWD_THREADS, 4)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ splat, const float* __restrict__ alpha_cut,
                 const float* __restrict__ bg, int W, int H,
                 int tiles_x, int tiles_y, int band_row0, int band_h,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinvdepth,
                 float* __restrict__ grad_rec, const uint8_t* __restrict__ masks,
                 const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ sched, uint32_t* queue)
{
    constexpr int NV = 6 + C;   // mean2D.xy, conic.xyw, opacity, colours
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char s_dyn[];       // BWD_WARPS x BwdWarpSmem (> 48 KB: opt-in)
    BwdWarpSmem* s_warp = reinterpret_cast<BwdWarpSmem*>(s_dyn);

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    BwdWarpSmem& sm = s_warp[warp];
    const int red_k = (int)(lane % NV), red_h = (int)(lane / NV);   // lane (k, h) adds columns [16h, 16h+16) of row k
    const int my_slot = lane < (uint32_t)NV ? (int)lane : -1;
    // The accumulators are kept un-scaled and un-signed; the constant factors of each gradient slot
    // are applied once per flush: mean2D gets -d(pixel)/d(ndc), the conic terms -1/2 (and the opacity, below).

#if EOGS_COUNT_PAIRS
    unsigned long long cnt_eval = 0ull, cnt_blend = 0ull, cnt_slots = 0ull, cnt_entries = 0ull, cnt_flush = 0ull;   // warp-uniform
#endif
    const uint32_t n_active = __ldg(sched);                      // tiles with max(n_contrib) > 0, longest first
    // Small images: with fewer active tiles than resident warps the launch time is the replay of the LONGEST tile by
    // one warp while most of the device idles (config 1, 512^2 = 1024 tiles: 87 us).  A tile is then split over 2 or 4
    // warps by regions — a work item replays the tile's list for its region subset only (the other regions' n_contrib
    // read as 0, which masks them everywhere below) and flushes its own partial sums.
    const uint32_t slots = gridDim.x * BWD_WARPS;
    const uint32_t split_log2 = 2u * n_active <= slots ? 2u : n_active <= slots ? 1u : 0u;
    const uint32_t n_items = n_active << split_log2;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_items) break;
        const uint32_t tile = __ldg(tile_order + (qi >> split_log2));
        const uint32_t part = qi & ((1u << split_log2) - 1u);
        const uint32_t region_set = split_log2 == 2u ? 1u << part : split_log2 == 1u ? 3u << (2u * part) : 0xFu;
        const int tile_x = (int)(tile % (uint32_t)tiles_x), brow = (int)(tile / (uint32_t)tiles_x);
        (void)tiles_y;
        const int tile_y = brow + band_row0;
        const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);
        const uint2 range = __ldg(ranges + (size_t)brow * tiles_x + tile_x);
        const uint32_t* list = point_list + range.x;

        // ---- per-pixel state: pixel (region q, half h, lane) = (tile_x*16 + 8*(q&1) + (lane&7), tile_y*16 + 8*(q>>1) + 4h + (lane>>3))
        const float pxl = tx0 + (float)(lane & 7u);
        const f2 neg_px = mk2(-pxl, -(pxl + 8.f));               // dx = mean.x - px as an add: region columns 0 / 1
        const float py_lane = ty0 + (float)(lane >> 3);
        // dy = mean.y - py as an add, for the pixel pair of the upper (q = 0, 1) and the lower (q = 2, 3) regions
        const f2 neg_py_up = mk2(-py_lane, -(py_lane + (float)PATCH_H));
        const f2 neg_py_lo = mk2(-(py_lane + 8.f), -(py_lane + 8.f + (float)PATCH_H));

        int ncon[2 * NREGION];
        f2 T2[NREGION], accum2[NREGION];
        float bgv[C];                          // per tile, so that it does not hold registers during the replay
#pragma unroll
        for (int ch = 0; ch < C; ch++) bgv[ch] = __ldg(bg + ch);
        int nmax = 0;
        int rmax[NREGION];                     // max(n_contrib) over each 8x8 region
#pragma unroll
        for (int q = 0; q < NREGION; q++) {
            float g[2][5], g_inv[2], Tf[2], nbg[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int px = tile_x * TILE + 8 * (q & 1) + (int)(lane & 7u);
                const int py = tile_y * TILE + 8 * (q >> 1) + PATCH_H * h + (int)(lane >> 3);
                const bool inside = px < W && py < H && ((region_set >> q) & 1u);
                const size_t pix_id = (size_t)(py - band_row0 * TILE) * W + px;     // band-compact buffers
                ncon[2 * q + h] = inside ? (int)__ldg(n_contrib + pix_id) : 0;
                Tf[h] = inside ? __ldg(final_T + pix_id) : 0.f;
                float bg_dot_g = 0.f;
#pragma unroll
                for (int ch = 0; ch < 5; ch++) {
                    g[h][ch] = (ch < C && inside) ? __ldg(dL_dpix + (size_t)ch * band_h * W + pix_id) : 0.f;
                    if (ch < C) bg_dot_g = fmaf(bgv[ch], g[h][ch], bg_dot_g);
                }
                g_inv[h] = (inside && dL_dinvdepth) ? __ldg(dL_dinvdepth + pix_id) : 0.f;
                nbg[h] = -Tf[h] * bg_dot_g;
            }
            rmax[q] = __reduce_max_sync(FULL, max(ncon[2 * q], ncon[2 * q + 1]));     // warp-uniform
            nmax = max(nmax, rmax[q]);
            sm.pix[q][0][lane] = make_float4(g[0][0], g[1][0], g[0][1], g[1][1]);
            sm.pix[q][1][lane] = make_float4(g[0][2], g[1][2], g[0][3], g[1][3]);
            sm.pix[q][2][lane] = make_float4(g[0][4], g[1][4], g_inv[0], g_inv[1]);
            sm.pix[q][3][lane] = make_float4(nbg[0], nbg[1], __int_as_float(ncon[2 * q]), __int_as_float(ncon[2 * q + 1]));
            T2[q] = mk2(Tf[0], Tf[1]);
            accum2[q] = bc2(0.f);
        }
        if (nmax == 0) continue;               // (cannot happen for a queued tile; kept for safety)
        const int rounds = (nmax + 31) >> 5;

        // Batch b, lane l holds list position nmax-1 - (32 b + l): back to front.  An entry's culling mask comes from
        // the forward (one byte per instance, blend_fwd.cu: which 8x8 regions of the tile the Gaussian can reach);
        // regions whose pixels all stopped earlier are masked too.  Only entries with a non-empty mask are fetched:
        // their records travel global -> shared with cp.async (no staging registers), one batch ahead of the replay;
        // ids and masks are prefetched two batches ahead.
        const uint8_t* mlist = masks + range.x;
        if (lane == 0) *reinterpret_cast<int4*>(&sm.rmax[0]) = make_int4(rmax[0], rmax[1], rmax[2], rmax[3]);
        __syncwarp();
        auto live_mask = [&](uint32_t raw, int pos) {
            const int4 rm = *reinterpret_cast<const int4*>(&sm.rmax[0]);
            const int rmv[4] = {rm.x, rm.y, rm.z, rm.w};
            uint32_t m = pos >= 0 ? raw : 0u;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (pos >= rmv[q]) m &= ~(1u << q);                // every pixel of region q stopped before pos
            return m;
        };
        auto fetch = [&](int stage, uint32_t id) {
            sm.rid[stage][lane] = id;
            const float4* src = splat + (size_t)id * REC_F4;
#pragma unroll
            for (int k = 0; k < REC_F4; k++) cp_async16(&sm.rec[stage][lane][k], src + k);
            cp_async4(&sm.cut[stage][lane], alpha_cut + id);
        };
        uint32_t id_next = 0u, m = 0u, m_next = 0u;
        {
            const int p0 = nmax - 1 - (int)lane;
            if (p0 >= 0) m = live_mask(__ldg(mlist + p0), p0);
            if (m) fetch(0, __ldg(list + p0));
            cp_async_commit();
            if (p0 - 32 >= 0) {
                m_next = live_mask(__ldg(mlist + p0 - 32), p0 - 32);
                if (m_next) id_next = __ldg(list + p0 - 32);
            }
        }

        int first = nmax - 1;                  // list position held by lane 0 in this batch
        for (int i = 0; i < rounds; i++, first -= 32) {
            const int pos = first - (int)lane;
            const int stage = i & 1;
            cp_async_wait<0>();                // this lane's record of batch i has landed
            uint32_t todo = __ballot_sync(FULL, m != 0u);
            __syncwarp();                      // all lanes' records visible; previous batch's stage is free

            // next batch's records are in flight while this one is replayed
            if (m_next) fetch(stage ^ 1, id_next);
            cp_async_commit();
            const uint32_t m_cur = m;
            m = m_next;
            m_next = 0u;
            if (pos - 64 >= 0) {
                m_next = live_mask(__ldg(mlist + pos - 64), pos - 64);
                if (m_next) id_next = __ldg(list + pos - 64);
            }

            while (todo) {
                const int e = __ffs(todo) - 1;
                todo &= todo - 1u;
                const uint32_t me = __shfl_sync(FULL, m_cur, e);
                const int pos_e = first - e;
                const float4 ra = sm.rec[stage][e][0];     // mean.x, mean.y, conic.x, conic.y
                const float4 rb = sm.rec[stage][e][1];     // conic.z, opacity, c0, c1
                const float4 rc = sm.rec[stage][e][2];     // c2, c3, c4, 1/depth
                const float col[5] = {rb.z, rb.w, rc.x, rc.y, rc.z};
                const float cut_e = sm.cut[stage][e];      // the forward accepted a pixel of this entry iff power >= cut_e

                f2 v2[NV];
#pragma unroll
                for (int k = 0; k < NV; k++) v2[k] = bc2(0.f);

                // the forward's exponent in its op order (forward.cu:361-365 as compiled: blend_fwd.cu), on the
                // lane's pixel pair (rows y, y+4 of a region; the pair shares dx); every half of a packed operation
                // rounds like the scalar one, so `power` has the forward's bits and the accept decision below is
                // the forward's
                const f2 dxc2 = add2(bc2(ra.x), neg_px);               // .lo: region column 0, .hi: column 1
                const f2 zdxc2 = mul2(bc2(ra.z), dxc2), wdxc2 = mul2(bc2(ra.w), dxc2);
                const float dxv[2] = {lo2(dxc2), hi2(dxc2)}, zdxv[2] = {lo2(zdxc2), hi2(zdxc2)}, wdxv[2] = {lo2(wdxc2), hi2(wdxc2)};
                const f2 dyr2[2] = {add2(bc2(ra.y), neg_py_up), add2(bc2(ra.y), neg_py_lo)};     // region rows 0 / 1
                const f2 cyyr2[2] = {mul2(mul2(bc2(rb.x), dyr2[0]), dyr2[0]), mul2(mul2(bc2(rb.x), dyr2[1]), dyr2[1])};

                // Stage A, per region the entry's mask reaches: G, alpha and the accept test of the lane's
                // pixels depend only on geometry, never on the replay state, so the chains
                // (FFMA2 -> ex2 -> select -> min) of the regions are independent and overlap each other's
                // latency.  A rejected pixel gets G = alpha = 0, which leaves T, accum and every accumulator
                // unchanged in stage B.  A region whose mask bit is clear cannot be accepted (the mask is
                // conservative) and is not evaluated at all.
                f2 a2[NREGION], Gv2[NREGION];
                uint32_t lane_live = 0u;                              // bit q: a pixel of THIS lane in region q accepts this entry
                auto stage_a = [&](const int q) {
                    const int c = q & 1, rr = q >> 1;
                    const f2 quad2 = fma2(bc2(dxv[c]), bc2(zdxv[c]), cyyr2[rr]);
                    const f2 power2 = fma2(quad2, bc2(-0.5f), neg2(mul2(bc2(wdxv[c]), dyr2[rr])));
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
                    // entry at list position pos_e is blended by a pixel iff pos_e < n_contrib (backward.cu:556-558)
#if EOGS_BWD_NCON_SMEM
                    const int2 ncp = *reinterpret_cast<const int2*>(&sm.pix[q][3][lane].z);
                    const int nc0 = ncp.x, nc1 = ncp.y;
#else
                    const int nc0 = ncon[2 * q], nc1 = ncon[2 * q + 1];
#endif
                    const bool v0 = pos_e < nc0 && !(lo2(power2) > 0.0f) && !(lo2(power2) < cut_e);
                    const bool v1 = pos_e < nc1 && !(hi2(power2) > 0.0f) && !(hi2(power2) < cut_e);
                    Gv2[q] = mk2(v0 ? G0 : 0.f, v1 ? G1 : 0.f);       // one select per pixel: alpha follows from G
                    const f2 og2 = mul2(bc2(rb.y), Gv2[q]);
                    a2[q] = mk2(fminf(0.99f, lo2(og2)), fminf(0.99f, hi2(og2)));
                    lane_live |= (v0 || v1) ? (1u << q) : 0u;
#if EOGS_COUNT_PAIRS
                    cnt_eval += __popc(__ballot_sync(FULL, pos_e < nc0)) + __popc(__ballot_sync(FULL, pos_e < nc1));
                    cnt_blend += __popc(__ballot_sync(FULL, v0)) + __popc(__ballot_sync(FULL, v1));
                    cnt_slots += 64ull;
#endif
                };
                if (me == 0xFu) {                                     // straight-line: four independent chains
#pragma unroll
                    for (int q = 0; q < NREGION; q++) stage_a(q);
                } else {
#pragma unroll
                    for (int q = 0; q < NREGION; q++)
                        if ((me >> q) & 1u) stage_a(q);               // warp-uniform
                }
                const uint32_t live = __reduce_or_sync(FULL, lane_live);    // bit r: some pixel of strip r accepts (one REDUX)
                const bool any = live != 0u;
#if EOGS_COUNT_PAIRS
                cnt_entries += 1ull; cnt_flush += any ? 1ull : 0ull;
#endif

                // Stage B: the sequential part (T, accum recurrences).  One 8x8 region per step = one pixel
                // PAIR per lane, all arithmetic packed (FFMA2).  Branch-free inside a region.
                // Per-Gaussian sums kept per lane, with u = G dL/dalpha (so dL/dG = opacity u):
                //   v2[0] = sum u dx   v2[1] = sum u dy   v2[2] = sum u dx dx   v2[3] = sum u dx dy   v2[4] = sum u dy dy
                //   v2[5] = sum u      v2[6 + ch] = sum alpha T dL_dpixel[ch]
                // The conic and the opacity are constants of the entry, so the mean gradient
                // (backward.cu:631-632: dL_dG dG/ddelta) is assembled from the two first moments at flush time
                // instead of per pixel: 27 packed operations per region instead of 33.
                auto stage_b = [&](const int q) {
                    const float dx = dxv[q & 1];
                    const f2 dy2 = dyr2[q >> 1];
                    const float4 pa = sm.pix[q][0][lane], pb = sm.pix[q][1][lane];
                    const float4 pc = sm.pix[q][2][lane];
                    const float2 pd = *reinterpret_cast<const float2*>(&sm.pix[q][3][lane]);
                    const f2 g2[5] = {mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(pb.x, pb.y), mk2(pb.z, pb.w), mk2(pc.x, pc.y)};
                    const f2 ginv2 = mk2(pc.z, pc.w), nbg2 = mk2(pd.x, pd.y);
                    const f2 Told2 = T2[q], accum_old2 = accum2[q];
                    const f2 om2 = fma2(a2[q], bc2(-1.f), bc2(1.f));                // 1 - alpha
#if EOGS_BWD_EXACT_DIV
                    const f2 inv2 = mk2(__fdiv_rn(1.f, lo2(om2)), __fdiv_rn(1.f, hi2(om2)));
#else
                    const f2 inv2 = mk2(fast_rcp(lo2(om2)), fast_rcp(hi2(om2)));    // exactly 1 for a rejected pixel
#endif
                    const f2 Tn2 = mul2(Told2, inv2);
                    const f2 w2 = mul2(a2[q], Tn2);
                    f2 cg2 = mul2(bc2(rc.w), ginv2);
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        fma2_acc(v2[6 + ch], w2, g2[ch]);
                        fma2_acc(cg2, bc2(col[ch]), g2[ch]);
                    }
                    // accum = (colour blended behind this entry) . dL_dpixel
                    const f2 behind2 = fma2(accum_old2, bc2(-1.f), cg2);
                    const f2 dLa2 = fma2(nbg2, inv2, mul2(behind2, Tn2));           // dL/dalpha (finite; G = 0 below if rejected)
                    const f2 acc_new2 = fma2(a2[q], behind2, accum_old2);
                    T2[q] = Tn2; accum2[q] = acc_new2;
                    const f2 u2 = mul2(Gv2[q], dLa2);
                    const f2 ux2 = mul2(u2, bc2(dx)), uy2 = mul2(u2, dy2);
                    v2[0] = add2(v2[0], ux2);
                    v2[1] = add2(v2[1], uy2);
                    fma2_acc(v2[2], ux2, bc2(dx));
                    fma2_acc(v2[3], ux2, dy2);
                    fma2_acc(v2[4], uy2, dy2);
                    v2[5] = add2(v2[5], u2);
                };
#pragma unroll
                for (int q = 0; q < NREGION; q++) {
                    if (!((live >> q) & 1u)) continue;               // warp-uniform
                    stage_b(q);
                }
                if (any) {
                    const uint32_t gid = sm.rid[stage][e];
                    float v[NV];
#pragma unroll
                    for (int k = 0; k < NV; k++) v[k] = lo2(v2[k]) + hi2(v2[k]);
                    // The record receives the plain sums: everything that is constant per Gaussian — the opacity on the
                    // five geometric sums, the conic that turns the two first moments into dL_dmean2D, the -1/2 of the conic
                    // terms, the pixel -> ndc factors — is applied once per Gaussian by preprocess_bwd_kernel (linear, so it commutes with the sum
                    // over tiles) instead of once per (tile, Gaussian) here.
                    __syncwarp();                                    // the previous flush's reads are done
#pragma unroll
                    for (int k = 0; k < NV; k++) sm.red[k][lane] = v[k];
                    __syncwarp();
                    if (red_h < 2) {
                        const float4* row = reinterpret_cast<const float4*>(&sm.red[red_k][16 * red_h]);
                        const float4 q0 = row[0], q1 = row[1], q2 = row[2], q3 = row[3];
                        const f2 s01 = add2(add2(mk2(q0.x, q0.y), mk2(q0.z, q0.w)), add2(mk2(q1.x, q1.y), mk2(q1.z, q1.w)));
                        const f2 s23 = add2(add2(mk2(q2.x, q2.y), mk2(q2.z, q2.w)), add2(mk2(q3.x, q3.y), mk2(q3.z, q3.w)));
                        const f2 s = add2(s01, s23);
                        v[0] = lo2(s) + hi2(s);
                    }
                    v[0] += __shfl_down_sync(FULL, v[0], NV);        // lane k < NV: its half + the half of lane k + NV
                    if (my_slot >= 0) atomicAdd(grad_rec + (size_t)gid * GRAD_STRIDE + my_slot, v[0]);
                }
            }
            __syncwarp();
        }
        cp_async_wait<0>();                    // (the last round committed an empty group)
        __syncwarp();                          // the warp's stages and pixel constants are free for the next tile
    }
#if EOGS_COUNT_PAIRS
    if (lane == 0) {
        atomicAdd(&g_counters_bwd[CNT_BWD_EVAL], cnt_eval); atomicAdd(&g_counters_bwd[CNT_BWD_BLEND], cnt_blend);
        atomicAdd(&g_counters_bwd[CNT_BWD_SLOTS], cnt_slots); atomicAdd(&g_counters_bwd[CNT_BWD_ENTRIES], cnt_entries);
        atomicAdd(&g_counters_bwd[CNT_BWD_FLUSHES], cnt_flush);
    }
#endif
}

// Longest-first tile order for the persistent backward (one block; runs right after the forward blend).
// work[t] = max(n_contrib) of tile t (written by blend_fwd_kernel); order = tiles with work > 0 sorted by
// descending work (1024 buckets, order inside a bucket arbitrary); sched[0] = their number.
__global__ void __launch_bounds__(1024)
tile_order_kernel(uint32_t tiles, const uint32_t* __restrict__ work, uint32_t* __restrict__ order,
                  uint32_t* __restrict__ sched)
{
    constexpr int NB = 1024;
    __shared__ uint32_t s_hist[NB];
    __shared__ uint32_t s_red[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t mx = 0u;
    for (uint32_t t = tid; t < tiles; t += NB) mx = max(mx, __ldg(work + t));
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) s_red[wid] = mx;
    s_hist[tid] = 0u;
    __syncthreads();
    mx = s_red[lane];
    mx = __reduce_max_sync(0xffffffffu, mx);
    int shift = 0;
    while ((mx >> shift) >= (uint32_t)NB) shift++;
    // bucket 0 = heaviest: b = NB-1 - (work >> shift); work == 0 is left out
    for (uint32_t t = tid; t < tiles; t += NB) {
        const uint32_t w = __ldg(work + t);
        if (w) atomicAdd(&s_hist[NB - 1 - (w >> shift)], 1u);
    }
    __syncthreads();
    // exclusive scan of the 1024 bucket counts (one per thread)
    const uint32_t cnt = s_hist[tid];
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    __syncthreads();
    if (lane == 31) s_red[wid] = incl;
    __syncthreads();
    uint32_t wsum = s_red[lane];
    uint32_t wincl = wsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d);
        if (lane >= (uint32_t)d) wincl += o;
    }
    const uint32_t wbase = __shfl_sync(0xffffffffu, wincl - wsum, wid);
    s_hist[tid] = wbase + incl - cnt;          // bucket start; becomes the bucket's cursor
    if (tid == NB - 1) sched[0] = wbase + incl;
    __syncthreads();
    for (uint32_t t = tid; t < tiles; t += NB) {
        const uint32_t w = __ldg(work + t);
        if (w) order[atomicAdd(&s_hist[NB - 1 - (w >> shift)], 1u)] = t;
    }
}

int launch_tile_order(cudaStream_t s, int W, int H, Band band, char* image, const ImageLayout& IL)
{
    const uint32_t tiles = (uint32_t)((W + TILE - 1) / TILE) * (uint32_t)band.rows();
    tile_order_kernel<<<1, 1024, 0, s>>>(tiles, reinterpret_cast<const uint32_t*>(image + IL.tile_work),
                                         reinterpret_cast<uint32_t*>(image + IL.tile_order),
                                         reinterpret_cast<uint32_t*>(image + IL.sched));
    EOGS_LAUNCH_CHECK("tile_order_kernel");
    return 0;
}

int launch_blend_bwd(cudaStream_t s, int W, int H, Band band, int channels, const char* geom,
                     const GeomLayout& GL, const uint32_t* point_list, const uint8_t* masks, const char* image,
                     const ImageLayout& IL, const float* bg, const float* dL_dpix,
                     const float* dL_dinvdepth, float* grad_rec, uint32_t* queue)
{
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = band.rows();
    // persistent: as many CTAs as fit on the device at once (4 per SM), never more warps than 4 per tile (the kernel
    // splits tiles over warps when there are fewer active tiles than warps)
    const int ctas_fit = sm_count_cached() * 4;
    const int ctas_need = tiles_x * tiles_y;
    const dim3 grid(ctas_need < ctas_fit ? ctas_need : ctas_fit, 1, 1);
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
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), dL_dpix, dL_dinvdepth, grad_rec, masks,
            reinterpret_cast<const uint32_t*>(image + IL.tile_order),
            reinterpret_cast<const uint32_t*>(image + IL.sched), queue);
    };
    if (channels == 5) run(blend_bwd_kernel<5>);
    else if (channels == 3) run(blend_bwd_kernel<3>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(blend_bwd_kernel)");
    EOGS_LAUNCH_CHECK("blend_bwd_kernel");
    return 0;
}

}  // namespace eogs
