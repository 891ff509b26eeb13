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

namespace eogs {

constexpr int BWD_WARPS = 4;                 // tiles per CTA: a 2x2 block of tiles
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr int NPATCH = (TILE / PATCH_W) * (TILE / PATCH_H);   // 8

struct BwdWarpSmem {
    float4 rec[2][REC_F4][32];        // two stages of 32 packed records
    float4 pix_a[NPATCH][32];         // dL_dpixel[0..3]           of pixel (patch, lane)
    float4 pix_b[NPATCH][32];         // dL_dpixel[4], dL_dinvdepth, T_final * (bg . dL_dpixel), -
};

// 4 CTAs (16 warps) per SM measured best: 3 (142 regs) and 5 (96 regs) are both 14 % slower.
template <int C>
__global__ void __launch_bounds__(BWD_THREADS, 4)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ splat, const float* __restrict__ bg, int W, int H,
                 int tiles_x, int tiles_y, int band_row0, int band_h,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinvdepth,
                 float* __restrict__ grad_rec)
{
    constexpr int NV = 6 + C;   // mean2D.xy, conic.xyw, opacity, colours
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ BwdWarpSmem s_warp[BWD_WARPS];

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // tiles_y = tile rows of the band; brow = row inside the band, tile_y = row in the image
    const int tile_x = (int)blockIdx.x * 2 + (int)(warp & 1u), brow = (int)blockIdx.y * 2 + (int)(warp >> 1);
    if (tile_x >= tiles_x || brow >= tiles_y) return;            // whole warp leaves; no block barriers below
    const int tile_y = brow + band_row0;
    BwdWarpSmem& sm = s_warp[warp];

    const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);
    const float img_x1 = (float)(W - 1), img_y1 = (float)(H - 1);
    const uint2 range = __ldg(ranges + (size_t)brow * tiles_x + tile_x);
    const uint32_t* list = point_list + range.x;

    // ---- per-pixel state: pixel (patch p, lane) = (tile_x*16 + 8*(p&1) + (lane&7), tile_y*16 + 4*(p>>1) + (lane>>3))
    float pxf[2], pyf[4];
    pxf[0] = tx0 + (float)(lane & 7u); pxf[1] = pxf[0] + (float)PATCH_W;
#pragma unroll
    for (int r = 0; r < 4; r++) pyf[r] = ty0 + (float)(PATCH_H * r) + (float)(lane >> 3);

    float bgv[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) bgv[ch] = __ldg(bg + ch);

    int ncon[NPATCH], pmax[NPATCH];
    float T[NPATCH], accum[NPATCH];
    int nmax = 0;
#pragma unroll
    for (int p = 0; p < NPATCH; p++) {
        const int px = tile_x * TILE + PATCH_W * (p & 1) + (int)(lane & 7u);
        const int py = tile_y * TILE + PATCH_H * (p >> 1) + (int)(lane >> 3);
        const bool inside = px < W && py < H;
        const size_t pix_id = (size_t)(py - band_row0 * TILE) * W + px;     // band-compact buffers
        ncon[p] = inside ? (int)__ldg(n_contrib + pix_id) : 0;
        const float Tf = inside ? __ldg(final_T + pix_id) : 0.f;
        T[p] = Tf;
        accum[p] = 0.f;
        float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        float bg_dot_g = 0.f;
#pragma unroll
        for (int ch = 0; ch < C; ch++) {
            g[ch] = inside ? __ldg(dL_dpix + (size_t)ch * band_h * W + pix_id) : 0.f;
            bg_dot_g = fmaf(bgv[ch], g[ch], bg_dot_g);
        }
        const float g_inv = (inside && dL_dinvdepth) ? __ldg(dL_dinvdepth + pix_id) : 0.f;
        sm.pix_a[p][lane] = make_float4(g[0], g[1], g[2], g[3]);
        sm.pix_b[p][lane] = make_float4(g[4], g_inv, Tf * bg_dot_g, 0.f);
        pmax[p] = __reduce_max_sync(FULL, ncon[p]);
        nmax = max(nmax, pmax[p]);
    }
    if (nmax == 0) return;                     // entries [nmax, n) were blended by no pixel of this tile
    const int rounds = (nmax + 31) >> 5;

    // Batch b, lane l holds list position nmax-1 - (32 b + l): back to front.
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    uint32_t rid = 0, id_next = 0;
    {
        const int p0 = nmax - 1 - (int)lane;
        if (p0 >= 0) { rid = __ldg(list + p0); fetch_record(splat, rid, r0, r1, r2); }
        if (p0 - 32 >= 0) id_next = __ldg(list + p0 - 32);
    }
    const int my_slot = fold_slot<NV>(lane);
    // the butterfly leaves the mean2D sums un-scaled: d(pixel)/d(ndc) is applied once per flush
    const float slot_scale = my_slot == 0 ? 0.5f * (float)W : (my_slot == 1 ? 0.5f * (float)H : 1.f);

    int first = nmax - 1;                      // list position held by lane 0 in this batch
    for (int i = 0; i < rounds; i++, first -= 32) {
        const int pos = first - (int)lane;
        uint32_t m = pos >= 0 ? patch_mask(r0, r1, tx0, ty0, img_x1, img_y1) : 0u;
#pragma unroll
        for (int p = 0; p < NPATCH; p++)
            if (pos >= pmax[p]) m &= ~(1u << p);               // every pixel of patch p stopped before pos
        const int stage = i & 1;
        if (m) {
            sm.rec[stage][0][lane] = r0;
            sm.rec[stage][1][lane] = r1;
            sm.rec[stage][2][lane] = r2;
        }
        const uint32_t cur_rid = rid;
        uint32_t todo = __ballot_sync(FULL, m != 0u);
        __syncwarp();

        // next batch's records are in flight while this one is replayed
        if (pos - 32 >= 0) { rid = id_next; fetch_record(splat, rid, r0, r1, r2); }
        if (pos - 64 >= 0) id_next = __ldg(list + pos - 64);

        while (todo) {
            const int e = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t me = __shfl_sync(FULL, m, e);
            const int pos_e = first - e;
            const float4 ra = sm.rec[stage][0][e];
            const float4 rb = sm.rec[stage][1][e];
            const float4 rc = sm.rec[stage][2][e];
            const float col[5] = {rb.z, rb.w, rc.x, rc.y, rc.z};

            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; k++) v[k] = 0.f;
            bool any = false;

            // Two patches (the left and right half of one 4-pixel-high band) per step: two independent
            // dependency chains per lane, written branch-free — a rejected pixel runs the same
            // arithmetic with alpha = G = dL_dalpha = 0, which leaves T, accum and every accumulator
            // unchanged.  A patch whose mask bit is clear cannot be accepted (the mask is conservative),
            // so the bits only decide whether the band is visited at all.
#pragma unroll
            for (int r = 0; r < NPATCH / 2; r++) {
                if (!((me >> (2 * r)) & 3u)) continue;           // warp-uniform
                const float dy = __fsub_rn(ra.y, pyf[r]);
                const float cyy = __fmul_rn(__fmul_rn(rb.x, dy), dy);
                float dx[2], G[2], alpha[2];
                bool valid[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    // the forward's exponent, same op order (pair_power); exp through ex2.approx:
                    // gradients carry a 1e-3 bar, not the forward's bit-exact one
                    dx[h] = __fsub_rn(ra.x, pxf[h]);
                    const float quad = __fmaf_rn(dx[h], __fmul_rn(ra.z, dx[h]), cyy);
                    const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(__fmul_rn(ra.w, dx[h]), dy));
                    G[h] = fast_exp(power);
                    alpha[h] = fminf(0.99f, rb.y * G[h]);
                    // Entry at list position pos_e is blended by a pixel iff pos_e < n_contrib (backward.cu:556-558).
                    valid[h] = pos_e < ncon[2 * r + h] && !(power > 0.0f) && !(alpha[h] < 1.0f / 255.0f);
                }
                if (!__any_sync(FULL, valid[0] || valid[1])) continue;
                any = true;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int p = 2 * r + h;
                    const float4 ga = sm.pix_a[p][lane];
                    const float4 gb = sm.pix_b[p][lane];
                    const float g[5] = {ga.x, ga.y, ga.z, ga.w, gb.x};
                    const float a = valid[h] ? alpha[h] : 0.f;
                    const float Gv = valid[h] ? G[h] : 0.f;
                    const float inv_1ma = fast_rcp(1.f - a);          // exactly 1 for a rejected pixel
                    const float Tn = T[p] * inv_1ma;
                    T[p] = Tn;
                    const float w = a * Tn;
                    float cg = rc.w * gb.y;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        v[6 + ch] = fmaf(w, g[ch], v[6 + ch]);
                        cg = fmaf(col[ch], g[ch], cg);
                    }
                    // accum = (colour blended behind this entry) . dL_dpixel
                    const float behind = cg - accum[p];
                    const float dL_dalpha = valid[h] ? fmaf(-gb.z, inv_1ma, behind * Tn) : 0.f;
                    accum[p] = fmaf(a, behind, accum[p]);

                    const float dL_dG = rb.y * dL_dalpha;
                    const float gdx = Gv * dx[h], gdy = Gv * dy;
                    v[0] = fmaf(-dL_dG, fmaf(gdx, ra.z, gdy * ra.w), v[0]);
                    v[1] = fmaf(-dL_dG, fmaf(gdy, rb.x, gdx * ra.w), v[1]);
                    const float hg = -0.5f * dL_dG;
                    const float hgx = hg * gdx;
                    v[2] = fmaf(hgx, dx[h], v[2]);
                    v[3] = fmaf(hgx, dy, v[3]);
                    v[4] = fmaf(hg * gdy, dy, v[4]);
                    v[5] = fmaf(Gv, dL_dalpha, v[5]);
                }
            }
            if (any) {
                warp_transpose_reduce<NV>(v, lane);
                const uint32_t gid = __shfl_sync(FULL, cur_rid, e);
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
    const dim3 grid((tiles_x + 1) / 2, (tiles_y + 1) / 2, 1);
    auto run = [&](auto kernel) {
        kernel<<<grid, BWD_THREADS, 0, s>>>(
            reinterpret_cast<const uint2*>(image + IL.ranges), point_list,
            reinterpret_cast<const float4*>(geom + GL.splat), bg, W, H, tiles_x, tiles_y,
            band.row_begin, band.height(H), reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), dL_dpix, dL_dinvdepth, grad_rec);
    };
    if (channels == 5) run(blend_bwd_kernel<5>);
    else if (channels == 3) run(blend_bwd_kernel<3>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    EOGS_LAUNCH_CHECK("blend_bwd_kernel");
    return 0;
}

}  // namespace eogs
