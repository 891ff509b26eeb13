// Backward blend: one 256-thread block per tile, one thread per pixel, back-to-front replay of
// the tile's list.  Replaces renderCUDA<5> backward (DGR/cuda_rasterizer/backward.cu:458-643).
//
// The reference issues 12 atomicAdd(float) to global memory per contributing (pixel, Gaussian)
// pair (backward.cu:598,608,631-640).  Here the 6+C per-pair terms are first summed over the
// 32 pixels of a warp with a transposing butterfly (fold): at each of the 5 shuffle levels a
// lane keeps half of its values and sends the other half, so N values cost N/2 + N/4 + ... + 1
// = 13 shuffles for N = 11 instead of 5*N = 55, and the result ends up spread over the lanes —
// one value per even lane — which then issue ONE red.global.add.f32 warp instruction into the
// Gaussian's contiguous 64-byte gradient record.  Warps in which no pixel is touched by the
// Gaussian (ballot == 0) skip everything after the alpha test.
//
// Other differences from the reference kernel
//   - the replay starts at the tile's max(n_contrib) instead of the end of the list, so list
//     entries nobody blended are never fetched (the reference stages and skips them); a warp
//     additionally never sees entries beyond its own 32 pixels' max(n_contrib);
//   - exact tile / warp-patch culling with per-warp entry lists and a register-prefetch pipeline
//     with one barrier per batch, exactly as in the forward (blend_common.cuh);
//   - the per-channel accum_rec / last_color recurrences (backward.cu:584-596) are collapsed
//     into one scalar recurrence on the dot product with dL_dpixel (same linear map);
//   - dL_dinvdepth per Gaussian is not accumulated: the reference computes it and drops it
//     (backward.cu:305-307 commented out; it never reaches a returned gradient).
//
// Bound: FP32 issue + shuffle.  Algorithmic HBM bytes: 52 B gathered + <= 8 warps x 44 B reduced per
// instance, 4*(C+1) + 8 B per pixel in.
#include "blend_common.cuh"

namespace eogs {

// Transposing butterfly level: N live values -> ceil(N/2), partner = lane ^ (1 << BIT).
template <int N, int BIT>
__device__ __forceinline__ void fold(float* v, uint32_t lane) {
    constexpr int HALF = (N + 1) / 2;
    const bool upper = (lane >> BIT) & 1u;
#pragma unroll
    for (int k = 0; k < HALF; k++) {
        const float hi = (k + HALF < N) ? v[k + HALF] : 0.f;
        const float keep = upper ? hi : v[k];
        const float send = upper ? v[k] : hi;
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << BIT);
    }
}

// After fold<N,4>, <N1,3>, <N2,2>, <N3,1> and a final xor-1 add, lane L holds the warp total of
// value slot(L) in v[0]; returns -1 for lanes that hold nothing.
template <int N>
__device__ __forceinline__ int fold_slot(uint32_t lane) {
    constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
    static_assert(N4 == 1, "fold supports up to 16 values");
    const int b1 = (lane >> 1) & 1, b2 = (lane >> 2) & 1, b3 = (lane >> 3) & 1, b4 = (lane >> 4) & 1;
    int pos = b1 * N4;
    if (pos >= N3) return -1;
    pos += b2 * N3;
    if (pos >= N2) return -1;
    pos += b3 * N2;
    if (pos >= N1) return -1;
    pos += b4 * N1;
    if (pos >= N) return -1;
    return (lane & 1u) ? -1 : pos;
}

template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float* v, uint32_t lane) {
    constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2;
    fold<N, 4>(v, lane);
    fold<N1, 3>(v, lane);
    fold<N2, 2>(v, lane);
    fold<N3, 1>(v, lane);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float fast_rcp(float x) {      // x in [0.01, 1]: no denormal handling needed
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int C>
__global__ void __launch_bounds__(BLEND_THREADS, 4)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ splat, const float* __restrict__ bg, int W, int H,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinvdepth,
                 float* __restrict__ grad_rec)
{
    constexpr int NV = 6 + C;   // mean2D.xy, conic.xyw, opacity, colours
    __shared__ BlendStage s_stage[2];
    __shared__ uint32_t s_id[2][BLEND_THREADS];
    __shared__ int s_wmax[BLEND_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t lx, ly;
    tile_pixel(tid, lx, ly);
    const uint32_t pix_x = blockIdx.x * TILE + lx, pix_y = blockIdx.y * TILE + ly;
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const size_t pix_id = (size_t)pix_y * W + pix_x;
    const float pixfx = (float)pix_x, pixfy = (float)pix_y;
    const float tx0 = (float)(blockIdx.x * TILE), ty0 = (float)(blockIdx.y * TILE);
    const float img_x1 = (float)(W - 1), img_y1 = (float)(H - 1);

    const uint2 range = __ldg(ranges + blockIdx.y * gridDim.x + blockIdx.x);
    const uint32_t* list = point_list + range.x;

    const int last_contributor = inside ? (int)__ldg(n_contrib + pix_id) : 0;
    const int wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_wmax[warp] = wmax;
    __syncthreads();
    int nmax = 0;                        // entries [nmax, n) were blended by no pixel of this tile
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; w++) nmax = max(nmax, s_wmax[w]);
    if (nmax == 0) return;
    const int rounds = (nmax + BLEND_THREADS - 1) / BLEND_THREADS;

    // Batch b, thread t holds list position nmax-1 - (b*256 + t): back to front.
    auto list_pos = [&](int batch) { return nmax - 1 - (batch * BLEND_THREADS + (int)tid); };
    // patches whose pixels all stopped before list position p never replay it
    auto mask_for = [&](int p, const float4& r0, const float4& r1) {
        uint32_t m = patch_mask(r0, r1, tx0, ty0, img_x1, img_y1);
#pragma unroll
        for (int w = 0; w < BLEND_WARPS; w++)
            if (p >= s_wmax[w]) m &= ~(1u << w);
        return m;
    };

    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    uint32_t rid = 0, id_next = 0;
    {   // prologue: batch 0 staged, ids of batch 1 in registers
        const int p = list_pos(0);
        if (p >= 0) { rid = __ldg(list + p); fetch_record(splat, rid, r0, r1, r2); }
        if (list_pos(1) >= 0) id_next = __ldg(list + list_pos(1));
        const uint32_t m = p >= 0 ? mask_for(p, r0, r1) : 0u;
        if (m) s_id[0][tid] = rid;
        stage_entry(s_stage[0], tid, m, r0, r1, r2);
    }

    const float T_final = inside ? __ldg(final_T + pix_id) : 0.f;
    float T = T_final;
    float g[C];
    float bg_dot_g = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
        g[ch] = inside ? __ldg(dL_dpix + (size_t)ch * H * W + pix_id) : 0.f;
        bg_dot_g = fmaf(__ldg(bg + ch), g[ch], bg_dot_g);
    }
    const float g_inv = (inside && dL_dinvdepth) ? __ldg(dL_dinvdepth + pix_id) : 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    float last_alpha = 0.f, last_cg = 0.f, accum_rec = 0.f;
    const int my_slot = fold_slot<NV>(lane);

    for (int i = 0; i < rounds; i++) {
        __syncthreads();                 // stage i&1 complete; the other stage is free again
        const bool more = i + 1 < rounds;
        const int p_next = list_pos(i + 1);
        const bool have_next = more && p_next >= 0;
        if (have_next) { rid = id_next; fetch_record(splat, rid, r0, r1, r2); }   // in flight during the replay below
        if (list_pos(i + 2) >= 0) id_next = __ldg(list + list_pos(i + 2));

        const int stage = i & 1;
        const BlendStage& st = s_stage[stage];
        const int first = nmax - 1 - i * BLEND_THREADS;            // list position of batch slot 0
        for (int seg = 0; seg < BLEND_WARPS; seg++) {
            const int cnt = st.cnt[warp][seg];
            for (int j = 0; j < cnt; j++) {
                const uint32_t e = st.list[warp][seg][j];
                // Entry at list position p is blended by this pixel iff p < n_contrib (backward.cu:556-558).
                const bool active = (first - (int)e) < last_contributor;
                const float4 ra = st.rec[0][e];
                const float4 rb = st.rec[1][e];
                float dx, dy;
                const float power = pair_power(ra, rb, pixfx, pixfy, dx, dy);
                const float G = expf(power);
                const float alpha = fminf(0.99f, __fmul_rn(rb.y, G));
                const bool valid = active && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                if (!__any_sync(0xffffffffu, valid)) continue;

                float v[NV];
#pragma unroll
                for (int k = 0; k < NV; k++) v[k] = 0.f;
                if (valid) {
                    const float4 rc = st.rec[2][e];
                    const float col[5] = {rb.z, rb.w, rc.x, rc.y, rc.z};
                    const float inv_1ma = fast_rcp(1.f - alpha);
                    T *= inv_1ma;
                    const float w = alpha * T;
                    float cg = rc.w * g_inv;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        v[6 + ch] = w * g[ch];
                        cg = fmaf(col[ch], g[ch], cg);
                    }
                    accum_rec = fmaf(last_alpha, last_cg, (1.f - last_alpha) * accum_rec);
                    last_cg = cg;
                    last_alpha = alpha;
                    float dL_dalpha = (cg - accum_rec) * T;
                    dL_dalpha = fmaf(-T_final * inv_1ma, bg_dot_g, dL_dalpha);

                    const float dL_dG = rb.y * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * ra.z - gdy * ra.w;
                    const float dG_ddely = -gdy * rb.x - gdx * ra.w;
                    v[0] = dL_dG * dG_ddelx * ddelx_dx;
                    v[1] = dL_dG * dG_ddely * ddely_dy;
                    v[2] = -0.5f * gdx * dx * dL_dG;
                    v[3] = -0.5f * gdx * dy * dL_dG;
                    v[4] = -0.5f * gdy * dy * dL_dG;
                    v[5] = G * dL_dalpha;
                }
                warp_transpose_reduce<NV>(v, lane);
                if (my_slot >= 0) atomicAdd(grad_rec + (size_t)s_id[stage][e] * GRAD_STRIDE + my_slot, v[0]);
            }
        }
        if (more) {
            const uint32_t m = have_next ? mask_for(p_next, r0, r1) : 0u;
            if (m) s_id[(i + 1) & 1][tid] = rid;
            stage_entry(s_stage[(i + 1) & 1], tid, m, r0, r1, r2);
        }
    }
}

int launch_blend_bwd(cudaStream_t s, int W, int H, int channels, const char* geom,
                     const GeomLayout& GL, const uint32_t* point_list, const char* image,
                     const ImageLayout& IL, const float* bg, const float* dL_dpix,
                     const float* dL_dinvdepth, float* grad_rec)
{
    const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
    auto run = [&](auto kernel) {
        kernel<<<grid, BLEND_THREADS, 0, s>>>(
            reinterpret_cast<const uint2*>(image + IL.ranges), point_list,
            reinterpret_cast<const float4*>(geom + GL.splat), bg, W, H,
            reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), dL_dpix, dL_dinvdepth, grad_rec);
    };
    if (channels == 5) run(blend_bwd_kernel<5>);
    else if (channels == 3) run(blend_bwd_kernel<3>);
    else { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    EOGS_LAUNCH_CHECK("blend_bwd_kernel");
    return 0;
}

}  // namespace eogs
