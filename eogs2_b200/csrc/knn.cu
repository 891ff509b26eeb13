// distCUDA2 for sm_100a: mean squared distance of every point to its 3 nearest neighbours
// (reference: submodules/simple-knn/simple_knn.cu:187-222, spatial.cu:15-26).
//
// What the reference computes is EXACT and independent of its traversal: for point i, the three smallest
// values of  d(i,j) = fma(dz,dz, fma(dx,dx, dy*dy)),  (dx,dy,dz) = p_j - p_i,  over all j != i (by index, so
// a duplicate of p_i at another index counts with distance 0), kept sorted, then ((b0 + b1) + b2) / 3.0f
// (op order read from the SASS of the reference built for sm_100a; `updateKBest`, simple_knn.cu:118-132;
// missing neighbours keep the reference's sentinel FLT_MAX = 1E+37, :26).  Its Morton boxes (:139-185) only
// prune the search, so a different search structure gives the same bits.
//
// The reference: thread per point; every thread tests all P/1024 boxes and scans each surviving box with
// 1024 dependent gathers points[indices[i]]; two blocking D2H copies and five device allocations per call.
// Here:
//   1. bounding box (warp redux on order-preserving integer images of the floats, 6 atomics per warp)
//   2. 30-bit Morton codes + CUB radix sort of (code, index) pairs, 4 passes
//   3. gather into Morton order as float4 {x, y, z, index} and build a 3-level box hierarchy with fan-out 32:
//      leaf = 32 consecutive points (one coalesced 512-B line), group = 32 leaves, super = 32 groups
//   4. search: ONE WARP PER LEAF.  The 32 queries of a leaf are neighbours in space, so they share one
//      traversal: lanes test 32 boxes of a level at once against the leaf's own box inflated by the warp's
//      worst third-best distance (ballot); a survivor is kept only if some query's own ball reaches its box
//      (vote; a leaf that straddles a jump of the Morton curve has a huge box of its own); a surviving leaf is
//      loaded once (coalesced), staged in shared memory and read back as 16-byte broadcasts; each lane
//      updates its sorted best-3 with 5 FMNMX.
//      The own leaf and its two Morton neighbours go first so the bound is tight before the traversal.
//   Everything is enqueued on the caller's stream; no host synchronisation, no allocation.
// Pruning is exact: the box-to-box bound is computed with the same mul/fma/fma shape as the point distance,
// and IEEE rounding is monotone, so bound <= d(i,j) for every pair the box could hold (no margin needed).
// Inputs are assumed finite (NaN coordinates poison the reference's Morton codes as well).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace eogs {

constexpr float KNN_FAR = 1e37f;      // the reference's sentinel (simple_knn.cu:26)
constexpr float KNN_EMPTY = 3.0e38f;  // empty-box corner: any gap against it squares to +inf
constexpr int KNN_LEAF = 32;
constexpr int KNN_SEARCH_WARPS = 8;

struct KnnLayout {
    size_t bbox;        // u32[8]: ordered images of min xyz [0..2], max xyz [4..6]
    size_t code_a, code_b, idx_a, idx_b;    // u32[P] each (radix sort ping-pong)
    size_t pts;         // float4[leaf_slots * 32] points in Morton order, w = original index bits
    size_t leaf_box;    // float4[2 * leaf_slots]  {lo, hi}
    size_t group_box;   // float4[2 * group_slots]
    size_t super_box;   // float4[2 * n_super]
    size_t temp, temp_bytes, total;
    int n_leaf, n_group, n_super, leaf_slots, group_slots;
};

static KnnLayout knn_layout(int P) {
    KnnLayout L{};
    const size_t n = (size_t)(P > 0 ? P : 1);
    L.n_leaf = (int)((n + KNN_LEAF - 1) / KNN_LEAF);
    L.n_group = (L.n_leaf + 31) / 32;
    L.n_super = (L.n_group + 31) / 32;
    L.leaf_slots = L.n_group * 32;
    L.group_slots = L.n_super * 32;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.bbox = take(32);
    L.code_a = take(4 * n); L.code_b = take(4 * n); L.idx_a = take(4 * n); L.idx_b = take(4 * n);
    L.pts = take(16 * (size_t)L.leaf_slots * KNN_LEAF);
    L.leaf_box = take(32 * (size_t)L.leaf_slots);
    L.group_box = take(32 * (size_t)L.group_slots);
    L.super_box = take(32 * (size_t)L.n_super);
    size_t tb = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k, v, (int)n, 0, 30);
    L.temp_bytes = tb + 256;
    L.temp = take(L.temp_bytes);
    L.total = off;
    return L;
}

// order-preserving map float -> u32 (so integer min/max reductions order like the floats)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void __launch_bounds__(256)
knn_bbox_kernel(int P, const float* __restrict__ points, uint32_t* __restrict__ bbox)
{
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t o = f2ord(points[3 * (size_t)i + k]);
            lo[k] = min(lo[k], o);
            hi[k] = max(hi[k], o);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = __reduce_min_sync(0xFFFFFFFFu, lo[k]);
        hi[k] = __reduce_max_sync(0xFFFFFFFFu, hi[k]);
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&bbox[k], lo[k]);
            atomicMax(&bbox[4 + k], hi[k]);
        }
    }
}

__device__ __forceinline__ uint32_t spread3(uint32_t x) {      // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ points, const uint32_t* __restrict__ bbox,
                  uint32_t* __restrict__ codes, uint32_t* __restrict__ ids)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    // one scale for the three axes (cubic cells): leaves stay compact in space when the cloud is flat,
    // unlike the reference's per-axis normalisation (simple_knn.cu:58-62)
    const float lo0 = ord2f(bbox[0]), lo1 = ord2f(bbox[1]), lo2 = ord2f(bbox[2]);
    const float ext = fmaxf(fmaxf(ord2f(bbox[4]) - lo0, ord2f(bbox[5]) - lo1), ord2f(bbox[6]) - lo2);
    const float scale = ext > 0.f ? 1023.f / ext : 0.f;
    const float lo[3] = {lo0, lo1, lo2};
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        q[k] = (uint32_t)fminf(fmaxf((points[3 * (size_t)i + k] - lo[k]) * scale, 0.f), 1023.f);
    codes[i] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
    ids[i] = (uint32_t)i;
}

// One 1024-thread block = one group = 32 leaves; warp w gathers leaf blockIdx.x*32 + w.
__global__ void __launch_bounds__(1024)
knn_leaves_kernel(int P, const float* __restrict__ points, const uint32_t* __restrict__ ids_sorted,
                  float4* __restrict__ pts, float4* __restrict__ leaf_box, float4* __restrict__ group_box)
{
    __shared__ float s_lo[32][3], s_hi[32][3];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const size_t leaf = (size_t)blockIdx.x * 32 + warp;
    const size_t s = leaf * KNN_LEAF + lane;
    float lo[3] = {KNN_EMPTY, KNN_EMPTY, KNN_EMPTY}, hi[3] = {-KNN_EMPTY, -KNN_EMPTY, -KNN_EMPTY};
    float4 p = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000), 0.f);
    if (s < (size_t)P) {
        const uint32_t id = ids_sorted[s];
        p = make_float4(points[3 * (size_t)id], points[3 * (size_t)id + 1], points[3 * (size_t)id + 2],
                        __uint_as_float(id));
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    }
    pts[s] = p;                                     // tail of the last leaf: +inf coordinates (never a neighbour)
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
        }
    }
    if (lane == 0) {
        leaf_box[2 * leaf] = make_float4(lo[0], lo[1], lo[2], 0.f);
        leaf_box[2 * leaf + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
#pragma unroll
        for (int k = 0; k < 3; k++) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float a = s_lo[lane][k], b = s_hi[lane][k];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xFFFFFFFFu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xFFFFFFFFu, b, o));
            }
            lo[k] = a; hi[k] = b;
        }
        if (lane == 0) {
            group_box[2 * (size_t)blockIdx.x] = make_float4(lo[0], lo[1], lo[2], 0.f);
            group_box[2 * (size_t)blockIdx.x + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
    }
}

// One warp per super box: reduce its 32 group slots, filling the slots past n_group with empty boxes.
__global__ void __launch_bounds__(128)
knn_supers_kernel(int n_group, int n_super, float4* __restrict__ group_box, float4* __restrict__ super_box)
{
    const int sup = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sup >= n_super) return;
    const uint32_t lane = lane_id();
    const int g = sup * 32 + (int)lane;
    float4 lo = make_float4(KNN_EMPTY, KNN_EMPTY, KNN_EMPTY, 0.f), hi = make_float4(-KNN_EMPTY, -KNN_EMPTY, -KNN_EMPTY, 0.f);
    if (g < n_group) { lo = group_box[2 * (size_t)g]; hi = group_box[2 * (size_t)g + 1]; }
    else { group_box[2 * (size_t)g] = lo; group_box[2 * (size_t)g + 1] = hi; }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xFFFFFFFFu, lo.x, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xFFFFFFFFu, hi.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xFFFFFFFFu, lo.y, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xFFFFFFFFu, hi.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xFFFFFFFFu, lo.z, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xFFFFFFFFu, hi.z, o));
    }
    if (lane == 0) { super_box[2 * (size_t)sup] = lo; super_box[2 * (size_t)sup + 1] = hi; }
}

// Lower bound of d(i,j) over i in box Q, j in box B, in the rounding shape of the point distance.
__device__ __forceinline__ float box_gap2(const float4& qlo, const float4& qhi, const float4& blo, const float4& bhi) {
    const float gx = fmaxf(0.f, fmaxf(__fadd_rn(blo.x, -qhi.x), __fadd_rn(qlo.x, -bhi.x)));
    const float gy = fmaxf(0.f, fmaxf(__fadd_rn(blo.y, -qhi.y), __fadd_rn(qlo.y, -bhi.y)));
    const float gz = fmaxf(0.f, fmaxf(__fadd_rn(blo.z, -qhi.z), __fadd_rn(qlo.z, -bhi.z)));
    return __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
}

// Lower bound of d(i,j) for ONE query q and every j in box `idx` (same rounding argument).
__device__ __forceinline__ float point_gap2(const float4& q, const float4* __restrict__ box, size_t idx) {
    const float4 blo = box[2 * idx], bhi = box[2 * idx + 1];            // warp-uniform address: one broadcast each
    const float gx = fmaxf(0.f, fmaxf(__fadd_rn(blo.x, -q.x), __fadd_rn(q.x, -bhi.x)));
    const float gy = fmaxf(0.f, fmaxf(__fadd_rn(blo.y, -q.y), __fadd_rn(q.y, -bhi.y)));
    const float gz = fmaxf(0.f, fmaxf(__fadd_rn(blo.z, -q.z), __fadd_rn(q.z, -bhi.z)));
    return __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
}

struct Best3 { float b0, b1, b2; };

// updateKBest<3> (simple_knn.cu:118-132) as a 5-op min/max insertion into the sorted triple
__device__ __forceinline__ void insert3(Best3& b, float d) {
    const float t0 = fminf(b.b0, d), d1 = fmaxf(b.b0, d);
    const float t1 = fminf(b.b1, d1), d2 = fmaxf(b.b1, d1);
    b.b2 = fminf(b.b2, d2);
    b.b1 = t1;
    b.b0 = t0;
}

template <bool OWN>
__device__ __forceinline__ void scan_leaf(const float4* __restrict__ pts, size_t leaf, float4* stage, uint32_t lane,
                                          const float4& q, Best3& best)
{
    const float4 c = pts[leaf * KNN_LEAF + lane];
    __syncwarp();
    stage[lane] = c;
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < KNN_LEAF; j++) {
        const float4 cj = stage[j];
        const float dx = __fadd_rn(cj.x, -q.x), dy = __fadd_rn(cj.y, -q.y), dz = __fadd_rn(cj.z, -q.z);
        float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (OWN && j == (int)lane) d = __int_as_float(0x7f800000);          // `if (i == idx) continue;`
        insert3(best, d);
    }
}

__device__ __forceinline__ float warp_bound(const Best3& best, bool valid) {
    // non-negative floats order like their bit patterns
    return __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, valid ? __float_as_uint(best.b2) : 0u));
}

__global__ void __launch_bounds__(KNN_SEARCH_WARPS * 32)
knn_search_kernel(int P, int n_leaf, int n_super, const float4* __restrict__ pts,
                  const float4* __restrict__ leaf_box, const float4* __restrict__ group_box,
                  const float4* __restrict__ super_box, float* __restrict__ mean_dist2)
{
    __shared__ float4 s_stage[KNN_SEARCH_WARPS][KNN_LEAF];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int leaf = blockIdx.x * KNN_SEARCH_WARPS + (int)warp;
    if (leaf >= n_leaf) return;                         // warp-uniform
    float4* stage = s_stage[warp];
    const size_t pos = (size_t)leaf * KNN_LEAF + lane;
    const bool valid = pos < (size_t)P;
    const float4 q = pts[pos];
    const float4 qlo = leaf_box[2 * (size_t)leaf], qhi = leaf_box[2 * (size_t)leaf + 1];
    Best3 best{KNN_FAR, KNN_FAR, KNN_FAR};

    scan_leaf<true>(pts, (size_t)leaf, stage, lane, q, best);
    if (leaf > 0) scan_leaf<false>(pts, (size_t)leaf - 1, stage, lane, q, best);
    if (leaf + 1 < n_leaf) scan_leaf<false>(pts, (size_t)leaf + 1, stage, lane, q, best);
    float bound = warp_bound(best, valid);

    for (int sbase = 0; sbase < n_super; sbase += 32) {
        const int sidx = sbase + (int)lane;
        float ds = __int_as_float(0x7f800000);
        if (sidx < n_super) ds = box_gap2(qlo, qhi, super_box[2 * (size_t)sidx], super_box[2 * (size_t)sidx + 1]);
        uint32_t smask = __ballot_sync(0xFFFFFFFFu, ds <= bound);
        while (smask) {
            const int sup = sbase + __ffs(smask) - 1;
            smask &= smask - 1;
            if (!__any_sync(0xFFFFFFFFu, valid && point_gap2(q, super_box, (size_t)sup) <= best.b2)) continue;
            const size_t g = (size_t)sup * 32 + lane;
            const float dg = box_gap2(qlo, qhi, group_box[2 * g], group_box[2 * g + 1]);
            uint32_t gmask = __ballot_sync(0xFFFFFFFFu, dg <= bound);
            while (gmask) {
                const int grp = sup * 32 + __ffs(gmask) - 1;
                gmask &= gmask - 1;
                if (!__any_sync(0xFFFFFFFFu, valid && point_gap2(q, group_box, (size_t)grp) <= best.b2)) continue;
                const int l = grp * 32 + (int)lane;
                const float dl = box_gap2(qlo, qhi, leaf_box[2 * (size_t)l], leaf_box[2 * (size_t)l + 1]);
                const bool seen = l >= leaf - 1 && l <= leaf + 1;
                uint32_t lmask = __ballot_sync(0xFFFFFFFFu, !seen && dl <= bound);
                while (lmask) {
                    const size_t cand = (size_t)grp * 32 + (__ffs(lmask) - 1);
                    lmask &= lmask - 1;
                    // the leaf's box is only a prefilter (a leaf that straddles a jump of the Morton curve has a
                    // huge box): scan a candidate only if some query's own ball, at its CURRENT radius, reaches it
                    if (!__any_sync(0xFFFFFFFFu, valid && point_gap2(q, leaf_box, cand) <= best.b2)) continue;
                    scan_leaf<false>(pts, cand, stage, lane, q, best);
                    bound = warp_bound(best, valid);
                }
            }
        }
    }
    if (valid)
        mean_dist2[__float_as_uint(q.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(best.b0, best.b1), best.b2), 3.0f);
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API size_t eogs_knn_bytes(int P) { return knn_layout(P).total; }

EOGS_API int eogs_knn_dist2(eogs_stream_t stream, int P, const float* points, void* scratch, size_t scratch_bytes,
                            float* mean_dist2)
{
    if (P < 0) { set_error("bad P"); return -1; }
    if (P == 0) return 0;
    if (!points || !scratch || !mean_dist2) { set_error("null argument"); return -4; }
    const KnnLayout L = knn_layout(P);
    if (scratch_bytes < L.total) { set_error("knn scratch %zu < %zu", scratch_bytes, L.total); return -3; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(scratch);
    uint32_t* bbox = reinterpret_cast<uint32_t*>(base + L.bbox);
    EOGS_CUDA(cudaMemsetAsync(bbox, 0xFF, 16, s));
    EOGS_CUDA(cudaMemsetAsync(bbox + 4, 0x00, 16, s));
    const int blocks = (P + 255) / 256;
    knn_bbox_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, 256, 0, s>>>(P, points, bbox);
    EOGS_LAUNCH_CHECK("knn_bbox_kernel");
    cub::DoubleBuffer<uint32_t> keys(reinterpret_cast<uint32_t*>(base + L.code_a), reinterpret_cast<uint32_t*>(base + L.code_b));
    cub::DoubleBuffer<uint32_t> vals(reinterpret_cast<uint32_t*>(base + L.idx_a), reinterpret_cast<uint32_t*>(base + L.idx_b));
    knn_morton_kernel<<<blocks, 256, 0, s>>>(P, points, bbox, keys.Current(), vals.Current());
    EOGS_LAUNCH_CHECK("knn_morton_kernel");
    size_t tb = L.temp_bytes;
    EOGS_CUDA(cub::DeviceRadixSort::SortPairs(base + L.temp, tb, keys, vals, P, 0, 30, s));
    float4* pts = reinterpret_cast<float4*>(base + L.pts);
    float4* leaf_box = reinterpret_cast<float4*>(base + L.leaf_box);
    float4* group_box = reinterpret_cast<float4*>(base + L.group_box);
    float4* super_box = reinterpret_cast<float4*>(base + L.super_box);
    knn_leaves_kernel<<<L.n_group, 1024, 0, s>>>(P, points, vals.Current(), pts, leaf_box, group_box);
    EOGS_LAUNCH_CHECK("knn_leaves_kernel");
    knn_supers_kernel<<<(L.n_super * 32 + 127) / 128, 128, 0, s>>>(L.n_group, L.n_super, group_box, super_box);
    EOGS_LAUNCH_CHECK("knn_supers_kernel");
    knn_search_kernel<<<(L.n_leaf + KNN_SEARCH_WARPS - 1) / KNN_SEARCH_WARPS, KNN_SEARCH_WARPS * 32, 0, s>>>(
        P, L.n_leaf, L.n_super, pts, leaf_box, group_box, super_box, mean_dist2);
    EOGS_LAUNCH_CHECK("knn_search_kernel");
    return 0;
}

}  // extern "C"
