// Binning: depth order of Gaussians, then the per-tile depth-sorted Gaussian lists and the tile ranges.
// Replaces cub InclusiveSum + duplicateWithKeys + 64-bit cub SortPairs + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:70-138, 280-321).
//
// The reference materialises I = sum(tiles touched) (key, value) instances and radix-sorts them on a 64-bit
// (tile | depth bits) key: ceil((32+bit)/8) = 6-7 passes over 12-byte pairs.  The order it produces is fully
// determined — (tile, depth bits, Gaussian id), because the LSD sort is stable and emission is id-ascending — so the
// identical list can be built without ever sorting instances:
//   1. sort the P Gaussians once by (depth bits, id)                    (P << I; 32-bit keys, library radix sort)
//   2. a tile's list is "the Gaussians whose tile rectangle covers it, in that order".  A rectangle is a set of ROW
//      RUNS [x0, x1) x {y}; the lists are built by two stable counting passes that only ever touch runs and ids:
//        rows:     each chunk of 1024 depth-ordered Gaussians counts its runs per tile row (difference array +
//                  prefix sum in shared memory), a column scan over chunks gives every (chunk, row) its base, and
//                  the chunk writes its runs {id, x0, x1} into the row's run list IN DEPTH ORDER — the rank of a
//                  Gaussian inside (chunk, row) is a popcount over a 1024-bit occupancy bitmap of that row;
//        columns:  the same three steps on sub-chunks of 1024 runs of one row, over tile columns: counts per
//                  (sub-chunk, tile), scan over sub-chunks (whose per-tile totals ARE the tile list lengths, i.e. the
//                  reference's ranges after one scan over tiles), and the scatter of the ids into point_list.
// No instance keys exist at any point; eogs_export_state rebuilds the reference's 64-bit keys from ranges, point_list
// and the depths for the parity tests, and the lists / ranges are bit-identical to the reference's.
//
// Algorithmic bytes: 12 B/Gaussian read twice, 8 B/run written once and read twice, 4 B/instance written once —
// about 11 B/instance on the bench scene (3.9 runs of 4 tiles per Gaussian) against 32 B/instance for emission +
// a 16-bit two-pass radix sort of the instances (the previous version) and 176-200 B/instance in the reference.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace eogs {

size_t sort_temp_bound(size_t n) { return (size_t(1) << 20) + n; }

int sm_count_cached() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n; cached_dev = dev;
    }
    return cached;
}

GeomLayout geom_layout(int P) {
    GeomLayout L;
    const size_t n = (size_t)(P > 0 ? P : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.splat = take(n * REC_F4 * sizeof(float4));
    L.cut = take(n * 4);
    L.depth = take(n * 4);
    L.rect = take(n * 8);
    L.tiles = take(n * 4);
    L.key_in = take(n * 4);
    L.key_out = take(n * 4);
    L.id_in = take(n * 4);
    L.order = take(n * 4);
    L.temp_bytes = sort_temp_bound(n);
    L.temp = take(L.temp_bytes);
    L.total = off;
    return L;
}

Band full_band(int H) { return Band{0, (H + TILE - 1) / TILE}; }

int check_band(int H, Band band) {
    const int grid_y = (H + TILE - 1) / TILE;
    if (band.row_begin < 0 || band.row_end > grid_y || band.row_begin >= band.row_end) {
        set_error("bad band: tile rows [%d, %d) of %d", band.row_begin, band.row_end, grid_y);
        return -1;
    }
    return 0;
}

ImageLayout image_layout(int W, int H, Band band) {
    ImageLayout L;
    const size_t n = (size_t)W * (size_t)band.height(H);
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * (size_t)band.rows();
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.final_T = take(n * 4);
    L.n_contrib = take(n * 4);
    L.ranges = take(tiles * 8);
    L.tile_work = take(tiles * 4);
    L.tile_order = take(tiles * 4);
    L.sched = take(16);
    L.total = off;
    return L;
}

constexpr int BIN_CH = 1024;                 // Gaussians per row chunk = runs per column sub-chunk = bits per bitmap row
constexpr int BIN_WORDS = BIN_CH / 32;
constexpr int BIN_THREADS = 256;
constexpr int BIN_IPT = BIN_CH / BIN_THREADS; // items per thread
constexpr int BIN_WIN = 1024;                // tile rows (columns) per shared-memory window
constexpr int BIN_ROW = BIN_WORDS + 1;       // bitmap row stride in words: rows land in distinct banks
constexpr int BIN_BWIN = 256;                // bins per bitmap window of the scatter kernels (bounds their shared memory)
constexpr int BIN_PRE = 2 * (BIN_WORDS / 2 + 1);   // halves per row of the word-prefix table (17 words: odd stride)

BinningLayout binning_layout(int W, int H, uint32_t I) {
    BinningLayout L;
    const size_t n = (size_t)(I > 0 ? I : 1);
    const size_t gx = (size_t)((W + TILE - 1) / TILE), gy = (size_t)((H + TILE - 1) / TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    // every visible Gaussian owns at least one instance and every run at least one: I bounds both counts
    L.chunk_cap = n / BIN_CH + 1;
    L.sub_cap = n / BIN_CH + gy + 1;
    L.row_count = take(L.chunk_cap * gy * 4);
    L.row_total = take((gy + 1) * 4);
    L.row_start = take((gy + 1) * 4);
    L.sub_first = take((gy + 1) * 4);
    L.runs = take(n * 8);
    L.col_count = take(L.sub_cap * gx * 4);
    L.total = off;
    return L;
}

// ---- stage 1b: depth order -------------------------------------------------------
static int depth_order_impl(cudaStream_t s, int P, uint32_t* key_in, uint32_t* key_out, uint32_t* order, uint32_t* id_in,
                            void* temp, size_t temp_bytes)
{
    // (depth bits, id): keys are non-negative floats, so their bit patterns order like the
    // values; the radix sort is stable and ids come in ascending, which yields the tie order.
    size_t need = 0;
    cub::DoubleBuffer<uint32_t> keys(key_in, key_out);
    // `order` holds 0..P-1 (written by the preprocess kernel): 4 passes land back in `order`
    cub::DoubleBuffer<uint32_t> vals(order, id_in);
    EOGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, vals, P, 0, 32, s));
    if (need > temp_bytes) { set_error("depth sort temp %zu > %zu", need, temp_bytes); return -3; }
    need = temp_bytes;
    EOGS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, keys, vals, P, 0, 32, s));
    if (vals.Current() != order)
        EOGS_CUDA(cudaMemcpyAsync(order, vals.Current(), (size_t)P * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int launch_depth_order(cudaStream_t s, int P, char* geom, const GeomLayout& L,
                       eogs_forward_info* info_dev)
{
    (void)info_dev;      // num_instances was already published by the preprocess kernel
    const int rc = depth_order_impl(s, P, reinterpret_cast<uint32_t*>(geom + L.key_in), reinterpret_cast<uint32_t*>(geom + L.key_out),
                                    reinterpret_cast<uint32_t*>(geom + L.order), reinterpret_cast<uint32_t*>(geom + L.id_in),
                                    geom + L.temp, L.temp_bytes);
    prof_mark(s, ST_DEPTH_SORT);
    return rc;
}

__global__ void iota_kernel(int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

size_t debug_depth_order_bytes(int P) {
    const size_t n = align_up((size_t)(P > 0 ? P : 1) * 4, 256);
    return 3 * n + sort_temp_bound((size_t)(P > 0 ? P : 1));
}

int debug_depth_order(cudaStream_t s, int P, const uint32_t* keys, uint32_t* order, void* scratch)
{
    if (P <= 0) return 0;
    const size_t n = align_up((size_t)P * 4, 256);
    char* base = static_cast<char*>(scratch);
    uint32_t* key_in = reinterpret_cast<uint32_t*>(base);
    uint32_t* key_out = reinterpret_cast<uint32_t*>(base + n);
    uint32_t* id_in = reinterpret_cast<uint32_t*>(base + 2 * n);
    EOGS_CUDA(cudaMemcpyAsync(key_in, keys, (size_t)P * 4, cudaMemcpyDeviceToDevice, s));
    iota_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, order);
    EOGS_LAUNCH_CHECK("iota_kernel");
    return depth_order_impl(s, P, key_in, key_out, order, id_in, base + 3 * n, sort_temp_bound((size_t)P));
}

// ---- stage 2: tile lists ------------------------------------------------------------------------
// Shared pieces of the two counting passes.  An ITEM is a half-open interval [lo, hi) of bins plus a payload:
//   rows pass:     item = depth-ordered Gaussian, bins = tile rows of the band,    payload = {id, x0 | x1 << 16}
//   columns pass:  item = run of one tile row,    bins = tile columns,             payload = id

// exclusive scan of one value per thread over a block of NT threads; returns the exclusive prefix, total in `total`
template <int NT>
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t w = lane < NT / 32 ? s_warp[lane] : 0u, wincl = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d);
        if (lane >= (uint32_t)d) wincl += o;
    }
    total = __shfl_sync(0xffffffffu, wincl, NT / 32 - 1);
    const uint32_t wbase = __shfl_sync(0xffffffffu, wincl - w, wid);
    __syncthreads();                                   // s_warp may be reused by the caller
    return wbase + incl - v;
}

// Which column sub-chunk is `s`: row y with sub_first[y] <= s < sub_first[y + 1], runs [beg, end) of that row.
// sub_first (one entry per tile row) is staged in shared memory once per block, so the binary search costs one global
// round trip instead of log2(rows) dependent ones.
constexpr int BIN_SF = 1024;                       // rows whose sub_first fits the shared-memory copy
__device__ __forceinline__ const uint32_t* stage_sub_first(int gy, const uint32_t* __restrict__ sub_first, uint32_t* s_sf) {
    if (gy > BIN_SF) return sub_first;
    for (int i = threadIdx.x; i <= gy; i += blockDim.x) s_sf[i] = __ldg(sub_first + i);
    __syncthreads();
    return s_sf;
}
__device__ __forceinline__ void locate_sub(uint32_t s, int gy, const uint32_t* sf, const uint32_t* __restrict__ row_start,
                                           const uint32_t* __restrict__ row_total, int& y, uint32_t& beg, uint32_t& end) {
    int lo = 0, hi = gy;                               // largest y with sub_first[y] <= s
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sf[mid] <= s) lo = mid; else hi = mid;
    }
    y = lo;
    const uint32_t rs = __ldg(row_start + y);
    beg = rs + (s - sf[y]) * BIN_CH;
    end = min(beg + BIN_CH, rs + __ldg(row_total + y));
}

// Counting: out[bin] = number of the block's items whose interval covers bin: +1 / -1 into a shared difference array
// (two shared-memory atomics per ITEM), then a prefix sum.
__device__ __forceinline__ void count_cover(const int (&lo)[BIN_IPT], const int (&hi)[BIN_IPT], int nbins,
                                            uint32_t* __restrict__ out, int* s_diff, uint32_t* s_warp) {
    for (int w0 = 0; w0 < nbins; w0 += BIN_WIN) {
        const int wn = min(BIN_WIN, nbins - w0);
        for (int i = threadIdx.x; i <= BIN_WIN; i += BIN_THREADS) s_diff[i] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BIN_IPT; k++) {
            const int a = max(lo[k], w0) - w0, b = min(hi[k], w0 + wn) - w0;
            if (a < b) { atomicAdd(&s_diff[a], 1); atomicAdd(&s_diff[b], -1); }
        }
        __syncthreads();
        // inclusive prefix over the window: thread t owns entries 4t .. 4t+3
        int d[BIN_IPT];
        uint32_t sum = 0u;
#pragma unroll
        for (int k = 0; k < BIN_IPT; k++) { d[k] = s_diff[BIN_IPT * threadIdx.x + k]; sum += (uint32_t)d[k]; }
        uint32_t tot;
        uint32_t run = block_scan_excl<BIN_THREADS>(sum, s_warp, tot);
#pragma unroll
        for (int k = 0; k < BIN_IPT; k++) {
            run += (uint32_t)d[k];
            const int i = BIN_IPT * (int)threadIdx.x + k;
            if (i < wn) out[w0 + i] = run;
        }
        __syncthreads();
    }
}

// Scatter: the block's items, in order, to out[base(bin) + rank]; rank = number of earlier items of the block that
// cover the same bin = popcount over the bin's occupancy bitmap (one bit per item, one 32-item word per (bin, batch)).
//   build    one shared-memory atomicOr per (item, bin).  (Building the words by 32x32 warp bit-transposes of the
//            items' interval masks instead, fully or for half of the batches, was measured slower: 106 / 91 vs 76 us
//            for the column scatter of the bench scene — the kernel is issue-bound, not LSU-bound.)
//   prefix   one thread per bin walks its 32 words: exclusive popcount prefix into a 16-bit table (a warp scan per bin
//            cost 34 instructions per bin and was 20 % of the kernel, ncu r2k)
//   emit     one warp per bin, lane = word: every set bit is one output, written at base + prefix + rank-in-word,
//            so a warp's stores are consecutive.
template <typename Payload, typename BaseFn>
__device__ __forceinline__ void scatter_cover(int n_local, const uint32_t* s_lohi, const Payload* s_pay, int nbins,
                                              int win, uint32_t* s_bits, uint16_t* s_pre, uint32_t* s_base,
                                              Payload* __restrict__ out, BaseFn base_of) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int w0 = 0; w0 < nbins; w0 += win) {                             // win = min(nbins, BIN_BWIN): the shared-memory window
        const int wn = min(win, nbins - w0);
        for (int i = threadIdx.x; i < wn * BIN_ROW; i += BIN_THREADS) s_bits[i] = 0u;
        for (int i = threadIdx.x; i < wn; i += BIN_THREADS) s_base[i] = base_of(w0 + i);      // coalesced, one round trip
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BIN_IPT; k++) {
            const int j = (int)threadIdx.x + k * BIN_THREADS;             // item j: word j >> 5, bit j & 31
            if (j >= n_local) continue;
            const uint32_t lh = s_lohi[j];
            const int a = max((int)(lh & 0xFFFFu), w0), b = min((int)(lh >> 16), w0 + wn);
            for (int bin = a; bin < b; bin++) atomicOr(&s_bits[(bin - w0) * BIN_ROW + (j >> 5)], 1u << (j & 31));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < wn; i += BIN_THREADS) {
            uint32_t acc = 0u;
#pragma unroll 8
            for (int w = 0; w < BIN_WORDS; w++) {
                s_pre[i * BIN_PRE + w] = (uint16_t)acc;
                acc += (uint32_t)__popc(s_bits[i * BIN_ROW + w]);
            }
        }
        __syncthreads();
        for (int r = (int)warp; r < wn; r += BIN_THREADS / 32) {
            uint32_t bits = s_bits[r * BIN_ROW + lane];
            uint32_t pos = s_base[r] + (uint32_t)s_pre[r * BIN_PRE + lane];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                out[pos++] = s_pay[lane * 32 + b];
            }
        }
        __syncthreads();
    }
}

// rows, step 1: per (chunk of 1024 depth-ordered Gaussians, tile row) the number of runs
__global__ void __launch_bounds__(BIN_THREADS)
bin_count_rows_kernel(int n_items, int gy, int band_y0, const uint32_t* __restrict__ order,
                      const uint2* __restrict__ rect, uint32_t* __restrict__ row_count)
{
    __shared__ int s_diff[BIN_WIN + 1];
    __shared__ uint32_t s_warp[32];
    int lo[BIN_IPT], hi[BIN_IPT];
#pragma unroll
    for (int k = 0; k < BIN_IPT; k++) {
        const int j = (int)blockIdx.x * BIN_CH + (int)threadIdx.x + k * BIN_THREADS;
        lo[k] = hi[k] = 0;
        if (j < n_items) {
            const uint2 r = __ldg(rect + __ldg(order + j));
            lo[k] = (int)(r.x >> 16) - band_y0;
            hi[k] = (int)(r.y >> 16) - band_y0;
            if ((r.y & 0xFFFFu) <= (r.x & 0xFFFFu)) hi[k] = lo[k];      // culled: empty rectangle
        }
    }
    count_cover(lo, hi, gy, row_count + (size_t)blockIdx.x * gy, s_diff, s_warp);
}

// rows, step 2: exclusive scan over chunks of every row's counts -> each (chunk, row)'s base.  One block per tile row;
// thread t owns a contiguous slice of chunks, so all loads of the block are independent (two memory round trips).
constexpr int BIN_SCAN_K = 8;                      // chunks per thread and sweep: 256 x 8 = 2048 chunks (2 M Gaussians)
__global__ void __launch_bounds__(256)
bin_scan_rows_kernel(int n_chunks, int gy, uint32_t* __restrict__ row_count, uint32_t* __restrict__ row_total)
{
    __shared__ uint32_t s_warp[32];
    const int y = (int)blockIdx.x;
    uint32_t carry = 0u;
    for (int c0 = 0; c0 < n_chunks; c0 += 256 * BIN_SCAN_K) {
        uint32_t v[BIN_SCAN_K], sum = 0u;
#pragma unroll
        for (int k = 0; k < BIN_SCAN_K; k++) {
            const int c = c0 + (int)threadIdx.x * BIN_SCAN_K + k;
            v[k] = c < n_chunks ? row_count[(size_t)c * gy + y] : 0u;
            sum += v[k];
        }
        uint32_t tot;
        uint32_t run = carry + block_scan_excl<256>(sum, s_warp, tot);
#pragma unroll
        for (int k = 0; k < BIN_SCAN_K; k++) {
            const int c = c0 + (int)threadIdx.x * BIN_SCAN_K + k;
            if (c < n_chunks) row_count[(size_t)c * gy + y] = run;
            run += v[k];
        }
        carry += tot;
    }
    if (threadIdx.x == 0) row_total[y] = carry;
}

// rows, step 3 (one block): where each row's run list starts, and which column sub-chunks (1024 runs) it is cut into
__global__ void __launch_bounds__(1024)
bin_row_offsets_kernel(int gy, const uint32_t* __restrict__ row_total, uint32_t* __restrict__ row_start,
                       uint32_t* __restrict__ sub_first)
{
    __shared__ uint32_t s_warp[32];
    uint32_t run_carry = 0u, sub_carry = 0u;
    for (int y0 = 0; y0 < gy; y0 += 1024) {
        const int y = y0 + (int)threadIdx.x;
        const uint32_t t = y < gy ? row_total[y] : 0u;
        uint32_t tot_runs, tot_subs;
        const uint32_t e_runs = block_scan_excl<1024>(t, s_warp, tot_runs);
        const uint32_t e_subs = block_scan_excl<1024>((t + BIN_CH - 1) / BIN_CH, s_warp, tot_subs);
        if (y < gy) { row_start[y] = run_carry + e_runs; sub_first[y] = sub_carry + e_subs; }
        run_carry += tot_runs; sub_carry += tot_subs;
    }
    if (threadIdx.x == 0) { row_start[gy] = run_carry; sub_first[gy] = sub_carry; }
}

// rows, step 4: every chunk writes its runs {id, x0 | x1 << 16} into the rows' run lists, in depth order
__global__ void __launch_bounds__(BIN_THREADS)
bin_scatter_rows_kernel(int n_items, int gy, int band_y0, const uint32_t* __restrict__ order,
                        const uint2* __restrict__ rect, const uint32_t* __restrict__ row_base,
                        const uint32_t* __restrict__ row_start, uint2* __restrict__ runs)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t* s_lohi = s_dyn;                                     // y0 | y1 << 16 (band-relative)
    uint2* s_pay = reinterpret_cast<uint2*>(s_dyn + BIN_CH);      // {id, x0 | x1 << 16}
    const int win = min(gy, BIN_BWIN);
    uint32_t* s_base = s_dyn + 3 * BIN_CH;                        // per bin of the window: where this chunk's runs go
    uint32_t* s_bits = s_dyn + 3 * BIN_CH + win;
    uint16_t* s_pre = reinterpret_cast<uint16_t*>(s_bits + win * BIN_ROW);
    const int base = (int)blockIdx.x * BIN_CH;
    const int n_local = min(BIN_CH, n_items - base);
    for (int j = (int)threadIdx.x; j < BIN_CH; j += BIN_THREADS) {
        uint32_t lh = 0u;
        uint2 pay = make_uint2(0u, 0u);
        if (j < n_local) {
            const uint32_t g = __ldg(order + base + j);
            const uint2 r = __ldg(rect + g);
            const uint32_t x0 = r.x & 0xFFFFu, x1 = r.y & 0xFFFFu;
            if (x1 > x0) lh = ((r.x >> 16) - (uint32_t)band_y0) | (((r.y >> 16) - (uint32_t)band_y0) << 16);
            pay = make_uint2(g, x0 | (x1 << 16));
        }
        s_lohi[j] = lh;
        s_pay[j] = pay;
    }
    __syncthreads();
    const uint32_t* my_base = row_base + (size_t)blockIdx.x * gy;
    scatter_cover<uint2>(n_local, s_lohi, s_pay, gy, win, s_bits, s_pre, s_base, runs,
                         [&](int y) { return __ldg(row_start + y) + __ldg(my_base + y); });
}

// columns, step 1: per (sub-chunk of 1024 runs of one row, tile column) the number of runs covering the tile.
// Persistent blocks stride over the sub-chunks (their number is only known on the device).
__global__ void __launch_bounds__(BIN_THREADS)
bin_count_cols_kernel(int gy, int gx, const uint32_t* __restrict__ sub_first, const uint32_t* __restrict__ row_start,
                      const uint32_t* __restrict__ row_total, const uint2* __restrict__ runs,
                      uint32_t* __restrict__ col_count)
{
    __shared__ int s_diff[BIN_WIN + 1];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_sf[BIN_SF + 1];
    const uint32_t* sf = stage_sub_first(gy, sub_first, s_sf);
    const uint32_t n_subs = sf[gy];
    for (uint32_t s = blockIdx.x; s < n_subs; s += gridDim.x) {
        int y; uint32_t beg, end;
        locate_sub(s, gy, sf, row_start, row_total, y, beg, end);
        int lo[BIN_IPT], hi[BIN_IPT];
#pragma unroll
        for (int k = 0; k < BIN_IPT; k++) {
            const uint32_t j = beg + threadIdx.x + (uint32_t)k * BIN_THREADS;
            lo[k] = hi[k] = 0;
            if (j < end) {
                const uint32_t xx = __ldg(&runs[j].y);
                lo[k] = (int)(xx & 0xFFFFu);
                hi[k] = (int)(xx >> 16);
            }
        }
        count_cover(lo, hi, gx, col_count + (size_t)s * gx, s_diff, s_warp);
    }
}

// columns, step 2: exclusive scan over the sub-chunks of a row, per tile (one warp per tile, lanes over sub-chunks:
// a row has ~30 at the bench size); the total is the tile's list length, parked in ranges[tile].y for the scan over tiles
__global__ void __launch_bounds__(256)
bin_scan_cols_kernel(int gy, int gx, const uint32_t* __restrict__ sub_first, uint32_t* __restrict__ col_count,
                     uint2* __restrict__ ranges)
{
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= (uint32_t)gy * (uint32_t)gx) return;
    const uint32_t y = t / (uint32_t)gx, x = t - y * (uint32_t)gx;
    const uint32_t s0 = __ldg(sub_first + y), s1 = __ldg(sub_first + y + 1);
    uint32_t running = 0u;
    for (uint32_t sb = s0; sb < s1; sb += 32u) {
        const uint32_t s = sb + lane;
        const size_t i = (size_t)s * gx + x;
        const uint32_t v = s < s1 ? col_count[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        if (s < s1) col_count[i] = running + incl - v;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) ranges[t] = make_uint2(0u, running);
}

// columns, step 3 (one block): scan over tiles -> ranges[tile] = [start, end) of its list (identifyTileRanges,
// rasterizer_impl.cu:116-138: empty tiles stay (0, 0))
__global__ void __launch_bounds__(1024)
bin_tile_ranges_kernel(uint32_t tiles, uint2* __restrict__ ranges)
{
    __shared__ uint32_t s_warp[32];
    constexpr int K = 16;                              // slabs of 1024 consecutive tiles per sweep (a 2048^2 image = one sweep)
    uint32_t carry = 0u;
    for (uint32_t t0 = 0; t0 < tiles; t0 += 1024u * K) {
        uint32_t c[K];
#pragma unroll
        for (int k = 0; k < K; k++) {                  // coalesced, all loads in flight before the first scan
            const uint32_t t = t0 + 1024u * k + threadIdx.x;
            c[k] = t < tiles ? ranges[t].y : 0u;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t t = t0 + 1024u * k + threadIdx.x;
            if (t0 + 1024u * k >= tiles) break;        // uniform
            uint32_t tot;
            const uint32_t start = carry + block_scan_excl<1024>(c[k], s_warp, tot);
            if (t < tiles) ranges[t] = c[k] ? make_uint2(start, start + c[k]) : make_uint2(0u, 0u);
            carry += tot;
        }
    }
}

// columns, step 4: every sub-chunk writes the ids of its runs into the tiles' lists, in depth order
__global__ void __launch_bounds__(BIN_THREADS)
bin_scatter_cols_kernel(int gy, int gx, const uint32_t* __restrict__ sub_first, const uint32_t* __restrict__ row_start,
                        const uint32_t* __restrict__ row_total, const uint2* __restrict__ runs,
                        const uint32_t* __restrict__ col_base, const uint2* __restrict__ ranges,
                        uint32_t* __restrict__ point_list)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t* s_lohi = s_dyn;                                     // x0 | x1 << 16
    uint32_t* s_pay = s_dyn + BIN_CH;                             // Gaussian id
    uint32_t* s_base = s_dyn + 2 * BIN_CH;                        // per tile of the window: where this sub-chunk's ids go
    const int win = min(gx, BIN_BWIN), nsf = min(gy, BIN_SF) + 1;
    uint32_t* s_sf = s_dyn + 2 * BIN_CH + win;                    // shared copy of sub_first
    uint32_t* s_bits = s_dyn + 2 * BIN_CH + win + nsf;
    uint16_t* s_pre = reinterpret_cast<uint16_t*>(s_bits + win * BIN_ROW);
    const uint32_t* sf = stage_sub_first(gy, sub_first, s_sf);
    const uint32_t n_subs = sf[gy];
    for (uint32_t s = blockIdx.x; s < n_subs; s += gridDim.x) {
        int y; uint32_t beg, end;
        locate_sub(s, gy, sf, row_start, row_total, y, beg, end);
        const int n_local = (int)(end - beg);
        for (int j = (int)threadIdx.x; j < BIN_CH; j += BIN_THREADS) {
            uint2 r = make_uint2(0u, 0u);
            if (j < n_local) r = __ldg(runs + beg + j);
            s_pay[j] = r.x;
            s_lohi[j] = r.y;
        }
        __syncthreads();
        const uint32_t* my_base = col_base + (size_t)s * gx;
        const uint2* row_ranges = ranges + (size_t)y * gx;
        scatter_cover<uint32_t>(n_local, s_lohi, s_pay, gx, win, s_bits, s_pre, s_base, point_list,
                                [&](int x) { return __ldg(&row_ranges[x].x) + __ldg(my_base + x); });
    }
}

int launch_binning(cudaStream_t s, int P, int W, int H, Band band, uint32_t I, const char* geom,
                   const GeomLayout& GL, uint32_t* point_list, char* binning,
                   const BinningLayout& BL, char* image, const ImageLayout& IL)
{
    const int gx = (W + TILE - 1) / TILE, gy = band.rows();
    const uint32_t tiles = (uint32_t)gx * (uint32_t)gy;
    uint2* ranges = reinterpret_cast<uint2*>(image + IL.ranges);
    if (I == 0) {
        EOGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s));
        return 0;
    }
    if (I > 0x7FFFFFFFu) {
        set_error("%u (Gaussian, tile) instances exceed 2^31-1: render the view in tile bands (eogs_*_band)", I);
        return -3;
    }
    const uint32_t* order = reinterpret_cast<const uint32_t*>(geom + GL.order);
    const uint2* rect = reinterpret_cast<const uint2*>(geom + GL.rect);
    uint32_t* row_count = reinterpret_cast<uint32_t*>(binning + BL.row_count);
    uint32_t* row_total = reinterpret_cast<uint32_t*>(binning + BL.row_total);
    uint32_t* row_start = reinterpret_cast<uint32_t*>(binning + BL.row_start);
    uint32_t* sub_first = reinterpret_cast<uint32_t*>(binning + BL.sub_first);
    uint2* runs = reinterpret_cast<uint2*>(binning + BL.runs);
    uint32_t* col_count = reinterpret_cast<uint32_t*>(binning + BL.col_count);

    // Gaussians without a tile in this band sort behind every contributing one (key 0xFFFFFFFF, preprocess.cu) and
    // every contributing one owns >= 1 instance: the first min(P, I) entries of the depth order hold all the work.
    const int n_items = (int)min((uint32_t)P, I);
    const int n_chunks = (n_items + BIN_CH - 1) / BIN_CH;
    const uint32_t n_subs = I / BIN_CH + (uint32_t)gy + 1u;      // >= sum over rows of ceil(runs / 1024)
    if ((size_t)n_chunks > BL.chunk_cap || (size_t)n_subs > BL.sub_cap) { set_error("binning scratch too small"); return -3; }

    // item arrays | bases | (sub_first copy) | bitmap window | 16-bit word-prefix table
    const int win_y = min(gy, BIN_BWIN), win_x = min(gx, BIN_BWIN);
    const size_t smem_rows = (size_t)(3 * BIN_CH + win_y + win_y * BIN_ROW) * 4 + (size_t)win_y * BIN_PRE * 2;
    const size_t smem_cols = (size_t)(2 * BIN_CH + win_x + min(gy, BIN_SF) + 1 + win_x * BIN_ROW) * 4 + (size_t)win_x * BIN_PRE * 2;
    static_assert((3 * BIN_CH + BIN_BWIN + BIN_SF + 1 + BIN_BWIN * BIN_ROW) * 4 + BIN_BWIN * BIN_PRE * 2 <= 100 * 1024, "scatter kernels: > 2 blocks per SM");
    EOGS_CUDA(cudaFuncSetAttribute(bin_scatter_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    EOGS_CUDA(cudaFuncSetAttribute(bin_scatter_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));

    bin_count_rows_kernel<<<n_chunks, BIN_THREADS, 0, s>>>(n_items, gy, band.row_begin, order, rect, row_count);
    EOGS_LAUNCH_CHECK("bin_count_rows_kernel");
    bin_scan_rows_kernel<<<gy, 256, 0, s>>>(n_chunks, gy, row_count, row_total);
    EOGS_LAUNCH_CHECK("bin_scan_rows_kernel");
    bin_row_offsets_kernel<<<1, 1024, 0, s>>>(gy, row_total, row_start, sub_first);
    EOGS_LAUNCH_CHECK("bin_row_offsets_kernel");
    bin_scatter_rows_kernel<<<n_chunks, BIN_THREADS, smem_rows, s>>>(n_items, gy, band.row_begin, order, rect, row_count,
                                                                      row_start, runs);
    EOGS_LAUNCH_CHECK("bin_scatter_rows_kernel");
    prof_mark(s, ST_BIN_ROWS);

    // the number of sub-chunks is known on the device only (n_subs bounds it): persistent blocks stride over them, as
    // many as are resident at once (one wave: no late blocks with a full share of the work)
    int per_sm_count = 0, per_sm_scatter = 0;
    EOGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_count, bin_count_cols_kernel, BIN_THREADS, 0));
    EOGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_scatter, bin_scatter_cols_kernel, BIN_THREADS, smem_cols));
    const uint32_t count_grid = min(n_subs, (uint32_t)(sm_count_cached() * max(per_sm_count, 1)));
    const uint32_t col_grid = min(n_subs, (uint32_t)(sm_count_cached() * max(per_sm_scatter, 1)));
    bin_count_cols_kernel<<<count_grid, BIN_THREADS, 0, s>>>(gy, gx, sub_first, row_start, row_total, runs, col_count);
    EOGS_LAUNCH_CHECK("bin_count_cols_kernel");
    bin_scan_cols_kernel<<<(tiles + 7) / 8, 256, 0, s>>>(gy, gx, sub_first, col_count, ranges);
    EOGS_LAUNCH_CHECK("bin_scan_cols_kernel");
    bin_tile_ranges_kernel<<<1, 1024, 0, s>>>(tiles, ranges);
    EOGS_LAUNCH_CHECK("bin_tile_ranges_kernel");
    prof_mark(s, ST_BIN_COUNT);

    bin_scatter_cols_kernel<<<col_grid, BIN_THREADS, smem_cols, s>>>(gy, gx, sub_first, row_start, row_total, runs, col_count,
                                                                    ranges, point_list);
    EOGS_LAUNCH_CHECK("bin_scatter_cols_kernel");
    prof_mark(s, ST_BIN_SCATTER);
    return 0;
}

}  // namespace eogs
