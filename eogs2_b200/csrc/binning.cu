// Binning: depth order of Gaussians, instance emission, tile sort, tile ranges.
// Replaces cub InclusiveSum + duplicateWithKeys + 64-bit cub SortPairs + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:70-138, 280-321).
//
// The reference sorts I instances on a 64-bit (tile | depth bits) key: ceil((32+bit)/8) = 6-7
// radix passes over 12-byte pairs.  The order it produces is fully determined — (tile, depth
// bits, Gaussian id) because the LSD sort is stable and emission is id-ascending — so we get
// the identical list with far less traffic:
//   1. sort the P Gaussians once by (depth bits, id)          (P << I, 32-bit keys)
//   2. emit instances in that order                            (balanced, coalesced)
//   3. stable-sort instances by tile id only                  (ceil(bit/8) = 2-3 passes over
//                                                              8-byte pairs)
// Algorithmic bytes per instance: 8 emitted + 16*passes*... see DESIGN.md.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace eogs {

size_t sort_temp_bound(size_t n) { return (size_t(1) << 20) + n; }

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
    L.offsets = take(n * 4);
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

BinningLayout binning_layout(int W, int H, uint32_t I) {
    (void)W; (void)H;
    BinningLayout L;
    const size_t n = (size_t)(I > 0 ? I : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.key_in = take(n * 4);
    L.key_out = take(n * 4);
    L.val_in = take(n * 4);
    L.temp_bytes = sort_temp_bound(n);
    L.temp = take(L.temp_bytes);
    L.total = off;
    return L;
}

// ---- stage 1b: depth order + offsets -------------------------------------------------------
struct GatherTiles {
    const uint32_t* tiles;
    __host__ __device__ uint32_t operator()(uint32_t g) const { return tiles[g]; }
};

int launch_depth_order(cudaStream_t s, int P, char* geom, const GeomLayout& L,
                       eogs_forward_info* info_dev)
{
    uint32_t* key_in = reinterpret_cast<uint32_t*>(geom + L.key_in);
    uint32_t* key_out = reinterpret_cast<uint32_t*>(geom + L.key_out);
    uint32_t* id_in = reinterpret_cast<uint32_t*>(geom + L.id_in);
    uint32_t* order = reinterpret_cast<uint32_t*>(geom + L.order);
    uint32_t* offsets = reinterpret_cast<uint32_t*>(geom + L.offsets);
    const uint32_t* tiles = reinterpret_cast<const uint32_t*>(geom + L.tiles);
    void* temp = geom + L.temp;

    // (depth bits, id): keys are non-negative floats, so their bit patterns order like the
    // values; the radix sort is stable and ids come in ascending, which yields the tie order.
    size_t need = 0;
    cub::DoubleBuffer<uint32_t> keys(key_in, key_out);
    // preprocess wrote 0..P-1 into `order` (see launch_preprocess_fwd): 4 passes land back in `order`
    cub::DoubleBuffer<uint32_t> vals(order, id_in);
    EOGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, vals, P, 0, 32, s));
    if (need > L.temp_bytes) { set_error("depth sort temp %zu > %zu", need, L.temp_bytes); return -3; }
    need = L.temp_bytes;
    EOGS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, keys, vals, P, 0, 32, s));
    if (vals.Current() != order)
        EOGS_CUDA(cudaMemcpyAsync(order, vals.Current(), (size_t)P * 4, cudaMemcpyDeviceToDevice, s));

    prof_mark(s, ST_DEPTH_SORT);
    auto in = thrust::make_transform_iterator(static_cast<const uint32_t*>(order), GatherTiles{tiles});
    need = 0;
    EOGS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, need, in, offsets, P, s));
    if (need > L.temp_bytes) { set_error("scan temp %zu > %zu", need, L.temp_bytes); return -3; }
    need = L.temp_bytes;
    EOGS_CUDA(cub::DeviceScan::InclusiveSum(temp, need, in, offsets, P, s));

    (void)info_dev;      // num_instances = offsets[P-1] was already published by the preprocess kernel
    prof_mark(s, ST_SCAN);
    return 0;
}

// ---- stage 2a: instance emission -----------------------------------------------------------
// One block per 256 depth-ordered Gaussians.  The block's instances form one contiguous output
// range; thread t writes outputs t, t+256, ... and finds the owning Gaussian by binary search in
// the block's 256 relative offsets (shared memory).  Balanced regardless of footprint size, and
// both stores are fully coalesced — the reference loops serially per Gaussian
// (rasterizer_impl.cu:96-108).
constexpr int EMIT_THREADS = 256;

template <typename KeyT>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(int P, int grid_x, uint32_t band_y0, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect,
                      KeyT* __restrict__ tile_keys, uint32_t* __restrict__ ids)
{
    __shared__ uint32_t s_end[EMIT_THREADS];
    __shared__ uint32_t s_id[EMIT_THREADS];
    __shared__ uint2 s_rect[EMIT_THREADS];

    const int base = blockIdx.x * EMIT_THREADS;
    const int i = base + threadIdx.x;
    const uint32_t block_start = base > 0 ? __ldg(offsets + base - 1) : 0u;
    const int last = min(base + EMIT_THREADS, P) - 1;
    const uint32_t total = __ldg(offsets + last) - block_start;
    if (total == 0u) return;

    if (i < P) {
        const uint32_t g = __ldg(order + i);
        s_end[threadIdx.x] = __ldg(offsets + i) - block_start;
        s_id[threadIdx.x] = g;
        s_rect[threadIdx.x] = __ldg(rect + g);
    } else {
        s_end[threadIdx.x] = total;
        s_id[threadIdx.x] = 0u;
        s_rect[threadIdx.x] = make_uint2(0u, 0u);
    }
    __syncthreads();

    for (uint32_t t = threadIdx.x; t < total; t += EMIT_THREADS) {
        // smallest k with s_end[k] > t
        int lo = 0;
#pragma unroll
        for (int step = EMIT_THREADS / 2; step > 0; step >>= 1)
            if (s_end[lo + step - 1] <= t) lo += step;
        const uint32_t start = lo > 0 ? s_end[lo - 1] : 0u;
        const uint32_t local = t - start;
        const uint2 r = s_rect[lo];
        const uint32_t x0 = r.x & 0xFFFFu, y0 = r.x >> 16, x1 = r.y & 0xFFFFu;
        const uint32_t w = x1 - x0;
        const uint32_t ry = local / w, rx = local - ry * w;   // row-major over (y, x), rasterizer_impl.cu:96-99
        tile_keys[block_start + t] = (KeyT)((y0 - band_y0 + ry) * (uint32_t)grid_x + (x0 + rx));   // band-relative tile id
        ids[block_start + t] = s_id[lo];
    }
}

// ---- stage 2c: tile ranges -------------------------------------------------------------------
// identifyTileRanges (rasterizer_impl.cu:116-138) on the sorted tile keys.  Each thread scans
// KPT consecutive keys fetched with one 16-byte load (the reference: one thread, two scalar loads
// per key), so the kernel streams at HBM rate; boundaries are rare (one per non-empty tile).
template <typename KeyT>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(uint32_t I, const KeyT* __restrict__ keys, uint2* __restrict__ ranges)
{
    constexpr uint32_t KPT = 16 / sizeof(KeyT);               // keys per thread: 8 (u16) or 4 (u32)
    const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * KPT;
    if (base >= I) return;
    KeyT k[KPT];
    if (base + KPT <= I) {
        *reinterpret_cast<uint4*>(k) = __ldg(reinterpret_cast<const uint4*>(keys + base));    // base*sizeof(KeyT) % 16 == 0
    } else {
#pragma unroll
        for (uint32_t j = 0; j < KPT; j++) k[j] = base + j < I ? __ldg(keys + base + j) : (KeyT)0;
    }
    uint32_t prev = base > 0 ? (uint32_t)__ldg(keys + base - 1) : 0u;
    if (base == 0) ranges[(uint32_t)k[0]].x = 0;
#pragma unroll
    for (uint32_t j = 0; j < KPT; j++) {
        const uint32_t idx = base + j;
        if (idx >= I) break;
        const uint32_t cur = (uint32_t)k[j];
        if (idx > 0 && cur != prev) { ranges[prev].y = idx; ranges[cur].x = idx; }
        if (idx == I - 1) ranges[cur].y = I;
        prev = cur;
    }
}

// getHigherMsb (rasterizer_impl.cu:35-50): number of key bits needed for tile ids
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

int launch_binning(cudaStream_t s, int P, int W, int H, Band band, uint32_t I, const char* geom,
                   const GeomLayout& GL, uint32_t* point_list, char* binning,
                   const BinningLayout& BL, char* image, const ImageLayout& IL)
{
    const int grid_x = (W + TILE - 1) / TILE;
    const uint32_t tiles = (uint32_t)grid_x * (uint32_t)band.rows();
    uint2* ranges = reinterpret_cast<uint2*>(image + IL.ranges);
    EOGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s));
    if (I == 0) return 0;
    if (I > 0x7FFFFFFFu) {        // cub::DeviceRadixSort takes a signed 32-bit item count
        set_error("%u (Gaussian, tile) instances exceed 2^31-1: render the view in tile bands (eogs_*_band)", I);
        return -3;
    }

    uint32_t* val_in = reinterpret_cast<uint32_t*>(binning + BL.val_in);
    // The reference sorts bits [0, 32 + getHigherMsb(tiles)) (rasterizer_impl.cu:306-311); tile ids are
    // < tiles, so the bits of tiles - 1 are all the significant ones and a stable sort over just those
    // gives the identical order.  This matters exactly at powers of two: the 4096^2 sun view has 65 536
    // tiles = 16 significant bits (two 8-bit passes on 16-bit keys), where getHigherMsb says 17.
    const int bit = tiles > 1 ? (int)higher_msb(tiles - 1) : 1;

    // Tile ids fit 16 bits up to 65 536 tiles (4096^2 images): 16-bit sort keys cut the traffic of
    // every sort pass from 16 to 12 bytes per instance.  Larger grids sort 32-bit keys.
    auto run = [&](auto key_tag) -> int {
        using KeyT = decltype(key_tag);
        KeyT* key_in = reinterpret_cast<KeyT*>(binning + BL.key_in);
        KeyT* key_out = reinterpret_cast<KeyT*>(binning + BL.key_out);
        // An onesweep sort of `bit` bits takes ceil(bit/8) passes and ping-pongs between the two value
        // buffers: emit into the one that makes the LAST pass land in point_list (no copy afterwards).
        const bool even_passes = (((bit + 7) / 8) & 1) == 0;
        uint32_t* ids_first = even_passes ? point_list : val_in;
        uint32_t* ids_other = even_passes ? val_in : point_list;
        emit_instances_kernel<KeyT><<<(P + EMIT_THREADS - 1) / EMIT_THREADS, EMIT_THREADS, 0, s>>>(
            P, grid_x, (uint32_t)band.row_begin, reinterpret_cast<const uint32_t*>(geom + GL.order),
            reinterpret_cast<const uint32_t*>(geom + GL.offsets),
            reinterpret_cast<const uint2*>(geom + GL.rect), key_in, ids_first);
        EOGS_LAUNCH_CHECK("emit_instances_kernel");
        prof_mark(s, ST_EMIT);

        cub::DoubleBuffer<KeyT> keys(key_in, key_out);
        cub::DoubleBuffer<uint32_t> vals(ids_first, ids_other);
        size_t need = 0;
        EOGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, vals, (int)I, 0, bit, s));
        if (need > BL.temp_bytes) { set_error("tile sort temp %zu > %zu", need, BL.temp_bytes); return -3; }
        need = BL.temp_bytes;
        EOGS_CUDA(cub::DeviceRadixSort::SortPairs(binning + BL.temp, need, keys, vals, (int)I, 0, bit, s));
        if (vals.Current() != point_list)
            EOGS_CUDA(cudaMemcpyAsync(point_list, vals.Current(), (size_t)I * 4, cudaMemcpyDeviceToDevice, s));
        prof_mark(s, ST_TILE_SORT);

        constexpr uint32_t per_block = 256u * (16u / sizeof(KeyT));
        tile_ranges_kernel<KeyT><<<(I + per_block - 1) / per_block, 256, 0, s>>>(I, keys.Current(), ranges);
        EOGS_LAUNCH_CHECK("tile_ranges_kernel");
        return 0;
    };
    if (int rc = tiles <= 0x10000u ? run(uint16_t{}) : run(uint32_t{})) return rc;
    prof_mark(s, ST_RANGES);
    return 0;
}

}  // namespace eogs
