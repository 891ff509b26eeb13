// Shared definitions of the sm_100a rasterizer kernels (internal; the public surface is
// include/eogs_raster.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/eogs_raster.h"

namespace eogs {

constexpr int TILE = EOGS_TILE;          // 16x16 pixel tiles (key parity with DGR config.h:15-16)
constexpr int TILE_PIXELS = TILE * TILE;
constexpr int REC_F4 = 3;                // float4s per packed splat record (48 bytes)
constexpr int GRAD_STRIDE = 16;          // floats per Gaussian in the blend-backward gradient record
constexpr int GRAD_TAIL = 16;            // floats after the P records of grad_scratch: [0] = the backward's tile-queue counter

// ---- packed per-Gaussian splat record (what the blend kernels gather) -----------------
//   rec[0] = { mean2D.x, mean2D.y, conic.x, conic.y }
//   rec[1] = { conic.z, opacity*aa_scale, color0, color1 }
//   rec[2] = { color2, color3, color4, 1/depth }
// One 48-byte record = 1.5 sectors; a gather touches exactly two 32-byte sectors, versus
// four separate arrays (means2D, conic_opacity, colors, depths) in the reference.

// ---- opaque buffer layouts -------------------------------------------------------------
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct GeomLayout {
    size_t splat;        // float4[3P]
    size_t cut;          // float[P]     alpha_cut: the forward accepts a pixel iff power >= cut (blend backward)
    size_t depth;        // float[P]     200 - altitude (garbage-free: culled entries hold +inf bits)
    size_t rect;         // uint2[P]     (x0 | y0<<16, x1 | y1<<16) tile rect, exclusive max
    size_t tiles;        // u32[P]       tiles touched
    size_t key_in;       // u32[P]       depth bits (culled: 0xFFFFFFFF)
    size_t key_out;      // u32[P]
    size_t id_in;        // u32[P]       0..P-1
    size_t order;        // u32[P]       Gaussian ids sorted by (depth bits, id)
    size_t temp;         // CUB temp storage (depth sort)
    size_t temp_bytes;
    size_t total;
};

size_t debug_depth_order_bytes(int P);
int debug_depth_order(cudaStream_t s, int P, const uint32_t* keys, uint32_t* order, void* scratch);

struct ImageLayout {
    size_t final_T;      // float[W*band_height]
    size_t n_contrib;    // u32[W*band_height]
    size_t ranges;       // uint2[band tiles]
    size_t tile_work;    // u32[band tiles]   max(n_contrib) of the tile (blend forward) = the backward's replay length
    size_t tile_order;   // u32[band tiles]   tiles with work > 0, longest first (the persistent backward's queue)
    size_t sched;        // u32[4]            [0] = number of queued tiles
    size_t total;
};

// A band = the tile rows [row_begin, row_end) of the image that one call renders (multi-GPU
// tile sharding of one huge view; the whole image is the band [0, grid_y)).  Per-image state
// (final_T, n_contrib, ranges), the output images and the upstream gradients are band-compact:
// pixel row y of the image is row y - 16*row_begin of the band buffers, tile (x, y) is entry
// (y - row_begin) * grid_x + x of `ranges`.
struct Band {
    int row_begin, row_end;      // tile rows
    __host__ __device__ int rows() const { return row_end - row_begin; }
    __host__ __device__ int y0() const { return row_begin * TILE; }                      // first pixel row
    __host__ __device__ int height(int H) const { return (row_end * TILE < H ? row_end * TILE : H) - row_begin * TILE; }
};

struct BinningLayout {       // scratch of the tile-list construction (binning.cu), freed after the forward
    size_t row_count;    // u32[chunks][band rows]   runs per (chunk of 1024 depth-ordered Gaussians, tile row) -> exclusive bases
    size_t row_total;    // u32[rows + 1]            runs per tile row
    size_t row_start;    // u32[rows + 1]            where the row's run list starts in `runs`
    size_t sub_first;    // u32[rows + 1]            first column sub-chunk (1024 runs) of the row
    size_t runs;         // uint2[<= I]              {Gaussian id, x0 | x1 << 16}, per row in depth order
    size_t col_count;    // u32[sub-chunks][grid_x]  runs covering a tile per sub-chunk -> exclusive bases
    size_t chunk_cap, sub_cap;
    size_t total;
};

size_t sort_temp_bound(size_t n);
int sm_count_cached();      // SMs of the current device (148 on B200)
GeomLayout geom_layout(int P);
ImageLayout image_layout(int W, int H, Band band);
Band full_band(int H);
int check_band(int H, Band band);
BinningLayout binning_layout(int W, int H, uint32_t I);

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define EOGS_CUDA(expr)                                             \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return eogs::cuda_fail(_e, #expr);   \
    } while (0)

#define EOGS_LAUNCH_CHECK(name)                                     \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return eogs::cuda_fail(_e, name);    \
    } while (0)

// ---- optional work counters (a separate build: -DEOGS_COUNT_PAIRS=1 -> libeogs_raster_count.so) -----------------
// The blend kernels count what SURVEY.md section 8d asks every report to carry: evaluated and blended (pixel, Gaussian)
// pairs, the lane slots spent on them and the list entries walked.  The product library compiles none of it.
#ifndef EOGS_COUNT_PAIRS
#define EOGS_COUNT_PAIRS 0
#endif
enum Counter : int {
    CNT_FWD_EVAL = 0, CNT_FWD_BLEND, CNT_FWD_SLOTS, CNT_FWD_ENTRIES,
    CNT_BWD_EVAL, CNT_BWD_BLEND, CNT_BWD_SLOTS, CNT_BWD_ENTRIES, CNT_BWD_FLUSHES, CNT_COUNT = 16
};
// each blend translation unit keeps its own __device__ counters (no relocatable device code) and copies them out
int read_counters_fwd(unsigned long long* out, bool reset);      // fills out[CNT_FWD_*]
int read_counters_bwd(unsigned long long* out, bool reset);      // fills out[CNT_BWD_*]

// ---- optional per-stage timing (instrumentation for bench.py; off by default) ----------------
// When enabled on the calling thread, every stage boundary records a CUDA event on the launch
// stream; eogs_profile_read() synchronises and returns the elapsed milliseconds per stage.
enum Stage : int {
    ST_BEGIN = 0, ST_PREPROCESS, ST_DEPTH_SORT, ST_BIN_ROWS, ST_BIN_COUNT, ST_BIN_SCATTER, ST_BLEND_FWD,
    ST_BWD_ZERO, ST_BLEND_BWD, ST_PREPROCESS_BWD, ST_COUNT
};
void prof_begin(cudaStream_t s);
void prof_mark(cudaStream_t s, int stage);

// ---- stage launchers (one per .cu) --------------------------------------------------------
int launch_preprocess_fwd(cudaStream_t s, int P, int W, int H, Band band, int channels, bool raw_params,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* view, const float* alt_affine, float scale_modifier, bool antialiasing,
                          int32_t* radii, char* geom, const GeomLayout& L, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host_mapped);

int launch_alpha_cut_debug(cudaStream_t s, int n, const float* op, float* cut, uint32_t* flags);

int launch_depth_order(cudaStream_t s, int P, char* geom, const GeomLayout& L,
                       eogs_forward_info* info_dev);

int launch_binning(cudaStream_t s, int P, int W, int H, Band band, uint32_t I, const char* geom,
                   const GeomLayout& GL, uint32_t* point_list, char* binning,
                   const BinningLayout& BL, char* image, const ImageLayout& IL);

int launch_blend_fwd(cudaStream_t s, int W, int H, Band band, int channels, const char* geom,
                     const GeomLayout& GL, const uint32_t* point_list, uint8_t* masks, char* image,
                     const ImageLayout& IL, const float* bg, float* out_color, float* out_invdepth);

int launch_blend_bwd(cudaStream_t s, int W, int H, Band band, int channels, const char* geom,
                     const GeomLayout& GL, const uint32_t* point_list, const uint8_t* masks, const char* image,
                     const ImageLayout& IL, const float* bg, const float* dL_dpix,
                     const float* dL_dinvdepth, float* grad_rec, uint32_t* queue);

int launch_tile_order(cudaStream_t s, int W, int H, Band band, char* image, const ImageLayout& IL);

int launch_preprocess_bwd(cudaStream_t s, int P, int W, int H, int channels, bool raw_params,
                          const float* alt_affine, float* alt_sums,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities,
                          const float* view, const float* proj, float scale_modifier,
                          bool antialiasing, const int32_t* radii, const float* grad_rec,
                          float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                          float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                          float* dL_drotations, float* cam_sums);

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

}  // namespace eogs
