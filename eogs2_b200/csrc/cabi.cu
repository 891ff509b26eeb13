// extern "C" surface of libeogs_raster.so (see include/eogs_raster.h for the contract and the
// reference interfaces each entry point replaces).
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace eogs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

struct Profiler {
    bool on = false;
    bool created = false;
    cudaEvent_t start{};
    cudaEvent_t ev[ST_COUNT]{};
    bool seen[ST_COUNT]{};
    int order[ST_COUNT * 2]{};
    int n = 0;
};
static thread_local Profiler g_prof;

void prof_begin(cudaStream_t s) {
    Profiler& p = g_prof;
    if (!p.on) return;
    if (!p.created) {
        cudaEventCreate(&p.start);
        for (int i = 0; i < ST_COUNT; i++) cudaEventCreate(&p.ev[i]);
        p.created = true;
    }
    for (int i = 0; i < ST_COUNT; i++) p.seen[i] = false;
    p.n = 0;
    cudaEventRecord(p.start, s);
}

void prof_mark(cudaStream_t s, int stage) {
    Profiler& p = g_prof;
    if (!p.on || !p.created || stage <= 0 || stage >= ST_COUNT || p.seen[stage]) return;
    p.seen[stage] = true;
    p.order[p.n++] = stage;
    cudaEventRecord(p.ev[stage], s);
}

// The geometry stage's words, stored by the device into the host's mapped pinned struct: payload, system fence, `ready`.
__global__ void publish_info_kernel(eogs_forward_info* info, volatile eogs_forward_info* host) {
    host->num_instances = info->num_instances;
    host->error = info->error;
    __threadfence_system();
    host->ready = 1u;
    info->ready = 1u;
}

__global__ void fill_u8_kernel(uint8_t* p, int n, uint8_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Rebuild the reference's per-Gaussian / per-instance views of our packed state.
__global__ void export_geom_kernel(int P, const float4* __restrict__ splat, const float* __restrict__ depth,
                                   const uint32_t* __restrict__ tiles, float* means2D, float* depths,
                                   float* conic_opacity, uint32_t* tiles_touched) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 a = splat[(size_t)i * REC_F4], b = splat[(size_t)i * REC_F4 + 1];
    if (means2D) { means2D[2 * (size_t)i] = a.x; means2D[2 * (size_t)i + 1] = a.y; }
    if (depths) depths[i] = depth[i];
    if (conic_opacity) {
        conic_opacity[4 * (size_t)i] = a.z; conic_opacity[4 * (size_t)i + 1] = a.w;
        conic_opacity[4 * (size_t)i + 2] = b.x; conic_opacity[4 * (size_t)i + 3] = b.y;
    }
    if (tiles_touched) tiles_touched[i] = tiles[i];
}

__global__ void export_keys_kernel(uint32_t num_tiles, uint32_t tile_base, const uint2* __restrict__ ranges,
                                   const uint32_t* __restrict__ point_list,
                                   const float* __restrict__ depth, uint64_t* keys) {
    const uint32_t tile = blockIdx.x;
    if (tile >= num_tiles) return;
    const uint2 r = ranges[tile];
    for (uint32_t j = r.x + threadIdx.x; j < r.y; j += blockDim.x)
        keys[j] = ((uint64_t)(tile + tile_base) << 32) | (uint64_t)__float_as_uint(depth[point_list[j]]);
}

static int check_common(int P, int W, int H, int channels) {
    if (P < 0 || W <= 0 || H <= 0) { set_error("bad sizes P=%d W=%d H=%d", P, W, H); return -1; }
    if (channels != 3 && channels != 5) { set_error("channels must be 3 or 5, got %d", channels); return -1; }
    if ((W + TILE - 1) / TILE > 0xFFFF || (H + TILE - 1) / TILE > 0xFFFF) { set_error("image too large"); return -1; }
    return 0;
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_abi_version(void) { return EOGS_ABI_VERSION; }
EOGS_API const char* eogs_last_error(void) { return g_err; }

// Instrumentation: per-stage device times of the calls made on this thread since the last
// eogs_profile_enable(1) / eogs_profile_read().  ms[stage] for the Stage enum of common.cuh;
// stages that did not run read 0.  Synchronises the recorded events.
EOGS_API int eogs_profile_enable(int on) { g_prof.on = on != 0; return 0; }
EOGS_API int eogs_profile_read(float* ms, int n) {
    Profiler& p = g_prof;
    for (int i = 0; i < n; i++) ms[i] = 0.f;
    if (!p.created) return 0;
    cudaEvent_t prev = p.start;
    for (int k = 0; k < p.n; k++) {
        const int st = p.order[k];
        EOGS_CUDA(cudaEventSynchronize(p.ev[st]));
        float t = 0.f;
        EOGS_CUDA(cudaEventElapsedTime(&t, prev, p.ev[st]));
        if (st < n) ms[st] += t;
        prev = p.ev[st];
    }
    return p.n;
}

// Work counters of the blend kernels since the last reset (pairs_eval, pairs_blend ...).  Returns 1 and
// fills out[16] (Counter enum of common.cuh) when the library was built with -DEOGS_COUNT_PAIRS=1, 0 (zeros) otherwise.
// Synchronises the device.
EOGS_API int eogs_debug_counters(unsigned long long* out, int reset) {
    for (int i = 0; i < CNT_COUNT; i++) out[i] = 0ull;
#if EOGS_COUNT_PAIRS
    EOGS_CUDA(cudaDeviceSynchronize());
    if (int rc = read_counters_fwd(out, reset != 0)) return rc < 0 ? rc : -rc;
    if (int rc = read_counters_bwd(out, reset != 0)) return rc < 0 ? rc : -rc;
    return 1;
#else
    (void)reset;
    return 0;
#endif
}

EOGS_API size_t eogs_geom_bytes(int P) { return geom_layout(P).total; }
EOGS_API size_t eogs_image_bytes(int W, int H) { return image_layout(W, H, full_band(H)).total; }
EOGS_API size_t eogs_image_bytes_band(int W, int H, int row_begin, int row_end) {
    const Band b{row_begin, row_end};
    if (W <= 0 || H <= 0 || check_band(H, b)) return 0;
    return image_layout(W, H, b).total;
}
EOGS_API size_t eogs_binning_bytes(int W, int H, uint32_t I) { return binning_layout(W, H, I).total; }
EOGS_API size_t eogs_point_list_words(uint32_t I) { return (size_t)I + ((size_t)I + 3) / 4; }
EOGS_API size_t eogs_grad_scratch_floats(int P) { return (size_t)(P > 0 ? P : 0) * GRAD_STRIDE + GRAD_TAIL; }

// Device-side alias of a pinned host struct, or nullptr when the device cannot write it directly (cached per thread:
// callers reuse one struct per stream).
static eogs_forward_info* mapped_alias(eogs_forward_info* host)
{
    static thread_local eogs_forward_info* last_host = nullptr;
    static thread_local eogs_forward_info* last_dev = nullptr;
    static thread_local int last_device = -1;
    if (!host) return nullptr;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (host == last_host && dev == last_device) return last_dev;
    cudaPointerAttributes attr{};
    eogs_forward_info* alias = nullptr;
    if (cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        alias = static_cast<eogs_forward_info*>(attr.devicePointer);
    else
        (void)cudaGetLastError();
    last_host = host; last_dev = alias; last_device = dev;
    return alias;
}

static int forward_geometry_impl(eogs_stream_t stream, int P, int W, int H, int channels,
                          int row_begin, int row_end, bool raw_params, const float* alt_affine,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    const Band band{row_begin, row_end};
    if (int rc = check_band(H, band)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!colors) { set_error("For non-RGB, provide precomputed Gaussian colors!"); return -4; }   // rasterizer_impl.cu:244-247
    if (!cov3D_precomp && (!scales || !rotations)) { set_error("need scales+rotations or cov3D_precomp"); return -4; }
    if (!means3D || !opacities || !viewmatrix || !radii || !geom || !info_dev) { set_error("null argument"); return -4; }
    EOGS_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(eogs_forward_info), s));
    prof_begin(s);
    // The host needs I once per forward.  Its copy to the host's pinned struct is enqueued right after the projection
    // kernel, BEFORE the depth sort: a host that polls info_host->ready has I while the sort runs and can enqueue the
    // render stage behind it — the GPU never waits for the host round trip (the reference blocks on a cudaMemcpy after
    // its scan, rasterizer_impl.cu:284).  Small scenes: the kernel's last warp writes the pinned struct itself (when it
    // is mapped into the device address space, which cudaHostAlloc / torch pin_memory memory is), see preprocess.cu.
    constexpr int DIRECT_PUBLISH_MAX_P = 1 << 18;
    eogs_forward_info* host_mapped = P > 0 ? mapped_alias(info_host) : nullptr;
    const bool in_kernel = host_mapped && P <= DIRECT_PUBLISH_MAX_P;
    if (P > 0) {
        const GeomLayout L = geom_layout(P);
        char* g = static_cast<char*>(geom);
        if (int rc = launch_preprocess_fwd(s, P, W, H, band, channels, raw_params, means3D, scales, rotations,
                                           cov3D_precomp, opacities, colors, viewmatrix, alt_affine, scale_modifier,
                                           antialiasing != 0, radii, g, L, info_dev, in_kernel ? host_mapped : nullptr)) return rc;
        prof_mark(s, ST_PREPROCESS);
    }
    if (host_mapped && !in_kernel) {
        // large scenes: one single-thread kernel behind the projection stores the words into the mapped host struct
        // (one launch instead of a memset and two copy operations in front of the depth sort)
        publish_info_kernel<<<1, 1, 0, s>>>(info_dev, host_mapped);
        EOGS_LAUNCH_CHECK("publish_info_kernel");
    } else if (!host_mapped) {
        // Host struct not mapped into the device address space (or P = 0): two stream-ordered copies, the payload (I, error)
        // first, the `ready` word second.  A host that sees `ready` set therefore reads a complete payload (one 16-byte copy
        // gives no such guarantee: CUDA does not promise that a host thread observes a device -> host copy atomically or in
        // word order before the stream operation completes).
        EOGS_CUDA(cudaMemsetAsync(&info_dev->ready, 0x01, sizeof(uint32_t), s));
        if (info_host) {
            EOGS_CUDA(cudaMemcpyAsync(info_host, info_dev, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            EOGS_CUDA(cudaMemcpyAsync(&info_host->ready, &info_dev->ready, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        }
    }
    if (P > 0) {
        if (int rc = launch_depth_order(s, P, static_cast<char*>(geom), geom_layout(P), info_dev)) return rc;
    }
    return 0;
}

EOGS_API int eogs_forward_geometry_band(eogs_stream_t stream, int P, int W, int H, int channels,
                          int row_begin, int row_end,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host)
{
    return forward_geometry_impl(stream, P, W, H, channels, row_begin, row_end, false, nullptr, means3D, scales,
                                 rotations, cov3D_precomp, opacities, colors, viewmatrix, scale_modifier, antialiasing,
                                 radii, geom, info_dev, info_host);
}

EOGS_API int eogs_forward_geometry_params_band(eogs_stream_t stream, int P, int W, int H,
                          int row_begin, int row_end,
                          const float* xyz, const float* log_scales, const float* raw_rotations,
                          const float* opacity_logits, const float* features_dc, const float* alt_affine,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host)
{
    if (!alt_affine || !log_scales || !raw_rotations || !features_dc) { set_error("null argument"); return -4; }
    return forward_geometry_impl(stream, P, W, H, 5, row_begin, row_end, true, alt_affine, xyz, log_scales,
                                 raw_rotations, nullptr, opacity_logits, features_dc, viewmatrix, scale_modifier,
                                 antialiasing, radii, geom, info_dev, info_host);
}

EOGS_API int eogs_forward_geometry(eogs_stream_t stream, int P, int W, int H, int channels,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    return eogs_forward_geometry_band(stream, P, W, H, channels, 0, (H + TILE - 1) / TILE, means3D, scales, rotations,
                                      cov3D_precomp, opacities, colors, viewmatrix, scale_modifier, antialiasing,
                                      radii, geom, info_dev, info_host);
}

EOGS_API int eogs_debug_alpha_cut(eogs_stream_t stream, int n, const float* opacity, float* cut, uint32_t* flags)
{
    if (n < 0) { set_error("bad n"); return -1; }
    if (n > 0 && (!opacity || !cut || !flags)) { set_error("null argument"); return -4; }
    return launch_alpha_cut_debug(static_cast<cudaStream_t>(stream), n, opacity, cut, flags);
}

EOGS_API size_t eogs_debug_depth_order_bytes(int P) { return debug_depth_order_bytes(P); }
EOGS_API int eogs_debug_depth_order(eogs_stream_t stream, int P, const uint32_t* keys, uint32_t* order, void* scratch)
{
    if (P < 0) { set_error("bad P"); return -1; }
    if (P > 0 && (!keys || !order || !scratch)) { set_error("null argument"); return -4; }
    return debug_depth_order(static_cast<cudaStream_t>(stream), P, keys, order, scratch);
}

EOGS_API int eogs_forward_render_band(eogs_stream_t stream, int P, int W, int H, int channels,
                        int row_begin, int row_end,
                        uint32_t num_instances, const void* geom, uint32_t* point_list,
                        void* binning, void* image, const float* bg,
                        float* out_color, float* out_invdepth)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    const Band band{row_begin, row_end};
    if (int rc = check_band(H, band)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!geom || !image || !bg || !out_color) { set_error("null argument"); return -4; }
    if (num_instances > 0 && (!point_list || !binning)) { set_error("null binning buffers"); return -4; }
    const GeomLayout GL = geom_layout(P);
    const ImageLayout IL = image_layout(W, H, band);
    const BinningLayout BL = binning_layout(W, H, num_instances);
    if (int rc = launch_binning(s, P, W, H, band, num_instances, static_cast<const char*>(geom), GL, point_list,
                                static_cast<char*>(binning), BL, static_cast<char*>(image), IL)) return rc;
    // one culling byte per instance lives behind the I ids of the point_list allocation (eogs_point_list_words)
    uint8_t* masks = reinterpret_cast<uint8_t*>(point_list + num_instances);
    if (int rc = launch_blend_fwd(s, W, H, band, channels, static_cast<const char*>(geom), GL, point_list, masks,
                                  static_cast<char*>(image), IL, bg, out_color, out_invdepth)) return rc;
    prof_mark(s, ST_BLEND_FWD);
    return 0;
}

EOGS_API int eogs_forward_render(eogs_stream_t stream, int P, int W, int H, int channels,
                        uint32_t num_instances, const void* geom, uint32_t* point_list,
                        void* binning, void* image, const float* bg,
                        float* out_color, float* out_invdepth)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    return eogs_forward_render_band(stream, P, W, H, channels, 0, (H + TILE - 1) / TILE, num_instances, geom,
                                    point_list, binning, image, bg, out_color, out_invdepth);
}

EOGS_API int eogs_rasterize_forward(eogs_stream_t stream, int P, int W, int H, int channels,
                           const float* means3D, const float* scales, const float* rotations,
                           const float* cov3D_precomp, const float* opacities, const float* colors,
                           const float* viewmatrix, float scale_modifier, int antialiasing,
                           const float* bg, eogs_alloc_fn alloc, void* user,
                           int32_t* radii, float* out_color, float* out_invdepth,
                           void** geom, uint32_t** point_list, void** image,
                           uint32_t* num_instances)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    if (!alloc || !geom || !point_list || !image || !num_instances) { set_error("null argument"); return -4; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    *geom = alloc(user, 0, eogs_geom_bytes(P) + 256);
    *image = alloc(user, 2, eogs_image_bytes(W, H));
    if (!*geom || !*image) { set_error("allocation callback returned NULL"); return -5; }
    // the info words live in the tail of the geometry buffer
    eogs_forward_info* info_dev = reinterpret_cast<eogs_forward_info*>(static_cast<char*>(*geom) + eogs_geom_bytes(P));
    eogs_forward_info info_host = {0u, 0u, 0u, 0u};
    if (int rc = eogs_forward_geometry(stream, P, W, H, channels, means3D, scales, rotations, cov3D_precomp,
                                       opacities, colors, viewmatrix, scale_modifier, antialiasing, radii,
                                       *geom, info_dev, nullptr)) return rc;
    EOGS_CUDA(cudaMemcpyAsync(&info_host, info_dev, sizeof(info_host), cudaMemcpyDeviceToHost, s));
    EOGS_CUDA(cudaStreamSynchronize(s));
    if (info_host.error & EOGS_ERR_ALTITUDE_ABOVE_200) {
        set_error("Point is too high: altitude above 200 (reference: __trap, forward.cu:267-272)");
        return -6;
    }
    if (info_host.error & EOGS_ERR_TOO_MANY_INSTANCES) {
        set_error("more than 2^32 (Gaussian, tile) instances: render the view in tile bands (eogs_*_band)");
        return -3;
    }
    *num_instances = info_host.num_instances;
    void* binning = nullptr;
    *point_list = nullptr;
    if (info_host.num_instances > 0) {
        *point_list = static_cast<uint32_t*>(alloc(user, 3, eogs_point_list_words(info_host.num_instances) * 4));
        binning = alloc(user, 1, eogs_binning_bytes(W, H, info_host.num_instances));
        if (!*point_list || !binning) { set_error("allocation callback returned NULL"); return -5; }
    }
    return eogs_forward_render(stream, P, W, H, channels, info_host.num_instances, *geom, *point_list,
                               binning, *image, bg, out_color, out_invdepth);
}

static int backward_impl(eogs_stream_t stream, int P, int W, int H, int channels,
                  int row_begin, int row_end, bool raw_params, const float* alt_affine, float* alt_sums,
                  uint32_t num_instances,
                  const float* means3D, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* opacities, const float* colors,
                  const float* viewmatrix, const float* projmatrix,
                  float scale_modifier, int antialiasing, const float* bg,
                  const int32_t* radii, const void* geom, const uint32_t* point_list,
                  const void* image, const float* dL_dpix, const float* dL_dinvdepth,
                  float* grad_scratch,
                  float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                  float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                  float* dL_drotations, float* cam_sums)
{
    (void)colors;
    if (int rc = check_common(P, W, H, channels)) return rc;
    const Band band{row_begin, row_end};
    if (int rc = check_band(H, band)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!cam_sums) { set_error("null argument"); return -4; }
    if (P == 0) {
        EOGS_CUDA(cudaMemsetAsync(cam_sums, 0, 16 * sizeof(float), s));
        if (alt_sums) EOGS_CUDA(cudaMemsetAsync(alt_sums, 0, 4 * sizeof(float), s));
        return 0;
    }
    if (!means3D || !opacities || !viewmatrix || !projmatrix || !bg || !radii || !geom || !image || !dL_dpix ||
        !grad_scratch || !dL_dmeans2D || !dL_dcolors || !dL_dopacity || !dL_dmeans3D) {
        set_error("null argument"); return -4;
    }
    if (!cov3D_precomp && (!scales || !rotations)) { set_error("need scales+rotations or cov3D_precomp"); return -4; }
    const GeomLayout GL = geom_layout(P);
    const ImageLayout IL = image_layout(W, H, band);
    EOGS_CUDA(cudaMemsetAsync(grad_scratch, 0, ((size_t)P * GRAD_STRIDE + GRAD_TAIL) * sizeof(float), s));
    prof_mark(s, ST_BWD_ZERO);
    if (num_instances > 0) {
        if (!point_list) { set_error("null point_list"); return -4; }
        if (int rc = launch_blend_bwd(s, W, H, band, channels, static_cast<const char*>(geom), GL, point_list,
                                      reinterpret_cast<const uint8_t*>(point_list + num_instances),
                                      static_cast<const char*>(image), IL, bg, dL_dpix, dL_dinvdepth,
                                      grad_scratch, reinterpret_cast<uint32_t*>(grad_scratch + (size_t)P * GRAD_STRIDE))) return rc;
    }
    prof_mark(s, ST_BLEND_BWD);
    const int rc_pre = launch_preprocess_bwd(s, P, W, H, channels, raw_params, alt_affine, alt_sums, means3D, scales, rotations, cov3D_precomp, opacities,
                                 viewmatrix, projmatrix, scale_modifier, antialiasing != 0, radii, grad_scratch,
                                 dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dscales,
                                 dL_drotations, cam_sums);
    prof_mark(s, ST_PREPROCESS_BWD);
    return rc_pre;
}

EOGS_API int eogs_backward_band(eogs_stream_t stream, int P, int W, int H, int channels,
                  int row_begin, int row_end, uint32_t num_instances,
                  const float* means3D, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* opacities, const float* colors,
                  const float* viewmatrix, const float* projmatrix,
                  float scale_modifier, int antialiasing, const float* bg,
                  const int32_t* radii, const void* geom, const uint32_t* point_list,
                  const void* image, const float* dL_dpix, const float* dL_dinvdepth,
                  float* grad_scratch,
                  float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                  float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                  float* dL_drotations, float* cam_sums)
{
    return backward_impl(stream, P, W, H, channels, row_begin, row_end, false, nullptr, nullptr, num_instances, means3D,
                         scales, rotations, cov3D_precomp, opacities, colors, viewmatrix, projmatrix, scale_modifier,
                         antialiasing, bg, radii, geom, point_list, image, dL_dpix, dL_dinvdepth, grad_scratch,
                         dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dscales, dL_drotations,
                         cam_sums);
}

EOGS_API int eogs_backward_params_band(eogs_stream_t stream, int P, int W, int H,
                  int row_begin, int row_end, uint32_t num_instances,
                  const float* xyz, const float* log_scales, const float* raw_rotations,
                  const float* opacity_logits, const float* alt_affine,
                  const float* viewmatrix, const float* projmatrix,
                  float scale_modifier, int antialiasing, const float* bg,
                  const int32_t* radii, const void* geom, const uint32_t* point_list,
                  const void* image, const float* dL_dpix, const float* dL_dinvdepth,
                  float* grad_scratch,
                  float* dL_dmeans2D, float* dL_dfeatures_dc, float* dL_dopacity_logits,
                  float* dL_dxyz, float* dL_dlog_scales, float* dL_draw_rotations,
                  float* cam_sums, float* alt_sums)
{
    if (!alt_affine || !alt_sums || !log_scales || !raw_rotations || !dL_dlog_scales || !dL_draw_rotations) {
        set_error("null argument"); return -4;
    }
    return backward_impl(stream, P, W, H, 5, row_begin, row_end, true, alt_affine, alt_sums, num_instances, xyz,
                         log_scales, raw_rotations, nullptr, opacity_logits, nullptr, viewmatrix, projmatrix,
                         scale_modifier, antialiasing, bg, radii, geom, point_list, image, dL_dpix, dL_dinvdepth,
                         grad_scratch, dL_dmeans2D, dL_dfeatures_dc, dL_dopacity_logits, dL_dxyz, nullptr,
                         dL_dlog_scales, dL_draw_rotations, cam_sums);
}

EOGS_API int eogs_backward(eogs_stream_t stream, int P, int W, int H, int channels,
                  uint32_t num_instances,
                  const float* means3D, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* opacities, const float* colors,
                  const float* viewmatrix, const float* projmatrix,
                  float scale_modifier, int antialiasing, const float* bg,
                  const int32_t* radii, const void* geom, const uint32_t* point_list,
                  const void* image, const float* dL_dpix, const float* dL_dinvdepth,
                  float* grad_scratch,
                  float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                  float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                  float* dL_drotations, float* cam_sums)
{
    if (int rc = check_common(P, W, H, channels)) return rc;
    return eogs_backward_band(stream, P, W, H, channels, 0, (H + TILE - 1) / TILE, num_instances, means3D, scales,
                              rotations, cov3D_precomp, opacities, colors, viewmatrix, projmatrix, scale_modifier,
                              antialiasing, bg, radii, geom, point_list, image, dL_dpix, dL_dinvdepth, grad_scratch,
                              dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dscales,
                              dL_drotations, cam_sums);
}

EOGS_API int eogs_mark_visible(eogs_stream_t stream, int P, const float* means3D,
                      const float* viewmatrix, const float* projmatrix, uint8_t* present)
{
    (void)means3D; (void)viewmatrix; (void)projmatrix;
    if (P < 0) { set_error("bad P"); return -1; }
    if (P == 0) return 0;
    if (!present) { set_error("null argument"); return -4; }
    fill_u8_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(present, P, 1);
    EOGS_LAUNCH_CHECK("fill_u8_kernel");
    return 0;
}

EOGS_API int eogs_export_state_band(eogs_stream_t stream, int P, int W, int H, int row_begin, int row_end,
                      uint32_t num_instances, const void* geom, const uint32_t* point_list, const void* image,
                      float* means2D, float* depths, float* conic_opacity,
                      uint32_t* tiles_touched, uint64_t* keys_sorted,
                      uint32_t* ranges, float* final_T, uint32_t* n_contrib)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P < 0 || W <= 0 || H <= 0) { set_error("bad sizes P=%d W=%d H=%d", P, W, H); return -1; }
    const Band band{row_begin, row_end};
    if (int rc = check_band(H, band)) return rc;
    const GeomLayout GL = geom_layout(P);
    const ImageLayout IL = image_layout(W, H, band);
    const char* g = static_cast<const char*>(geom);
    const char* im = static_cast<const char*>(image);
    const uint32_t grid_x = (uint32_t)((W + TILE - 1) / TILE);
    const uint32_t tiles = grid_x * (uint32_t)band.rows();
    const size_t npix = (size_t)W * (size_t)band.height(H);
    if (P > 0 && (means2D || depths || conic_opacity || tiles_touched)) {
        export_geom_kernel<<<(P + 255) / 256, 256, 0, s>>>(
            P, reinterpret_cast<const float4*>(g + GL.splat), reinterpret_cast<const float*>(g + GL.depth),
            reinterpret_cast<const uint32_t*>(g + GL.tiles), means2D, depths, conic_opacity, tiles_touched);
        EOGS_LAUNCH_CHECK("export_geom_kernel");
    }
    if (keys_sorted && num_instances > 0) {
        export_keys_kernel<<<tiles, 128, 0, s>>>(tiles, grid_x * (uint32_t)band.row_begin, reinterpret_cast<const uint2*>(im + IL.ranges), point_list,
                                                  reinterpret_cast<const float*>(g + GL.depth), keys_sorted);
        EOGS_LAUNCH_CHECK("export_keys_kernel");
    }
    if (ranges) EOGS_CUDA(cudaMemcpyAsync(ranges, im + IL.ranges, (size_t)tiles * 8, cudaMemcpyDeviceToDevice, s));
    if (final_T) EOGS_CUDA(cudaMemcpyAsync(final_T, im + IL.final_T, npix * 4, cudaMemcpyDeviceToDevice, s));
    if (n_contrib) EOGS_CUDA(cudaMemcpyAsync(n_contrib, im + IL.n_contrib, npix * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

EOGS_API int eogs_export_state(eogs_stream_t stream, int P, int W, int H, uint32_t num_instances,
                      const void* geom, const uint32_t* point_list, const void* image,
                      float* means2D, float* depths, float* conic_opacity,
                      uint32_t* tiles_touched, uint64_t* keys_sorted,
                      uint32_t* ranges, float* final_T, uint32_t* n_contrib)
{
    return eogs_export_state_band(stream, P, W, H, 0, (H + TILE - 1) / TILE, num_instances, geom, point_list, image,
                                  means2D, depths, conic_opacity, tiles_touched, keys_sorted, ranges, final_T,
                                  n_contrib);
}

}  // extern "C"
