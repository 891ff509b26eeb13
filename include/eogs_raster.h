/*
 * eogs_raster.h — C ABI of libeogs_raster.so, the B200 (sm_100a) affine-camera
 * Gaussian-splatting rasterizer that stands in for EOGS++'s
 * diff-gaussian-rasterization extension ("DGR" below =
 * src/gaussiansplatting/submodules/diff-gaussian-rasterization of gardiens/EOGS2).
 *
 * Boundary rules
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - Every pointer marked "dev" is a CUDA device pointer owned by the caller
 *     (torch owns every buffer, as in DGR/rasterize_points.cu:69-85,163-174).
 *     The library never allocates or frees device memory.
 *   - Every call enqueues work on the given stream and returns without
 *     synchronising, except where stated.  No global mutable state: calls on
 *     different streams / devices are independent.
 *   - Return value: 0 = ok, <0 = argument error, >0 = cudaError_t.
 *     eogs_last_error() returns a thread-local message for the last failure.
 *
 * What each entry point replaces in the reference
 *   eogs_forward_geometry + eogs_forward_render
 *        = _C.rasterize_gaussians            (DGR/ext.cpp:16, DGR/rasterize_points.cu:35-124,
 *                                             CudaRasterizer::Rasterizer::forward,
 *                                             DGR/cuda_rasterizer/rasterizer_impl.cu:198-341)
 *   eogs_rasterize_forward (single call with allocation callbacks)
 *        = the same, in the shape of Rasterizer::forward's std::function<char*(size_t)>
 *          resize callbacks (DGR/cuda_rasterizer/rasterizer.h:31-56)
 *   eogs_backward
 *        = _C.rasterize_gaussians_backward   (DGR/ext.cpp:17, DGR/rasterize_points.cu:126-224,
 *                                             Rasterizer::backward, rasterizer_impl.cu:345-452)
 *          plus the three torch reductions of _RasterizeGaussians.backward
 *          (DGR/diff_gaussian_rasterization/__init__.py:174-202), returned as 14 sums.
 *   eogs_mark_visible
 *        = _C.mark_visible                   (DGR/ext.cpp:18, rasterize_points.cu:226-245)
 *
 * The forward is split in two because the number of (Gaussian, tile) instances is
 * only known after the geometry stage; the reference hides the same dependency
 * behind a blocking cudaMemcpy (rasterizer_impl.cu:284).
 */
#ifndef EOGS_RASTER_H_INCLUDED
#define EOGS_RASTER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EOGS_API __attribute__((visibility("default")))
#else
#define EOGS_API
#endif

#define EOGS_ABI_VERSION 7
#define EOGS_TILE 16            /* BLOCK_X = BLOCK_Y = 16, DGR/cuda_rasterizer/config.h:15-16 */
#define EOGS_MAX_CHANNELS 5     /* NUM_CHANNELS 5,        DGR/cuda_rasterizer/config.h:14    */

/* error bits reported in eogs_forward_info.error */
#define EOGS_ERR_ALTITUDE_ABOVE_200 1u  /* reference: printf + __trap(), forward.cu:267-272 */
#define EOGS_ERR_TOO_MANY_INSTANCES 2u  /* sum of tiles touched wrapped 32 bits (the reference's scan wraps silently,
                                           rasterizer_impl.cu:280-284): render the view in tile bands */

typedef void* eogs_stream_t;    /* cudaStream_t */

/* Host-visible result of the geometry stage.  info_host must be pinned host memory.  Its words are written right after
 * the projection kernel, BEFORE the depth sort — by the device itself when info_host is mapped into the device address
 * space (cudaHostAlloc / cudaMallocHost / torch pin_memory under unified addressing: payload, system-scope fence,
 * `ready`; by the projection kernel's last warp up to 2^18 Gaussians, by a one-thread kernel behind it above), otherwise
 * by two stream-ordered copies (payload, then `ready`): a host
 * that zeroes info_host->ready before the call and polls it afterwards has I while the rest of the geometry stage still
 * runs; a host that synchronises the stream sees the same values. */
typedef struct eogs_forward_info {
    uint32_t num_instances;     /* I = sum of tiles touched = reference's num_rendered */
    uint32_t error;             /* EOGS_ERR_* bits */
    uint32_t ready;             /* non-zero once num_instances / error are final */
    uint32_t reserved;          /* device copy: warps of the projection kernel that have published (internal) */
} eogs_forward_info;

EOGS_API int eogs_abi_version(void);
EOGS_API const char* eogs_last_error(void);

/* Instrumentation (developer builds): work counters of the blend kernels — evaluated / blended (pixel, Gaussian)
 * pairs, lane slots, list entries — since the last reset.  out[16]; returns 1 when the library
 * was compiled with -DEOGS_COUNT_PAIRS=1 (libeogs_raster_count.so), 0 and zeros for the product library. */
EOGS_API int eogs_debug_counters(unsigned long long* out, int reset);

/* ---- scratch sizing (host only, no CUDA calls that touch a device) ---------------- */
/* Per-Gaussian state kept from forward to backward (packed splat records, depths, tile
 * rects, depth order, offsets) + the temporary space of the depth sort and scan.
 * Counterpart of required<GeometryState>(P), rasterizer_impl.cu:155-170. */
EOGS_API size_t eogs_geom_bytes(int P);
/* Per-image state kept for backward: final transmittance, last contributor, tile ranges.
 * Counterpart of required<ImageState>(W*H), rasterizer_impl.cu:172-179. */
EOGS_API size_t eogs_image_bytes(int W, int H);
/* Temporary space of the tile sort for `num_instances` instances (freed after forward).
 * Counterpart of required<BinningState>(I), rasterizer_impl.cu:181-194, minus the sorted
 * Gaussian list, which is the separate `point_list` argument (4*I bytes) because it is
 * the only part backward needs. */
EOGS_API size_t eogs_binning_bytes(int W, int H, uint32_t num_instances);
/* 32-bit words of the `point_list` buffer for I instances: the I sorted Gaussian ids (identical to the reference's
 * binningState.point_list, rasterizer_impl.h:56-61) followed by one culling byte per instance that the forward blend
 * writes and the backward blend reads (which 8x8 regions of its tile the instance can reach). */
EOGS_API size_t eogs_point_list_words(uint32_t num_instances);
/* Floats of backward scratch for P Gaussians: one 16-float gradient record per Gaussian (what the
 * reference keeps in dL_dconic / dL_dmeans2D / dL_dopacity / dL_dcolors between its two backward
 * kernels, rasterize_points.cu:163-183) + a small tail holding the tile-queue counter of the blend
 * backward.  Zeroed by every backward call. */
EOGS_API size_t eogs_grad_scratch_floats(int P);

/* ---- forward, stage 1: per-Gaussian geometry -------------------------------------- */
/* Affine projection, 3D->2D covariance, conic, radius, tile rect, depth = 200 - altitude,
 * depth ordering and instance offsets.  Replaces preprocessCUDA (forward.cu:154-283) and
 * cub InclusiveSum (rasterizer_impl.cu:280).
 *   means3D  [P,3]   dev   scales [P,3] / rotations [P,4] dev, or cov3D_precomp [P,6] dev
 *   opacities[P]     dev   colors [P,channels] dev (colors_precomp; required, as in
 *                          rasterizer_impl.cu:244-247)
 *   viewmatrix[16]   dev   transposed [[A,b],[0,1]]: A[r][c] = v[4c+r], b[r] = v[12+r]
 *   radii    [P] i32 dev   out, 0 for culled Gaussians
 *   geom            dev   eogs_geom_bytes(P) bytes, 256-byte aligned
 *   info_dev        dev   8 bytes; info_host: 8 bytes of pinned host memory, filled by an
 *                          async copy on `stream` — synchronise the stream before reading. */
EOGS_API int eogs_forward_geometry(eogs_stream_t stream, int P, int W, int H, int channels,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host);

/* ---- forward, stage 2: binning + blend --------------------------------------------- */
/* Instance emission, tile sort, tile ranges, front-to-back alpha blend.  Replaces
 * duplicateWithKeys, cub SortPairs, identifyTileRanges (rasterizer_impl.cu:70-138,292-321)
 * and renderCUDA (forward.cu:288-411).
 *   point_list [eogs_point_list_words(I)] u32 dev  out: first I words = Gaussian ids sorted by (tile, depth bits, id)
 *                           — identical to the reference's binningState.point_list; then I culling bytes
 *   binning          dev    eogs_binning_bytes(W,H,I) bytes of scratch
 *   image            dev    eogs_image_bytes(W,H) bytes (kept for backward)
 *   bg [channels]    dev
 *   out_color [channels,H,W] dev, out_invdepth [H,W] dev (may be NULL) */
EOGS_API int eogs_forward_render(eogs_stream_t stream, int P, int W, int H, int channels,
                        uint32_t num_instances, const void* geom, uint32_t* point_list,
                        void* binning, void* image, const float* bg,
                        float* out_color, float* out_invdepth);

/* ---- forward, single call with allocation callbacks -------------------------------- */
/* Same as the two stages back to back; `alloc(user, which, bytes)` must return a device
 * pointer of at least `bytes` bytes (which: 0 = geom, 1 = binning scratch, 2 = image,
 * 3 = point_list).  Synchronises `stream` once, like rasterizer_impl.cu:284.
 * Returns the pointers it obtained through *geom / *point_list / *image so the caller can
 * hand them to eogs_backward.  *num_instances receives I. */
typedef void* (*eogs_alloc_fn)(void* user, int which, size_t bytes);
EOGS_API int eogs_rasterize_forward(eogs_stream_t stream, int P, int W, int H, int channels,
                           const float* means3D, const float* scales, const float* rotations,
                           const float* cov3D_precomp, const float* opacities, const float* colors,
                           const float* viewmatrix, float scale_modifier, int antialiasing,
                           const float* bg, eogs_alloc_fn alloc, void* user,
                           int32_t* radii, float* out_color, float* out_invdepth,
                           void** geom, uint32_t** point_list, void** image,
                           uint32_t* num_instances);

/* ---- backward ----------------------------------------------------------------------- */
/* Blend backward + preprocess backward + camera-gradient reductions.
 *   dL_dpix [channels,H,W] dev; dL_dinvdepth [H,W] dev or NULL
 *   grad_scratch dev: eogs_grad_scratch_floats(P) floats (16 per Gaussian + 16), zeroed by this call
 * outputs (all dev, all written by this call; culled Gaussians get zeros):
 *   dL_dmeans2D [P,3] (z = 0)   dL_dcolors [P,channels]   dL_dopacity [P]
 *   dL_dmeans3D [P,3]           dL_dcov3D [P,6] or NULL    dL_dscales [P,3] or NULL
 *   dL_drotations [P,4] or NULL
 *   cam_sums [16]: [0..5] = sum_p dL_dT[p][0..5]                  (__init__.py:180-192)
 *                  [6..11] = means3D^T @ dL_dmeans2D[:, :2], row-major 3x2 (__init__.py:195-196)
 *                  [12..13] = sum_p dL_dmeans2D[p][0..1]          (__init__.py:199-202)
 *   projmatrix[16] dev: used for dL_dmeans3D like backward.cu:439-445 (pass viewmatrix when
 *   they are the same tensor). */
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
                  float* dL_drotations, float* cam_sums);

/* ---- tile-band variants (multi-GPU sharding of ONE view) ---------------------------- */
/* Not in the reference (it has no multi-device code): tiles are independent
 * after binning, so a huge view (BASELINE configs[4]: 8192^2) is split into horizontal bands of
 * tile rows [row_begin, row_end), one band per GPU.  Each *_band call behaves like its
 * whole-image counterpart restricted to the band's tiles:
 *   - every Gaussian is projected (radii keep the whole-image meaning, identical on every rank);
 *     only the (Gaussian, tile) instances inside the band are emitted, sorted and blended, so
 *     the band's sorted list and ranges equal the whole-image ones restricted to those tiles
 *     (list offsets shifted by the band's base);
 *   - per-image buffers are band-compact: `image` has eogs_image_bytes_band() bytes, and
 *     out_color [channels, band_h, W], out_invdepth [band_h, W], dL_dpix [channels, band_h, W],
 *     dL_dinvdepth [band_h, W] hold pixel rows [16*row_begin, min(H, 16*row_end)) only;
 *   - eogs_backward_band returns the band's share of every gradient and of cam_sums; the
 *     whole-image gradient is the sum over bands (an NCCL all-reduce across ranks).
 * The whole-image entry points above are the band [0, ceil(H/16)). */
EOGS_API size_t eogs_image_bytes_band(int W, int H, int row_begin, int row_end);
EOGS_API int eogs_forward_geometry_band(eogs_stream_t stream, int P, int W, int H, int channels,
                          int row_begin, int row_end,
                          const float* means3D, const float* scales, const float* rotations,
                          const float* cov3D_precomp, const float* opacities, const float* colors,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host);
EOGS_API int eogs_forward_render_band(eogs_stream_t stream, int P, int W, int H, int channels,
                        int row_begin, int row_end,
                        uint32_t num_instances, const void* geom, uint32_t* point_list,
                        void* binning, void* image, const float* bg,
                        float* out_color, float* out_invdepth);
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
                  float* dL_drotations, float* cam_sums);

/* ---- fused render glue ---------------------------------- */
/* EOGS++'s render() (gaussian_renderer/renderer.py:84-107) runs ~10 torch kernels before every
 * rasterizer call and as many after its backward: exp / normalize / sigmoid activations
 * (scene/gaussian_model.py:41-53,109-137), SH2RGB of the DC coefficients (utils/sh_utils.py:125-126),
 * the altitude colour A_2 . xyz + b_2 (AffineCamera.ECEF_to_UVA, scene/cameras/affine_cameras.py:
 * 432-438), a ones column and a concatenation.  These entry points take the RAW parameters of
 * GaussianModel instead and do all of that inside the geometry kernels, forward and backward:
 *   xyz [P,3]   log_scales [P,3] (_scaling)   raw_rotations [P,4] (_rotation)
 *   opacity_logits [P] (_opacity)   features_dc [P,3] (_features_dc)   alt_affine [4] dev = (a, b) of
 *   altitude = a . xyz + b (column 2 of the camera's transposed `affine`)
 * and render 5 channels [rgb, altitude, 1].  The render stage is the ordinary
 * eogs_forward_render_band.  Backward outputs are gradients w.r.t. the raw parameters; alt_sums [4]
 * = sum_p dL_daltitude_p * (x, y, z, 1), the gradient of the altitude affine row. */
EOGS_API int eogs_forward_geometry_params_band(eogs_stream_t stream, int P, int W, int H,
                          int row_begin, int row_end,
                          const float* xyz, const float* log_scales, const float* raw_rotations,
                          const float* opacity_logits, const float* features_dc, const float* alt_affine,
                          const float* viewmatrix, float scale_modifier, int antialiasing,
                          int32_t* radii, void* geom, eogs_forward_info* info_dev,
                          eogs_forward_info* info_host);
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
                  float* cam_sums, float* alt_sums);

/* ---- virtual-camera resample (sun-view shadow pass) --------- */
/* render_resample_virtual_camera, steps 2-3 (gaussian_renderer/renderer_cc_shadow.py:32-46):
 * virtual_uv = (cam2virt @ rendered_uva)[..., :2]; bilinear grid_sample of the virtual camera's render
 * (align_corners=True, zeros padding) at virtual_uv; altitude = -100 where |u| > 1 or |v| > 1.
 *   virtual_render [Cv,Hv,Wv] dev (Cv >= 4: rgb, altitude, ...)   cam2virt [3,3] dev row-major
 *   rendered_uva [H,W,3] dev   ->   out_rgb [3,H,W], out_altitude [H,W], out_uv [H,W,2]
 * Backward: dL_drgb / dL_daltitude / dL_duv may be NULL (no upstream gradient); dL_dvirtual
 * [Cv,Hv,Wv] and dL_dcam2virt [9] are zeroed by the call, dL_duva [H,W,3] is written. */
EOGS_API int eogs_resample_forward(eogs_stream_t stream, int Cv, int Hv, int Wv, int H, int W,
                                   const float* virtual_render, const float* cam2virt, const float* rendered_uva,
                                   float* out_rgb, float* out_altitude, float* out_uv);
EOGS_API int eogs_resample_backward(eogs_stream_t stream, int Cv, int Hv, int Wv, int H, int W,
                                    const float* virtual_render, const float* cam2virt, const float* rendered_uva,
                                    const float* dL_drgb, const float* dL_daltitude, const float* dL_duv,
                                    float* dL_dvirtual, float* dL_duva, float* dL_dcam2virt);

/* ---- photometric loss --------------------------------------- */
/* L = (1 - lambda) * mean|image - gt| + lambda * (1 - SSIM(image, gt))   (loss/shadow.py:21-29 with
 * l1_loss / ssim of utils/loss_utils.py:18-85: 11x11 Gaussian window, sigma 1.5, zero padding).
 *   window11 [11] HOST floats: the normalised 1-D window (gaussian(11, 1.5))
 *   image, gt [C,H,W] dev      maps [3,C,H,W] dev: per-pixel partials kept for the backward
 *   sums2 [2] dev scratch      out3 [3] dev: {loss, mean SSIM, mean L1}
 * Backward: dL_dloss [1] dev or NULL (= 1); dL_dimage [C,H,W] dev is written in the planar layout the
 * rasterizer's backward consumes. */
EOGS_API int eogs_photometric_forward(eogs_stream_t stream, int C, int H, int W, const float* window11,
                                      const float* image, const float* gt, float lambda_dssim,
                                      float* maps, float* sums2, float* out3);
EOGS_API int eogs_photometric_backward(eogs_stream_t stream, int C, int H, int W, const float* window11,
                                       const float* image, const float* gt, float lambda_dssim,
                                       const float* maps, const float* dL_dloss, float* dL_dimage);

/* ---- replicated optimiser step and prune compaction --------- */
/* Adam of all parameter groups in one launch over a flat fp32 buffer split into `num_segments` contiguous
 * segments (GaussianModel.training_setup, scene/gaussian_model.py:223-271: one group per parameter, eps 1e-15).
 *   segment_end [num_segments] HOST: exclusive end offset of each segment (last = n); lr [num_segments] HOST
 *   step: 1-based step count (bias corrections are computed in double on the host, like torch)
 *   params / exp_avg / exp_avg_sq [n] dev are updated in place; grads [n] dev (e.g. the all-reduced bucket). */
EOGS_API int eogs_adam_step(eogs_stream_t stream, unsigned long long n, int num_segments,
                            const unsigned long long* segment_end, const float* lr,
                            double beta1, double beta2, double eps, int step,
                            float* params, const float* grads, float* exp_avg, float* exp_avg_sq);
/* Prune compaction (prune_points / _prune_optimizer, scene/gaussian_model.py:466-505):
 *   eogs_prune_offsets: exclusive scan of keep[P] (u8) -> offsets[P] (destination row of every kept row),
 *                       *count_dev = number of kept rows; temp: eogs_prune_temp_bytes(P) bytes of scratch.
 *   eogs_prune_gather:  dst[offsets[r], :] = src[r, :] for kept rows of a [P, width] fp32 array. */
EOGS_API size_t eogs_prune_temp_bytes(int P);
EOGS_API int eogs_prune_offsets(eogs_stream_t stream, int P, const uint8_t* keep, uint32_t* offsets,
                                void* temp, size_t temp_bytes, uint32_t* count_dev);
EOGS_API int eogs_prune_gather(eogs_stream_t stream, int P, int width, const uint8_t* keep,
                               const uint32_t* offsets, const float* src, float* dst);

/* Densification (densify_and_prune = densify_and_clone + densify_and_split + prune, scene/gaussian_model.py:573-704)
 * on the same flat layout; the row copies reuse eogs_prune_offsets / eogs_prune_gather.
 *   eogs_densify_select: clone_flag / split_flag [P] u8 from grads = grad_accum / denom (NaN -> 0), grad_threshold and
 *                        size_threshold = percent_dense * scene_extent (:581-586, :633-640)
 *   eogs_densify_split_children: in place on the K = N*Ks child rows (copies of their parents):
 *                        xyz += R(rotation) (noise * exp(log_scale)); log_scale = log(exp(log_scale) / (0.8 N)) (:590-603)
 *   eogs_densify_keep:   keep [Pn] u8 after densification: not a split parent (rows < P_old), sigmoid(opacity) >=
 *                        min_opacity, and max exp(log_scale) <= ws_threshold when ws_threshold >= 0 (:690-700) */
EOGS_API int eogs_densify_select(eogs_stream_t stream, int P, const float* grad_accum, const float* denom,
                                 const float* log_scales, float grad_threshold, float size_threshold,
                                 uint8_t* clone_flag, uint8_t* split_flag);
EOGS_API int eogs_densify_split_children(eogs_stream_t stream, int K, int N, float* xyz, float* log_scales,
                                         const float* rotations, const float* noise);
EOGS_API int eogs_densify_keep(eogs_stream_t stream, int Pn, int P_old, const uint8_t* split_flag,
                               const float* opacity_logits, const float* log_scales, float min_opacity,
                               float ws_threshold, uint8_t* keep);

/* ---- simple-knn distCUDA2 ------------------------------------ */
/* distCUDA2 (submodules/simple-knn/spatial.cu:15-26 -> SimpleKNN::knn, simple_knn.cu:187-222): for every point the
 * mean of the squared distances to its 3 nearest neighbours (other indices; duplicates count with distance 0;
 * missing neighbours count as 1e37 like the reference's FLT_MAX).  Bit-exact against the reference: same
 * per-pair expression fma(dz,dz,fma(dx,dx,dy*dy)) and ((b0+b1)+b2)/3.  Finite inputs.
 *   points [P,3] dev fp32   scratch: eogs_knn_bytes(P) bytes dev   mean_dist2 [P] dev fp32
 * Stream-ordered, no host synchronisation (the reference blocks twice on D2H copies of the bounding box). */
EOGS_API size_t eogs_knn_bytes(int P);
EOGS_API int eogs_knn_dist2(eogs_stream_t stream, int P, const float* points, void* scratch, size_t scratch_bytes,
                            float* mean_dist2);

/* ---- DSM splat ------------------------------------------------ */
/* The `plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma)` call of compute_dsm_from_view
 * (utils/dsm_utils.py:27-37): every point (x, y, v) is averaged, with weight exp(-dist^2 / (2 sigma^2)), into
 * the raster cells within `radius` cells of its own; unreached cells are NaN.
 *   cloud [N,3] dev fp64 (x, y, value)   accum [2*xsize*ysize] dev fp64 scratch (zeroed by the call)
 *   raster [ysize, xsize] dev fp32 (the [:, :, 0] plane of plyflatten's result) */
EOGS_API int eogs_dsm_splat(eogs_stream_t stream, long long N, const double* cloud, double xoff, double yoff,
                            double resolution, int xsize, int ysize, int radius, float sigma,
                            double* accum, float* raster);

/* ---- NVLS gradient all-reduce (data parallel over views) --------------------------- */
/* In-switch all-reduce (SUM, fp32) of a bucket that lives in symmetric memory mapped to one multicast address on all
 * ranks: this rank reduces its 1/world slice with multimem.ld_reduce and broadcasts it with multimem.st.  The caller
 * brackets the call with cross-GPU barriers (before: all ranks' gradients written; after: all slices landed).
 *   multicast_ptr: the multicast mapping of the bucket (NULL -> error: fall back to NCCL); n_floats % 4 == 0 */
EOGS_API int eogs_nvls_allreduce(eogs_stream_t stream, void* multicast_ptr, unsigned long long n_floats, int rank, int world);
/* Peer-to-peer variant for small worlds (2 GPUs: nothing to gain from the switch's reduction): this rank sums its 1/world
 * slice out of every rank's copy with plain NVLink loads and stores the result into every copy.  buffer_ptrs: HOST array
 * of `world` device pointers, entry r = rank r's copy of the bucket as mapped into this process (symmetric memory).
 * Same barrier contract as eogs_nvls_allreduce. */
EOGS_API int eogs_p2p_allreduce(eogs_stream_t stream, void* const* buffer_ptrs, unsigned long long n_floats, int rank, int world);

/* ---- markVisible ------------------------------------------------------------------- */
/* The reference's in_frustum culls nothing for affine cameras (its body is
 * commented out, auxiliary.h:151-176): every Gaussian is reported visible. */
EOGS_API int eogs_mark_visible(eogs_stream_t stream, int P, const float* means3D,
                      const float* viewmatrix, const float* projmatrix, uint8_t* present);

/* ---- instrumentation ---------------------------------------------------------------- */
/* Per-stage device times (CUDA events on the launch stream) of the calls made on this
 * thread; off by default.  eogs_profile_read fills ms[stage] (milliseconds) for stage ids
 * 1 preprocess, 2 depth sort, 3 scan, 4 emit, 5 tile sort, 6 ranges, 7 blend fwd,
 * 8 bwd zeroing, 9 blend bwd, 10 preprocess bwd; it synchronises the recorded events and
 * returns the number of stages recorded.  The forward render stage continues the timeline
 * of the geometry stage, so any gap the host leaves between them is attributed to stage 4 (emit); the copy of the
 * instance count to the host sits in stage 2 (it is enqueued before the depth sort). */
EOGS_API int eogs_profile_enable(int on);
EOGS_API int eogs_profile_read(float* ms, int n);

/* ---- inspection (parity tests) ----------------------------------------------------- */
/* Copies internal state into caller buffers in the reference's layouts so that tests can
 * compare bit for bit.  Any output pointer may be NULL.
 *   means2D [P,2], depths [P], conic_opacity [P,4], tiles_touched [P] u32 : geomState fields
 *   keys_sorted [I] u64 = (tile << 32) | depth bits, rebuilt from point_list
 *   ranges [tiles,2] u32, final_T [H*W], n_contrib [H*W] u32 : imgState fields */
EOGS_API int eogs_export_state(eogs_stream_t stream, int P, int W, int H, uint32_t num_instances,
                      const void* geom, const uint32_t* point_list, const void* image,
                      float* means2D, float* depths, float* conic_opacity,
                      uint32_t* tiles_touched, uint64_t* keys_sorted,
                      uint32_t* ranges, float* final_T, uint32_t* n_contrib);

/* alpha_cut of an array of (antialias-scaled) opacities: the smallest exponent `power` for which the forward's
 * test !(opacity * expf(power) < 1/255) (forward.cu:367-372) accepts a pair; +inf when it never does.  The blend
 * backward takes its accept decision from this per-Gaussian threshold.  flags bit 0: the test accepts power = cut,
 * bit 1: it still accepts the next float below cut (must be 0). */
EOGS_API int eogs_debug_alpha_cut(eogs_stream_t stream, int n, const float* opacity, float* cut, uint32_t* flags);

/* The depth-order stage alone (test hook): order[P] = indices 0..P-1 sorted by (keys[i], i), keys = depth bit patterns
 * with 0xFFFFFFFF for culled Gaussians, exactly what eogs_forward_geometry runs between projection and binning
 * (the depth half of the reference's instance sort, rasterizer_impl.cu:306-311).
 * scratch: device memory of eogs_debug_depth_order_bytes(P). */
EOGS_API size_t eogs_debug_depth_order_bytes(int P);
EOGS_API int eogs_debug_depth_order(eogs_stream_t stream, int P, const uint32_t* keys, uint32_t* order, void* scratch);

/* Band variant: tiles_touched counts the band's tiles, keys_sorted carry whole-image tile ids,
 * ranges [band tiles,2], final_T / n_contrib [band_h*W]. */
EOGS_API int eogs_export_state_band(eogs_stream_t stream, int P, int W, int H, int row_begin, int row_end,
                      uint32_t num_instances,
                      const void* geom, const uint32_t* point_list, const void* image,
                      float* means2D, float* depths, float* conic_opacity,
                      uint32_t* tiles_touched, uint64_t* keys_sorted,
                      uint32_t* ranges, float* final_T, uint32_t* n_contrib);

#ifdef __cplusplus
}
#endif
#endif /* EOGS_RASTER_H_INCLUDED */
