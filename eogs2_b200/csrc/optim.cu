// Replicated optimiser step and prune compaction on ONE flat parameter buffer.
//
// EOGS++ keeps six Adam parameter groups (xyz, f_dc, f_rest, opacity, scaling, rotation;
// scene/gaussian_model.py:223-271, eps 1e-15) and, on every prune, rebuilds each parameter and both moment
// tensors with boolean-mask indexing plus Python-side optimizer-state surgery (:451-505).  Here the
// parameters live segment by segment in one flat fp32 buffer with the layout of the data-parallel gradient
// bucket (dp.py), so that
//   - the Adam step of ALL groups is ONE kernel over the flat buffer (per-segment learning rates), reading
//     the all-reduced gradient bucket in place;
//   - pruning is an exclusive scan of the keep flags (CUB) + one gather kernel per segment for parameters
//     and both moments.
// HBM-bound streaming: Adam reads 16 B and writes 12 B per scalar parameter; compaction reads and writes
// each surviving row once.
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace eogs {

constexpr int ADAM_MAX_SEG = 8;
struct AdamSegs {
    int n;
    unsigned long long end[ADAM_MAX_SEG];    // exclusive end offset of each segment in the flat buffer
    float step_size[ADAM_MAX_SEG];           // lr / (1 - beta1^t)
};

// torch.optim.Adam (single tensor path, amsgrad = False, weight_decay = 0):
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
//   denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps;  param.addcdiv_(exp_avg, denom, value = -lr / (1 - beta1^t))
__global__ void __launch_bounds__(256)
adam_step_kernel(unsigned long long n, AdamSegs segs, float one_minus_beta1, float beta2, float one_minus_beta2,
                 float eps, float bc2_sqrt,
                 float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float step_size = segs.step_size[0];
#pragma unroll
    for (int k = 1; k < ADAM_MAX_SEG; k++)
        if (k < segs.n && i >= segs.end[k - 1]) step_size = segs.step_size[k];
    const float g = grad[i];
    float m = exp_avg[i], v = exp_avg_sq[i];
    m = fmaf(g - m, one_minus_beta1, m);                   // lerp, weight < 0.5 form
    v = fmaf(g * g, one_minus_beta2, v * beta2);
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    param[i] = param[i] - step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(int P, int width, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ offsets,
                   const float* __restrict__ src, float* __restrict__ dst)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P * width;
    if (t >= total) return;
    const int row = (int)(t / width), col = (int)(t - (long long)row * width);
    if (keep[row]) dst[(size_t)offsets[row] * width + col] = src[t];
}

struct KeepToU32 {
    __host__ __device__ uint32_t operator()(uint8_t k) const { return k ? 1u : 0u; }
};

__global__ void prune_count_kernel(int P, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ offsets,
                                   uint32_t* __restrict__ count) {
    *count = offsets[P - 1] + (keep[P - 1] ? 1u : 0u);
}

// ---- densification (densify_and_clone / densify_and_split / densify_and_prune, gaussian_model.py:573-704) ----
// Selection masks on the P rows that existed before densification.  grads = xyz_gradient_accum / denom with
// NaN -> 0 (:682-683); clone: |grads| >= thr and max(exp(log_scale)) <= percent_dense * extent (:633-640);
// split: grads >= thr and max(exp(log_scale)) > percent_dense * extent (:581-586; rows appended by the clone
// step have a padded gradient of 0 and are never selected).
__global__ void __launch_bounds__(256)
densify_select_kernel(int P, const float* __restrict__ grad_accum, const float* __restrict__ denom,
                      const float* __restrict__ log_scales, float grad_threshold, float size_threshold,
                      uint8_t* __restrict__ clone_flag, uint8_t* __restrict__ split_flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float g = grad_accum[i] / denom[i];
    if (g != g) g = 0.f;
    const float ms = fmaxf(fmaxf(expf(log_scales[3 * (size_t)i]), expf(log_scales[3 * (size_t)i + 1])),
                           expf(log_scales[3 * (size_t)i + 2]));
    clone_flag[i] = (fabsf(g) >= grad_threshold && ms <= size_threshold) ? 1 : 0;
    split_flag[i] = (g >= grad_threshold && ms > size_threshold) ? 1 : 0;
}

// The K = N * Ks child rows of a split hold copies of their parents (gathered N times).  In place (:592-603):
//   xyz += build_rotation(rotation) @ (noise * exp(log_scale));   log_scale = log(exp(log_scale) / (0.8 N))
// build_rotation (utils/general_utils.py:82-105) normalises the quaternion.  noise [K,3] is standard normal,
// drawn by the caller (torch.normal(0, stds) = randn * stds, :590-591) so that ranks share the stream.
__global__ void __launch_bounds__(256)
densify_split_children_kernel(int K, int N, float* __restrict__ xyz, float* __restrict__ log_scales,
                              const float* __restrict__ rotations, const float* __restrict__ noise)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    float s[3], smp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s[k] = expf(log_scales[3 * (size_t)c + k]);
        smp[k] = noise[3 * (size_t)c + k] * s[k];
    }
    const float q0 = rotations[4 * (size_t)c], q1 = rotations[4 * (size_t)c + 1], q2 = rotations[4 * (size_t)c + 2],
                q3 = rotations[4 * (size_t)c + 3];
    const float nrm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    const float r = q0 / nrm, x = q1 / nrm, y = q2 / nrm, z = q3 / nrm;
    const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                           {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                           {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    const float div = 0.8f * (float)N;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        xyz[3 * (size_t)c + k] += R[k][0] * smp[0] + R[k][1] * smp[1] + R[k][2] * smp[2];
        log_scales[3 * (size_t)c + k] = logf(s[k] / div);
    }
}

// Keep mask over the Pn rows after densification (:60x prune_filter of the split parents, then :690-700):
// drop split parents, sigmoid(opacity) < min_opacity, and — when ws_threshold >= 0 — max(exp(log_scale)) >
// ws_threshold (big_points_ws).  big_points_vs tests max_radii2D, which densification_postfix has just reset
// to zeros (:571), so it never fires in the reference and is not evaluated here.
__global__ void __launch_bounds__(256)
densify_keep_kernel(int Pn, int P_old, const uint8_t* __restrict__ split_flag, const float* __restrict__ opacity_logits,
                    const float* __restrict__ log_scales, float min_opacity, float ws_threshold,
                    uint8_t* __restrict__ keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Pn) return;
    bool k = !(i < P_old && split_flag[i]);
    const float op = 1.f / (1.f + expf(-opacity_logits[i]));
    if (op < min_opacity) k = false;
    if (ws_threshold >= 0.f) {
        const float ms = fmaxf(fmaxf(expf(log_scales[3 * (size_t)i]), expf(log_scales[3 * (size_t)i + 1])),
                               expf(log_scales[3 * (size_t)i + 2]));
        if (ms > ws_threshold) k = false;
    }
    keep[i] = k ? 1 : 0;
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_adam_step(eogs_stream_t stream, unsigned long long n, int num_segments,
                            const unsigned long long* segment_end, const float* lr,
                            double beta1, double beta2, double eps, int step,
                            float* params, const float* grads, float* exp_avg, float* exp_avg_sq)
{
    if (num_segments <= 0 || num_segments > ADAM_MAX_SEG || step <= 0) { set_error("bad segments/step"); return -1; }
    if (n == 0) return 0;
    if (!segment_end || !lr || !params || !grads || !exp_avg || !exp_avg_sq) { set_error("null argument"); return -4; }
    AdamSegs segs;
    segs.n = num_segments;
    // bias corrections in double like the Python reference, then rounded once
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int k = 0; k < ADAM_MAX_SEG; k++) {
        segs.end[k] = k < num_segments ? segment_end[k] : n;
        segs.step_size[k] = k < num_segments ? (float)((double)lr[k] / bc1) : 0.f;
    }
    if (segs.end[num_segments - 1] != n) { set_error("last segment must end at n"); return -1; }
    const unsigned long long blocks = (n + 255) / 256;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, segs, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)sqrt(bc2), params, grads,
        exp_avg, exp_avg_sq);    // scalars are formed in double like torch's Python side, then rounded once
    EOGS_LAUNCH_CHECK("adam_step_kernel");
    return 0;
}

EOGS_API size_t eogs_prune_temp_bytes(int P) {
    size_t need = 0;
    auto in = thrust::make_transform_iterator(static_cast<const uint8_t*>(nullptr), KeepToU32{});
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, static_cast<uint32_t*>(nullptr), P > 0 ? P : 1);
    return need + 256;
}

EOGS_API int eogs_prune_offsets(eogs_stream_t stream, int P, const uint8_t* keep, uint32_t* offsets,
                                void* temp, size_t temp_bytes, uint32_t* count_dev)
{
    if (P < 0) { set_error("bad P"); return -1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!count_dev) { set_error("null argument"); return -4; }
    if (P == 0) { EOGS_CUDA(cudaMemsetAsync(count_dev, 0, 4, s)); return 0; }
    if (!keep || !offsets || !temp) { set_error("null argument"); return -4; }
    auto in = thrust::make_transform_iterator(keep, KeepToU32{});
    size_t need = 0;
    EOGS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, in, offsets, P, s));
    if (need > temp_bytes) { set_error("prune scan temp %zu > %zu", need, temp_bytes); return -3; }
    EOGS_CUDA(cub::DeviceScan::ExclusiveSum(temp, need, in, offsets, P, s));
    prune_count_kernel<<<1, 1, 0, s>>>(P, keep, offsets, count_dev);
    EOGS_LAUNCH_CHECK("prune_count_kernel");
    return 0;
}

EOGS_API int eogs_prune_gather(eogs_stream_t stream, int P, int width, const uint8_t* keep,
                               const uint32_t* offsets, const float* src, float* dst)
{
    if (P < 0 || width <= 0) { set_error("bad sizes"); return -1; }
    if (P == 0) return 0;
    if (!keep || !offsets || !src || !dst) { set_error("null argument"); return -4; }
    const long long total = (long long)P * width;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, width, keep, offsets, src, dst);
    EOGS_LAUNCH_CHECK("gather_rows_kernel");
    return 0;
}

EOGS_API int eogs_densify_select(eogs_stream_t stream, int P, const float* grad_accum, const float* denom,
                                 const float* log_scales, float grad_threshold, float size_threshold,
                                 uint8_t* clone_flag, uint8_t* split_flag)
{
    if (P < 0) { set_error("bad P"); return -1; }
    if (P == 0) return 0;
    if (!grad_accum || !denom || !log_scales || !clone_flag || !split_flag) { set_error("null argument"); return -4; }
    densify_select_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, grad_accum, denom, log_scales, grad_threshold, size_threshold, clone_flag, split_flag);
    EOGS_LAUNCH_CHECK("densify_select_kernel");
    return 0;
}

EOGS_API int eogs_densify_split_children(eogs_stream_t stream, int K, int N, float* xyz, float* log_scales,
                                         const float* rotations, const float* noise)
{
    if (K < 0 || N <= 0) { set_error("bad sizes"); return -1; }
    if (K == 0) return 0;
    if (!xyz || !log_scales || !rotations || !noise) { set_error("null argument"); return -4; }
    densify_split_children_kernel<<<(K + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        K, N, xyz, log_scales, rotations, noise);
    EOGS_LAUNCH_CHECK("densify_split_children_kernel");
    return 0;
}

EOGS_API int eogs_densify_keep(eogs_stream_t stream, int Pn, int P_old, const uint8_t* split_flag,
                               const float* opacity_logits, const float* log_scales, float min_opacity,
                               float ws_threshold, uint8_t* keep)
{
    if (Pn < 0 || P_old < 0 || P_old > Pn) { set_error("bad sizes"); return -1; }
    if (Pn == 0) return 0;
    if ((P_old > 0 && !split_flag) || !opacity_logits || !log_scales || !keep) { set_error("null argument"); return -4; }
    densify_keep_kernel<<<(Pn + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        Pn, P_old, split_flag, opacity_logits, log_scales, min_opacity, ws_threshold, keep);
    EOGS_LAUNCH_CHECK("densify_keep_kernel");
    return 0;
}

}  // extern "C"
