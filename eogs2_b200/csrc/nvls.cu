// Gradient all-reduce through the NVSwitch's in-network reduction (NVLS "multimem" instructions), for the
// data-parallel-over-views step: one kernel that, for this rank's 1/N slice of the flat
// gradient bucket, LOADS the sum of all ranks' copies with multimem.ld_reduce (the switch reads the slice from the
// N GPUs and adds in flight) and STORES it to all ranks' copies with multimem.st (the switch replicates).  Two-shot
// all-reduce in one pass over the slice: per GPU 1x the bucket out and 1x in over NVLink, no intermediate buffers,
// no ring steps.  The bucket lives in symmetric memory (torch.distributed._symmetric_memory: same allocation on
// every rank, mapped into one multicast address); cross-GPU ordering is the caller's: a symmetric-memory barrier
// before (every rank's backward has written its gradients) and after (every slice has landed everywhere) —
// eogs2_b200/nvls.py.  The reference has no communication code at all; the baseline this replaces is
// ncclAllReduce on the same bucket (eogs2_b200/dp.py).
//
// HBM/NVLink-bound streaming: 16-byte multimem accesses, grid = 2 CTAs per SM, grid-stride over the slice.
#include "common.cuh"

namespace eogs {

__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Four independent 16-byte reductions in flight per thread: an NVLS round trip is microseconds long.
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(float4* __restrict__ mc, size_t begin, size_t end)
{
    constexpr int UNROLL = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < end; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) v[k] = mc_ld_reduce(mc + i + k * stride);
#pragma unroll
        for (int k = 0; k < UNROLL; k++) mc_st(mc + i + k * stride, v[k]);
    }
    for (; i < end; i += stride) mc_st(mc + i, mc_ld_reduce(mc + i));
}

// Peer-to-peer variant for small worlds (2 GPUs: no in-switch reduction to gain): this rank sums its 1/N slice out of
// every rank's copy (plain NVLink loads, the own copy first, then the peers in ring order so that the ranks do not all
// hit the same link) and stores the result into every copy.  Per GPU: (N-1)/N of the bucket read and written remotely,
// both directions of the links busy at once.  Each element is touched by exactly one rank, so no cross-GPU hazard exists
// inside the kernel; ordering against the producers / consumers is the caller's pair of barriers, as for the NVLS kernel.
struct PeerPtrs { float4* p[8]; };

__device__ __forceinline__ float4 peer_ld(const float4* p) {            // system scope, not cached in L1: peers rewrite it every step
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float4* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(512)
p2p_allreduce_kernel(PeerPtrs pp, int world, int rank, size_t begin, size_t end)
{
    constexpr int UNROLL = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < end; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) v[k] = peer_ld(pp.p[rank] + i + k * stride);
        for (int r = 1; r < world; r++) {
            const float4* src = pp.p[(rank + r) % world];
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                const float4 a = peer_ld(src + i + k * stride);
                v[k].x += a.x; v[k].y += a.y; v[k].z += a.z; v[k].w += a.w;
            }
        }
        for (int r = 0; r < world; r++) {
            float4* dst = pp.p[(rank + r) % world];
#pragma unroll
            for (int k = 0; k < UNROLL; k++) peer_st(dst + i + k * stride, v[k]);
        }
    }
    for (; i < end; i += stride) {
        float4 v = peer_ld(pp.p[rank] + i);
        for (int r = 1; r < world; r++) {
            const float4 a = peer_ld(pp.p[(rank + r) % world] + i);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        for (int r = 0; r < world; r++) peer_st(pp.p[(rank + r) % world] + i, v);
    }
}

}  // namespace eogs

using namespace eogs;

extern "C" {

EOGS_API int eogs_nvls_allreduce(eogs_stream_t stream, void* multicast_ptr, unsigned long long n_floats, int rank, int world)
{
    if (!multicast_ptr) { set_error("no multicast mapping (NVLS unavailable): use the NCCL all-reduce"); return -4; }
    if (world <= 0 || rank < 0 || rank >= world || (n_floats & 3ull) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15u)) {
        set_error("bad rank/world, or the bucket is not a 16-byte-aligned multiple of 4 floats");
        return -1;
    }
    const size_t n4 = (size_t)(n_floats / 4), per = (n4 + (size_t)world - 1) / (size_t)world;
    const size_t begin = per * (size_t)rank < n4 ? per * (size_t)rank : n4;
    const size_t end = begin + per < n4 ? begin + per : n4;
    if (end > begin) {
        const size_t want = (end - begin + 511) / 512;
        // 2 CTAs per SM; measured on 2 and 8 B200: 148 ... 2368 CTAs all within 5 % (the switch is the limit)
        const unsigned blocks = (unsigned)(want < 148u * 2u ? want : 148u * 2u);
        nvls_allreduce_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float4*>(multicast_ptr), begin, end);
        EOGS_LAUNCH_CHECK("nvls_allreduce_kernel");
    }
    return 0;
}

EOGS_API int eogs_p2p_allreduce(eogs_stream_t stream, void* const* buffer_ptrs, unsigned long long n_floats, int rank, int world)
{
    if (!buffer_ptrs) { set_error("no peer buffer pointers: use the NCCL all-reduce"); return -4; }
    if (world <= 0 || world > 8 || rank < 0 || rank >= world || (n_floats & 3ull)) {
        set_error("bad rank/world (1..8 ranks), or the bucket is not a multiple of 4 floats");
        return -1;
    }
    PeerPtrs pp{};
    for (int r = 0; r < world; r++) {
        if (!buffer_ptrs[r] || (reinterpret_cast<uintptr_t>(buffer_ptrs[r]) & 15u)) { set_error("peer buffer %d is null or not 16-byte aligned", r); return -1; }
        pp.p[r] = static_cast<float4*>(buffer_ptrs[r]);
    }
    const size_t n4 = (size_t)(n_floats / 4), per = (n4 + (size_t)world - 1) / (size_t)world;
    const size_t begin = per * (size_t)rank < n4 ? per * (size_t)rank : n4;
    const size_t end = begin + per < n4 ? begin + per : n4;
    if (end > begin) {
        const size_t want = (end - begin + 511) / 512;
        const unsigned blocks = (unsigned)(want < 148u * 2u ? want : 148u * 2u);
        p2p_allreduce_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(stream)>>>(pp, world, rank, begin, end);
        EOGS_LAUNCH_CHECK("p2p_allreduce_kernel");
    }
    return 0;
}

}  // extern "C"
