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

}  // extern "C"
