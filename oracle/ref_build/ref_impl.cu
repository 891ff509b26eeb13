// Build wrapper for the reference's rasterizer_impl.cu (oracle/_ref only; TEST INFRASTRUCTURE).
// Same in_frustum patch as ref_forward.cu; <cstdint> is needed by rasterizer_impl.h on gcc 13.
#include <cstdint>
#include <cstddef>
#include <stdexcept>
#include <cuda_runtime.h>
// CUB must come before config.h: its histogram templates have a parameter called NUM_CHANNELS,
// which config.h #defines (the reference includes CUB first too, rasterizer_impl.cu:19-20).
#include <cub/cub.cuh>
#include <cub/device/device_radix_sort.cuh>
#define in_frustum in_frustum_reference_ub
#include "cuda_rasterizer/auxiliary.h"
#undef in_frustum
__forceinline__ __device__ bool in_frustum(int idx, const float* orig_points, const float* viewmatrix,
                                           const float* projmatrix, bool prefiltered, float3& p_view)
{
    float3 p_orig = { orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2] };
    p_view = transformPoint4x3(p_orig, viewmatrix);
    return true;
}
#include "cuda_rasterizer/rasterizer_impl.cu"
