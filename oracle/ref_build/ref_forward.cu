// Build wrapper for the reference's forward.cu (oracle/_ref only; TEST INFRASTRUCTURE).
// The reference source is compiled from where it lies under /root/reference; nothing is copied.
//
// The one semantic patch the reference needs on nvcc 12.9: in_frustum
// (DGR/cuda_rasterizer/auxiliary.h:151-176) falls off its end without a return
// statement, which is UB and makes nvcc emit preprocessCUDA as an empty kernel.
// We rename the UB original out of the way and supply the only reading under which
// the rest of preprocessCUDA is reachable: compute p_view, return true.
#include <cstdint>
#include <cuda_runtime.h>
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
#include "cuda_rasterizer/forward.cu"
