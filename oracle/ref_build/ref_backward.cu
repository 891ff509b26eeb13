// Build wrapper for the reference's backward.cu (oracle/_ref only; TEST INFRASTRUCTURE).
#include <cstdint>
#include <cuda_runtime.h>
#include "cuda_rasterizer/backward.cu"
