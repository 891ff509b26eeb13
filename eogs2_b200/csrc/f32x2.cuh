// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2 etc., sm_100+).
//
// One FFMA2 performs two IEEE round-to-nearest FMAs on a 64-bit register pair for ONE issue slot.
// The blend kernels are issue-bound (ncu: smsp__issue_active 70-80 %, fma pipe the busiest), and a
// lane that owns two pixels runs the same arithmetic on both — so the per-pair math is written on
// pixel PAIRS.  Measured on B200 (tools/microbench/ffma2.cu, profiles/r01c_ffma2_microbench.txt):
// FFMA2 sustains the same 128 FMA/clk/SM as scalar FFMA with half the instructions.
// ptxas folds scalar broadcasts (R.F32), negation and immediates into the FFMA2 operands, so bc()
// and neg() cost nothing.  Each lane of a packed op rounds exactly like the scalar op; note that
// ptxas may contract mul2 followed by add2 into one FFMA2 — where bit-exactness matters (forward)
// the code only uses shapes that cannot be contracted (explicit fma2, add -> mul).
#pragma once
#include <cuda_runtime.h>

namespace eogs {

struct f2 { float2 v; };

__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; r.v = make_float2(lo, hi); return r; }
__device__ __forceinline__ f2 bc2(float a) { return mk2(a, a); }
__device__ __forceinline__ void un2(f2 x, float& lo, float& hi) { lo = x.v.x; hi = x.v.y; }
__device__ __forceinline__ float lo2(f2 x) { return x.v.x; }
__device__ __forceinline__ float hi2(f2 x) { return x.v.y; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
// acc = a * b + acc
__device__ __forceinline__ void fma2_acc(f2& acc, f2 a, f2 b) { acc.v = __ffma2_rn(a.v, b.v, acc.v); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-a.v.x, -a.v.y); }

}  // namespace eogs
