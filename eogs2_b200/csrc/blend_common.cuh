// Pieces shared by the forward and backward blend kernels: the exact culling test and the approximate MUFU wrappers.
//
// A 16x16 tile is cut into four 8x8 pixel REGIONS (bit q of a mask = region (8*(q&1), 8*(q>>1))).  The forward thread
// that stages a list entry decides ONCE which regions the Gaussian can reach with alpha >= 1/255 (region_mask below):
// the forward kernel (128 threads, warp = one region) appends the entry only to the regions it reaches and stores the
// mask, one byte per instance; the backward kernel (one warp per tile, the same pixel-to-lane map) reads it, never
// fetches entries nobody reaches and evaluates only the regions of the mask.
// The reference evaluates every entry of the tile list in every one of the 256 threads (forward.cu:356-372,
// backward.cu:551-577).
// Culling is exact and conservative: it only removes (pixel, Gaussian) pairs that the per-pixel test
// would reject, so images, n_contrib and gradients are unchanged; the tile lists stay the reference's.
#pragma once
#include "common.cuh"

namespace eogs {

constexpr int PATCH_W = 8, PATCH_H = 4;

// 4-bit mask of the 8x8 REGIONS of a tile (bit 2R + c: rows 8R..8R+7, columns 8c..8c+7) this Gaussian may contribute to.
//
// alpha >= 1/255 somewhere in a region  <=>  the ellipse E = { p : q(p - m) <= qmax },
// qmax = 2 ln(255 op), meets the region's rectangle.  E is cut by horizontal lines at the region-row
// boundaries: on a line at offset dy from the centre, E spans x in [c - h, c + h] with
// c = mx - (B/A) dy and h = sqrt(A qmax - det dy^2) / A.  Over a band between two lines the
// right edge c + h is concave in dy, so its maximum is the bounding-box extreme mx + hx when the
// extreme's dy = -(B/C) hx lies inside the band and otherwise sits on one of the two lines;
// same for the left edge.  One sqrt per line (3 lines for 2 bands, bands widened by half a pixel
// so neighbours share a line) gives all 4 answers exactly.  Margin 0.1 on q (0.05 on the exponent) covers rounding and
// the approximate sqrt / divide; NaNs compare false and keep the entry.
__device__ __forceinline__ float fast_sqrt(float x) {     // MUFU.SQRT: 2 ulp, inside region_mask's margin (an IEEE sqrt is ~8 instructions)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t region_mask(const float4& r0, const float4& r1, float tx0, float ty0,
                                                float img_x1, float img_y1) {
    constexpr int ROWS = 2, COLS = 2, RH = 8, RW = 8;
    const float mx = r0.x, my = r0.y, A = r0.z, B = r0.w, C = r1.x, op = r1.y;
    const float qmax = 2.f * __logf(255.f * op) + 0.1f;
    if (qmax <= 0.f) return 0u;
    const float det = A * C - B * B;
    const float inv_det = __fdividef(1.f, det), invA = __fdividef(1.f, A);
    const float hx = fast_sqrt(qmax * C * inv_det), hy = fast_sqrt(qmax * A * inv_det);
    if (mx + hx < tx0 || mx - hx > tx0 + (TILE - 1) || my + hy < ty0 || my - hy > ty0 + (TILE - 1)) return 0u;
    const float slope = -B * invA;
    const float dy_right = __fdividef(-B * hx, C), dy_left = -dy_right;
    const float Aq = A * qmax;
    float dyl[ROWS + 1], xr[ROWS + 1], xl[ROWS + 1];
#pragma unroll
    for (int k = 0; k <= ROWS; k++) {
        dyl[k] = (ty0 + (float)(RH * k) - 0.5f) - my;
        const float dyc = fminf(fmaxf(dyl[k], -hy), hy);
        const float h = fast_sqrt(fmaxf(0.f, Aq - det * dyc * dyc)) * invA;
        const float c = mx + slope * dyc;
        xr[k] = c + h;
        xl[k] = c - h;
    }
    uint32_t m = 0u;
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        if (dyl[r + 1] < -hy || dyl[r] > hy) continue;
        if (ty0 + (float)(RH * r) > img_y1) continue;
        const float right = (dy_right > dyl[r] && dy_right < dyl[r + 1]) ? mx + hx : fmaxf(xr[r], xr[r + 1]);
        const float left = (dy_left > dyl[r] && dy_left < dyl[r + 1]) ? mx - hx : fminf(xl[r], xl[r + 1]);
#pragma unroll
        for (int c = 0; c < COLS; c++) {
            const float px0 = tx0 + (float)(RW * c);
            if (px0 > img_x1) continue;
            if (!(right < px0 - 0.01f || left > px0 + (RW - 1) + 0.01f)) m |= 1u << (r * COLS + c);
        }
    }
    return m;
}

__device__ __forceinline__ float fast_rcp(float x) {      // x in [0.01, 1]: no denormal handling needed
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace eogs
