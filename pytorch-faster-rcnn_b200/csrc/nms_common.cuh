// nms_common.cuh -- the exact suppression decision shared by nms.cu (K4) and rpn_back.cu (fused K4 + merge).
#pragma once
#include <math.h>

#include "common.cuh"

namespace b2d {

// (double)iou > thr with iou = inter / ((aa + ab) - inter), evaluated exactly like the
// reference.  The IEEE divide is only executed inside the narrow band |iou/thr - 1| <
// 2^-18 where the cheap products cannot decide (or for NaN/inf inputs and thr <= 0, where the
// host makes the bounds +-inf): outside it the sign of inter - thr*u already determines
// fl(inter/u) > thr (margin >> half an ulp).
__device__ __forceinline__ bool suppresses(const float4& a, float aa, const float4& b, float ab, float thr,
                                           float thr_lo, float thr_hi) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
    const float inter = w * h;
    const float u = (aa + ab) - inter;
    const bool yes = inter > thr_hi * u, no = inter < thr_lo * u;
    // the product test is only decisive for a finite positive union (malformed boxes with
    // x2 < x1 give u <= 0; the reference then compares the signed / NaN quotient)
    if ((yes || no) && u > 0.0f && u < 3.0e38f) return yes;
    return inter / u > thr;
}

// Same decision when both boxes are well-formed (x2 >= x1, y2 >= y1, no NaN): then 0 <= inter <=
// min(aa, ab), so u >= 0, and u = +inf / NaN / 0 all fall through to the exact quotient or give
// the same answer as it (see suppresses) -- the two range tests on u are not needed.
__device__ __forceinline__ bool suppresses_wf(const float4& a, float aa, const float4& b, float ab, float thr,
                                              float thr_lo, float thr_hi) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
    const float inter = w * h;
    const float u = (aa + ab) - inter;
    const bool yes = inter > thr_hi * u, no = inter < thr_lo * u;
    if (yes || no) return yes;
    return inter / u > thr;
}
// (sides <= 1e18 keep thr * u far from fp32 overflow, where the product test would stop being decisive)
__device__ __forceinline__ bool well_formed(const float4& b) {
    return b.z >= b.x && b.w >= b.y && (b.z - b.x) <= 1.0e18f && (b.w - b.y) <= 1.0e18f;
}

// Padding box for rows/columns past n: far away from anything finite a caller passes, with a
// finite positive area, so a real-vs-pad pair takes the cheap "no" exit.
__device__ __forceinline__ float4 pad_box() { return make_float4(-2.0e18f, -2.0e18f, -1.0e18f, -1.0e18f); }
__device__ __forceinline__ float area_of(const float4& b) { return (b.z - b.x) * (b.w - b.y); }


// decisive bounds around the threshold (see suppresses): outside [thr_lo, thr_hi] * u the products decide
static inline void nms_thr_bounds(float thr, float& t, float& lo, float& hi) {
    t = thr;
    if (thr > 0.0f && thr < 3.0e38f) {
        hi = thr * (1.0f + 1.0f / 262144.0f);
        lo = thr * (1.0f - 1.0f / 262144.0f);
    } else {                                         // thr <= 0 / inf / NaN: always take the exact divide
        hi = INFINITY;
        lo = -INFINITY;
    }
}

}  // namespace b2d
