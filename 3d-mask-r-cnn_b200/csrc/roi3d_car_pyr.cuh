// roi3d_car_pyr.cuh -- fused PyramidROIAlign routing shared by the crop-and-resize kernels (plane-staged and row-walk).
#pragma once
#include "roi3d_common.cuh"
#include <math.h>
#include <string.h>

namespace roi3d {

// ---- fused PyramidROIAlign -------------------------------------------------------
// In pyramid mode a launch serves all four levels: every CTA clips its ROI, routes it to a level with the
// reference's formula (PyramidROIAlign.call, core/models.py:615-649) and then runs the ordinary plane-staged crop on
// that level's map, writing the crop at the ROI's original position (no where / gather_nd / concat / top_k glue).
struct PyrParams {
    const float *image[4];      // P2..P5 (forward) -- or grad images (backward)
    int H[4], W[4], D[4];       // level shapes [B, H_l, W_l, D_l, C]
    float imH, imW, imD;        // image_shape from image_meta (routing and the z min-size rule)
    int rois_per_image;         // boxes are [B, R, 6]; box_index = roi / R
    float lt[3];                // smallest normalized box volume routed to level 3 / 4 / 5 (host: pyr_level_thresholds)
};

struct PyrRoute { float box[6]; int level; };

// The level expression of PyramidROIAlign.call (core/models.py:637-649), fp32 like the TF graph:
//   level = clamp(4 + round(log2(cbrt(h*w*d) / (224 / cbrt(H*W*D)))), 2, 5)
// It depends on the box only through vol = h*w*d and is monotonic in it, so the host turns it into three volume
// thresholds once per call (bisection over the fp32 values of vol, same formula) and a CTA routes its box with three
// compares instead of two powf and two logf on its critical path (round 1: ~0.4 us per CTA before anything else could
// start).  As before, a box whose expression lies within an ulp of x.5 may land on the other side than with another libm.
static inline int pyr_level_of(float vol, float imH, float imW, float imD) {
    const float area = (imH * imW) * imD;
    const float ratio = powf(vol, 1.0f / 3.0f) / (224.0f / powf(area, 1.0f / 3.0f));
    const float lvl = logf(ratio) / logf(2.0f);
    const int k = 4 + (int)nearbyintf(lvl);                   // tf.round: half to even
    return k < 2 ? 2 : (k > 5 ? 5 : k);
}
static inline void pyr_level_thresholds(float imH, float imW, float imD, float lt[3]) {
    for (int target = 3; target <= 5; ++target) {
        uint32_t lo = 0x00800000u, hi = 0x7f000000u;          // smallest normal .. huge: vol > 0 always (min sizes)
        while (lo < hi) {                                      // smallest vol whose level is >= target
            const uint32_t mid = lo + (hi - lo) / 2;
            float v;
            memcpy(&v, &mid, 4);
            if (pyr_level_of(v, imH, imW, imD) >= target) hi = mid; else lo = mid + 1;
        }
        memcpy(&lt[target - 3], &lo, 4);
    }
}

// tf.clip_by_value + min sizes (core/models.py:615-632), then the level from the host's thresholds
__device__ __forceinline__ PyrRoute pyr_route(const float *b6, const PyrParams &P) {
    PyrRoute r;
    float y1 = fminf(fmaxf(b6[0], 0.f), 1.f), x1 = fminf(fmaxf(b6[1], 0.f), 1.f), z1 = fminf(fmaxf(b6[2], 0.f), 1.f);
    float y2 = fminf(fmaxf(b6[3], 0.f), 1.f), x2 = fminf(fmaxf(b6[4], 0.f), 1.f), z2 = fminf(fmaxf(b6[5], 0.f), 1.f);
    y2 = fmaxf(y2, __fadd_rn(y1, 1e-6f));
    x2 = fmaxf(x2, __fadd_rn(x1, 1e-6f));
    z2 = fmaxf(z2, __fadd_rn(z1, __fdiv_rn(1.0f, fmaxf(P.imD, 1.0f))));
    r.box[0] = y1; r.box[1] = x1; r.box[2] = z1; r.box[3] = y2; r.box[4] = x2; r.box[5] = z2;
    const float vol = __fmul_rn(__fmul_rn(__fsub_rn(y2, y1), __fsub_rn(x2, x1)), __fsub_rn(z2, z1));
    r.level = 2 + (vol >= P.lt[0]) + (vol >= P.lt[1]) + (vol >= P.lt[2]);
    return r;
}

// level l of the pyramid parameters with constant indices only (a run-time index into a kernel parameter array makes
// the compiler copy the struct to local memory: round 1's PYR kernels carried a 96-byte stack frame for it)
__device__ __forceinline__ void pyr_level(const PyrParams &P, int lv, int &H, int &W, int &D, const float *&img) {
    H = lv == 0 ? P.H[0] : (lv == 1 ? P.H[1] : (lv == 2 ? P.H[2] : P.H[3]));
    W = lv == 0 ? P.W[0] : (lv == 1 ? P.W[1] : (lv == 2 ? P.W[2] : P.W[3]));
    D = lv == 0 ? P.D[0] : (lv == 1 ? P.D[1] : (lv == 2 ? P.D[2] : P.D[3]));
    img = lv == 0 ? P.image[0] : (lv == 1 ? P.image[1] : (lv == 2 ? P.image[2] : P.image[3]));
}

__device__ __forceinline__ float4 scrub4(const float4 v) {       // tf.where(is_finite(x), x, 0), core/models.py:683
    return make_float4(isfinite(v.x) ? v.x : 0.f, isfinite(v.y) ? v.y : 0.f, isfinite(v.z) ? v.z : 0.f, isfinite(v.w) ? v.w : 0.f);
}

__device__ __forceinline__ float4 sel4(bool bad, const float4 a, const float4 b) {
    return make_float4(bad ? a.x : b.x, bad ? a.y : b.y, bad ? a.z : b.z, bad ? a.w : b.w);
}

}  // namespace roi3d
