// Gaussian target geometry shared by the render kernel (rr_render.cu) and the fused render+focal kernels
// (rr_focal.cu).  Every operation is an explicit round-to-nearest intrinsic, so the result does not depend
// on the translation unit's --fmad setting.  Internal header.
#pragma once
#include "rr_common.cuh"

namespace rr {

// functional.py:177-198 with min_overlap = 0.7; python scalars enter the tensor ops as fp32
__device__ __forceinline__ float gaussian_radius_f32(float height, float width) {
    const float c_1m = (float)(1 - 0.7), c_1p = (float)(1 + 0.7);
    const float b1 = __fadd_rn(height, width);
    const float c1 = __fdiv_rn(__fmul_rn(__fmul_rn(width, height), c_1m), c_1p);
    const float sq1 = __fsqrt_rn(__fsub_rn(__fmul_rn(b1, b1), __fmul_rn(4.0f, c1)));
    const float r1 = __fmul_rn(__fadd_rn(b1, sq1), 0.5f);
    const float b2 = __fmul_rn(2.0f, __fadd_rn(height, width));
    const float c2 = __fmul_rn(__fmul_rn(c_1m, width), height);
    const float sq2 = __fsqrt_rn(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.0f, c2)));
    const float r2 = __fmul_rn(__fadd_rn(b2, sq2), 0.5f);
    const float a3x4 = (float)(4 * (4 * 0.7));
    const float b3 = __fmul_rn((float)(-2 * 0.7), __fadd_rn(height, width));
    const float c3 = __fmul_rn(__fmul_rn((float)(0.7 - 1), width), height);
    const float sq3 = __fsqrt_rn(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(a3x4, c3)));
    const float r3 = __fmul_rn(__fadd_rn(b3, sq3), 0.5f);
    return fminf(fminf(r1, r2), r3);
}

// One object's Gaussian as draw_umich_gaussian sees it (functional.py:212-227, 230-262).
struct ObjGauss {
    float cxi, cyi, denom;      // integer centre (as float), 2*sigma*sigma
    int xa, xb, ya, yb;         // clipped window [xa,xb) x [ya,yb); empty when xb <= xa or yb <= ya
    int cls;                    // plane index, -1: nothing to draw
    float bw, bh, ox, oy, msk, idx;   // side outputs: wh, offset, reg_mask, ind
};

__device__ __forceinline__ ObjGauss obj_gauss(const float* __restrict__ a, int img_w, int Hh, int Wh, float sf, int cls_num) {
    ObjGauss o;
    float x1 = a[0], y1 = a[1];
    float x2 = __fadd_rn(a[2], a[0]), y2 = __fadd_rn(a[3], a[1]);                 // :246-247
    x1 = __fdiv_rn(x1, sf); y1 = __fdiv_rn(y1, sf); x2 = __fdiv_rn(x2, sf); y2 = __fdiv_rn(y2, sf);
    o.bh = __fsub_rn(y2, y1); o.bw = __fsub_rn(x2, x1);                           // :250
    const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f), cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
    o.cxi = floorf(cx); o.cyi = floorf(cy);
    o.ox = __fsub_rn(cx, o.cxi); o.oy = __fsub_rn(cy, o.cyi);
    o.msk = (o.bh > 0.f && o.bw > 0.f) ? 1.f : 0.f;
    o.idx = __fadd_rn(__fmul_rn(o.cyi, (float)(img_w / 4)), o.cxi);               // :257 hard-coded 4
    const float rad = fmaxf(floorf(gaussian_radius_f32(ceilf(o.bh), ceilf(o.bw))), 0.f);      // :258-259
    int cls = (int)__fsub_rn(a[5], 1.f);
    if (cls < 0) cls += cls_num;                                                  // python negative index
    o.cls = (cls >= 0 && cls < cls_num) ? cls : -1;
    const float sigma = __fdiv_rn(__fadd_rn(__fmul_rn(2.f, rad), 1.f), 6.f);
    o.denom = __fmul_rn(__fmul_rn(2.f, sigma), sigma);
    const float left = fminf(o.cxi, rad), right = fminf((float)Wh - o.cxi, rad + 1.f);
    const float top = fminf(o.cyi, rad), bottom = fminf((float)Hh - o.cyi, rad + 1.f);
    o.ya = (int)(o.cyi - top); o.yb = min((int)(o.cyi + bottom), Hh);
    o.xa = (int)(o.cxi - left); o.xb = min((int)(o.cxi + right), Wh);
    if (o.ya < 0 || o.xa < 0 || o.yb <= o.ya || o.xb <= o.xa) { o.xb = o.xa = 0; o.yb = o.ya = 0; }
    return o;
}

__device__ __forceinline__ float obj_value(const ObjGauss& o, int y, int x) {
    const float dx = (float)x - o.cxi, dy = (float)y - o.cyi;
    return expf(-__fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), o.denom));
}

}  // namespace rr
