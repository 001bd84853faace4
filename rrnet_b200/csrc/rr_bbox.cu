// Stage-2 box decode for sm_100a.  Replaces RRNetOperator.generate_bbox
// (operators/rrnet_operator.py:188-209) for every row of the batch in one launch.
// Built with --fmad=false: each torch op in the reference rounds once, e.g.
// ctr = (reg*w' + X1) + w'/2 is three roundings.
#include "rr_common.cuh"

namespace rr {

__global__ void __launch_bounds__(256)
generate_bbox_kernel(const float* __restrict__ bxyxy, const float* __restrict__ reg,
                     const float* __restrict__ scores, const float* __restrict__ clses,
                     const int* __restrict__ n_rois_dev, int n_cap, float scale,
                     float* __restrict__ s1, float* __restrict__ s2) {
    RR_PDL_PROLOGUE();
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < live; i += gridDim.x * blockDim.x) {
        const float* r = bxyxy + (size_t)i * 5;
        const float X1 = __fmul_rn(r[1], scale), Y1 = __fmul_rn(r[2], scale);     // :192
        const float X2 = __fmul_rn(r[3], scale), Y2 = __fmul_rn(r[4], scale);
        const float w = __fsub_rn(X2, X1), h = __fsub_rn(Y2, Y1);                 // :197
        const float sc = scores[i];
        float* a = s1 + (size_t)i * 6;
        a[0] = X1; a[1] = Y1; a[2] = w; a[3] = h; a[4] = sc; a[5] = 0.f;          // :198
        const float w1 = __fadd_rn(w, 1.f), h1 = __fadd_rn(h, 1.f);               // :201
        const float4 g = reinterpret_cast<const float4*>(reg)[i];
        const float cx = __fadd_rn(__fadd_rn(__fmul_rn(g.x, w1), X1), __fmul_rn(w1, 0.5f));   // :202
        const float cy = __fadd_rn(__fadd_rn(__fmul_rn(g.y, h1), Y1), __fmul_rn(h1, 0.5f));   // :203
        const float ow = __fmul_rn(expf(g.z), w1), oh = __fmul_rn(expf(g.w), h1);             // :204-205
        float* q = s2 + (size_t)i * 6;
        q[0] = __fsub_rn(cx, __fmul_rn(ow, 0.5f));                                 // :206
        q[1] = __fsub_rn(cy, __fmul_rn(oh, 0.5f));                                 // :207
        q[2] = ow; q[3] = oh; q[4] = sc;
        q[5] = __fadd_rn(clses[i], 1.f);                                           // :208
    }
}

int generate_bbox_launch(const float* bxyxy, const float* reg, const float* scores, const float* clses,
                         const int32_t* n_rois_dev, int n_cap, float scale, float* s1, float* s2,
                         cudaStream_t st) {
    int rc = 0;
    int grid = (n_cap + 255) / 256;
    launch_pdl(generate_bbox_kernel, dim3(grid), dim3(256), 0, st, bxyxy, reg, scores, clses, n_rois_dev, n_cap, scale, s1, s2);
    RR_LAUNCHED_K(rc, "generate_bbox_kernel", st);
    return rc;
}

}  // namespace rr

using namespace rr;

RR_API int rr_generate_bbox(const float* bxyxy, const float* reg, const float* scores, const float* clses,
                            const int32_t* n_rois_dev, int n_cap, float scale,
                            float* s1, float* s2, void* stream) {
    if (n_cap == 0) return 0;
    if (!bxyxy || !reg || !scores || !clses || !s1 || !s2 || n_cap < 0) return RR_E_BADARG;
    if ((uintptr_t)reg & 15) return RR_E_ALIGN;
    return generate_bbox_launch(bxyxy, reg, scores, clses, n_rois_dev, n_cap, scale, s1, s2, (cudaStream_t)stream);
}
