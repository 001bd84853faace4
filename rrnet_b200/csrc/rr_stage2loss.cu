// Stage-2 regression loss for sm_100a: the per-image Python loop of RRNetOperator.criterion
// (operators/rrnet_operator.py:64-84) -- box_iou against the padded ground truth, max over the ground truth, IoU > 0.5
// selection with the "row 0, factor 0" rule when an image has no positive, generate_bbox_target (:86-102, "+1" widths),
// smooth_l1_loss(mean) / bs -- and its backward, one CTA per image, no host sync (the reference synchronises twice per
// image: `pos_idx.sum() == 0` and the boolean-mask indexing).
//
// Gradients: d loss / d s2_reg, and d loss / d bxyxy through the targets (the reference does NOT detach the predicted
// boxes the targets are built from, Appendix A.5), both scaled by `grad_scale`.  The IoU / max / threshold selection is
// piecewise constant, as in autograd.
#include "rr_common.cuh"

namespace rr {

constexpr int kS2Threads = 512;

__device__ __forceinline__ double block_sum_s2(double v, double* s_red) {
    v = warp_sum(v);
    if (lane_id() == 0) s_red[warp_id()] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kS2Threads / 32; ++i) t += s_red[i];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kS2Threads)
stage2_loss_kernel(const float* __restrict__ bxyxy, const int* __restrict__ seg, const float* __restrict__ reg,
                   const float* __restrict__ gt, int max_n, int gt_stride, float scale, float inv_bs, float grad_scale,
                   float* __restrict__ loss_parts, float* __restrict__ grad_reg, float* __restrict__ grad_box) {
    extern __shared__ float s_gt[];                  // [max_n][4] ground truth xyxy of this image
    __shared__ double s_red[kS2Threads / 32];
    __shared__ int s_npos;
    const int b = blockIdx.x;
    const int lo = seg[b], hi = seg[b + 1];
    for (int i = threadIdx.x; i < max_n * 4; i += kS2Threads)
        s_gt[i] = gt[((size_t)b * max_n + (i >> 2)) * gt_stride + (i & 3)];
    if (threadIdx.x == 0) s_npos = 0;
    __syncthreads();

    // ---- pass 1: best ground-truth box per RoI, positives of this image (rrnet_operator.py:72-74) ----
    // the row maxima of the first kCache strides of a thread stay in registers for pass 2
    constexpr int kCache = 4;                        // 4 x 512 = 2048 RoIs per image without recomputation
    float c_best[kCache];
    int c_arg[kCache];
    auto row_max = [&](int r, float& best, int& arg) {
        const float* bx = bxyxy + (size_t)r * 5 + 1;
        const float x1 = __fmul_rn(bx[0], scale), y1 = __fmul_rn(bx[1], scale);
        const float x2 = __fmul_rn(bx[2], scale), y2 = __fmul_rn(bx[3], scale);
        const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
        best = -1.f;
        arg = 0;
        bool nan_seen = false;
        for (int k = 0; k < max_n; ++k) {
            const float gx1 = s_gt[4 * k], gy1 = s_gt[4 * k + 1], gx2 = s_gt[4 * k + 2], gy2 = s_gt[4 * k + 3];
            const float iw = fmaxf(__fsub_rn(fminf(x2, gx2), fmaxf(x1, gx1)), 0.f);
            const float ih = fmaxf(__fsub_rn(fminf(y2, gy2), fmaxf(y1, gy1)), 0.f);
            const float inter = __fmul_rn(iw, ih);
            const float ga = __fmul_rn(__fsub_rn(gx2, gx1), __fsub_rn(gy2, gy1));
            const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area, ga), inter));
            nan_seen |= iou != iou;                               // 0/0: zero-area RoI against a zero-padded gt row
            if (iou > best) { best = iou; arg = k; }              // first maximum, like torch.max
        }
        // torch.max propagates NaN: the reference's max_iou of such a row is NaN and `NaN > 0.5` makes it a negative
        // (rrnet_operator.py:73-74).  Same decision here; `arg` of a negative row is never used.
        if (nan_seen) best = __int_as_float(0x7fc00000);
    };
    int n_pos = 0;
    {
        int it = 0;
        for (int r = lo + threadIdx.x; r < hi; r += kS2Threads, ++it) {
            float best;
            int arg;
            row_max(r, best, arg);
#pragma unroll
            for (int q = 0; q < kCache; ++q)
                if (q == it) { c_best[q] = best; c_arg[q] = arg; }
            n_pos += best > 0.5f;
        }
    }
    n_pos = (int)block_sum_s2((double)n_pos, s_red);

    // ---- pass 2: targets, smooth-L1, gradients ----
    const float w = n_pos > 0 ? __fdiv_rn(inv_bs, (float)(4 * n_pos)) : 0.f;      // mean over n_pos x 4, then / bs
    double sum = 0.0;
    int it2 = 0;
    for (int r = lo + threadIdx.x; r < hi; r += kS2Threads, ++it2) {
        const float* bx = bxyxy + (size_t)r * 5 + 1;
        const float x1 = __fmul_rn(bx[0], scale), y1 = __fmul_rn(bx[1], scale);
        const float x2 = __fmul_rn(bx[2], scale), y2 = __fmul_rn(bx[3], scale);
        float best = -1.f;
        int arg = 0;
        if (it2 < kCache) {
#pragma unroll
            for (int q = 0; q < kCache; ++q)
                if (q == it2) { best = c_best[q]; arg = c_arg[q]; }
        } else {
            row_max(r, best, arg);
        }
        float4 g_reg = make_float4(0.f, 0.f, 0.f, 0.f), g_box = g_reg;
        if (n_pos > 0 && best > 0.5f) {
            const float gx1 = s_gt[4 * arg], gy1 = s_gt[4 * arg + 1], gx2 = s_gt[4 * arg + 2], gy2 = s_gt[4 * arg + 3];
            const float ew = x2 - x1 + 1.0f, eh = y2 - y1 + 1.0f;                  // :88-95
            const float ecx = x1 + 0.5f * ew, ecy = y1 + 0.5f * eh;
            const float gw = gx2 - gx1 + 1.0f, gh = gy2 - gy1 + 1.0f;
            const float gcx = gx1 + 0.5f * gw, gcy = gy1 + 0.5f * gh;
            const float t[4] = {(gcx - ecx) / ew, (gcy - ecy) / eh, logf(gw / ew), logf(gh / eh)};
            const float4 p = reinterpret_cast<const float4*>(reg)[r];
            const float d[4] = {p.x - t[0], p.y - t[1], p.z - t[2], p.w - t[3]};
            float gd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = fabsf(d[j]);
                sum += (double)(a < 1.0f ? 0.5f * d[j] * d[j] : a - 0.5f);         // smooth_l1, beta = 1
                gd[j] = (a < 1.0f ? d[j] : (d[j] > 0.f ? 1.0f : -1.0f)) * w * grad_scale;
            }
            g_reg = make_float4(gd[0], gd[1], gd[2], gd[3]);
            // targets are functions of the predicted box: d loss / d t = -gd
            const float a_x = -gd[0], a_y = -gd[1], c_w = -gd[2], c_h = -gd[3];
            const float g_ecx = -a_x / ew, g_ew = -a_x * t[0] / ew - c_w / ew;
            const float g_ecy = -a_y / eh, g_eh = -a_y * t[1] / eh - c_h / eh;
            g_box.x = (0.5f * g_ecx - g_ew) * scale;
            g_box.z = (0.5f * g_ecx + g_ew) * scale;
            g_box.y = (0.5f * g_ecy - g_eh) * scale;
            g_box.w = (0.5f * g_ecy + g_eh) * scale;
        }
        if (grad_reg) reinterpret_cast<float4*>(grad_reg)[r] = g_reg;
        if (grad_box) reinterpret_cast<float4*>(grad_box)[r] = g_box;
    }
    sum = block_sum_s2(sum, s_red);
    if (threadIdx.x == 0) loss_parts[b] = n_pos > 0 ? (float)(sum * (double)w) : 0.f;
}

}  // namespace rr

using namespace rr;

RR_API int rr_stage2_loss(const float* bxyxy, const int32_t* seg_offsets, const float* s2_reg, const float* gt_xyxy,
                          int B, int max_n, int gt_stride, float scale, float grad_scale,
                          float* loss_parts, float* grad_reg, float* grad_box, void* stream) {
    if (!bxyxy || !seg_offsets || !s2_reg || !loss_parts || B <= 0 || max_n < 0 || gt_stride < 4) return RR_E_BADARG;
    if (max_n > 0 && !gt_xyxy) return RR_E_BADARG;
    if (((uintptr_t)s2_reg & 15) || ((uintptr_t)grad_reg & 15) || ((uintptr_t)grad_box & 15)) return RR_E_ALIGN;
    const size_t smem = sizeof(float) * 4 * (size_t)max_n;
    if (smem > 200 * 1024) return RR_E_BADARG;
    int rc = 0;
    static OncePerDevice attr_once; int attr_dev;
    if (attr_once.need(&attr_dev) && smem > 48 * 1024) {
        RR_CUDA(cudaFuncSetAttribute(stage2_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), rc);
        if (rc == 0) attr_once.mark(attr_dev);
    }
    stage2_loss_kernel<<<B, kS2Threads, smem, (cudaStream_t)stream>>>(bxyxy, seg_offsets, s2_reg, gt_xyxy, max_n, gt_stride,
                                                                      scale, 1.0f / (float)B, grad_scale, loss_parts,
                                                                      grad_reg, grad_box);
    RR_LAUNCHED_K(rc, "stage2_loss_kernel", (cudaStream_t)stream);
    return rc;
}
