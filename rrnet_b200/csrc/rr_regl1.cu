// RegL1Loss of the wh / offset maps for sm_100a.  Replaces modules/loss/regl1loss.py:9-17 -- a
// permute + contiguous copy of the whole NCHW map, a gather, two expand/multiply passes and an l1_loss -- and its
// autograd graph (scatter through the gather, permute back) with one launch that reads the <= B*max_n*c gathered
// values straight from the NCHW map, and one that writes the (sparse) gradient:
//     pred[b,k,ch] = output[b,ch,ind[b,k]]
//     loss = sum |pred*mask - target*mask| / (c * sum(mask) + 1e-4)            (mask expanded over the c channels)
//     d loss / d output[b,ch,ind[b,k]] += sign(pred*mask - target*mask) * mask / (c * sum(mask) + 1e-4)
// The sums run in double over a fixed order (one CTA), so the loss is bit-reproducible.
#include "rr_common.cuh"

namespace rr {

constexpr int kRegThreads = 1024;

__device__ __forceinline__ double block_sum(double v, double* s_red) {     // fixed order; result in every thread
    v = warp_sum(v);
    if (lane_id() == 0) s_red[warp_id()] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kRegThreads / 32; ++i) t += s_red[i];
    __syncthreads();
    return t;
}

// grid = 1.  ind arrives as float (the collate pads it so, datasets/drones_det.py:70-94).
__global__ void __launch_bounds__(kRegThreads)
regl1_kernel(const float* __restrict__ output, const float* __restrict__ mask, const float* __restrict__ ind,
             const float* __restrict__ target, int B, int c, int HW, int max_n, float grad_scale,
             float* __restrict__ loss, float* __restrict__ grad) {
    __shared__ double s_red[kRegThreads / 32];
    const int n = B * max_n * c;
    double sum_abs = 0.0, sum_mask = 0.0;
    for (int i = threadIdx.x; i < n; i += kRegThreads) {
        const int ch = i % c, bk = i / c, b = bk / max_n;
        const float m = mask[bk];
        const int at = (int)ind[bk];
        float d = 0.f;
        if (at >= 0 && at < HW) {
            const float p = output[((size_t)b * c + ch) * HW + at];
            d = __fsub_rn(__fmul_rn(p, m), __fmul_rn(target[i], m));        // regl1loss.py:15
        }
        sum_abs += (double)fabsf(d);
        sum_mask += (double)m;                                               // expand_as(pred): every channel counts
    }
    sum_abs = block_sum(sum_abs, s_red);
    sum_mask = block_sum(sum_mask, s_red);
    const float denom = __fadd_rn((float)sum_mask, 1e-4f);                   // :16
    if (threadIdx.x == 0) loss[0] = __fdiv_rn((float)sum_abs, denom);
    if (grad == nullptr) return;
    const float g = __fdiv_rn(grad_scale, denom);
    for (int i = threadIdx.x; i < n; i += kRegThreads) {
        const int ch = i % c, bk = i / c, b = bk / max_n;
        const float m = mask[bk];
        const int at = (int)ind[bk];
        if (m == 0.f || at < 0 || at >= HW) continue;
        const size_t o = ((size_t)b * c + ch) * HW + at;
        const float d = __fsub_rn(__fmul_rn(output[o], m), __fmul_rn(target[i], m));
        const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        if (s != 0.f) atomicAdd(grad + o, __fmul_rn(__fmul_rn(s, m), g));    // two objects may share a centre cell
    }
}

}  // namespace rr

using namespace rr;

RR_API int rr_regl1_fwd_bwd(const float* output, const float* mask, const float* ind, const float* target,
                            int B, int c, int H, int W, int max_n, float grad_scale,
                            float* loss, float* grad, void* stream) {
    if (!output || !loss || B <= 0 || c <= 0 || H <= 0 || W <= 0 || max_n < 0) return RR_E_BADARG;
    if (max_n > 0 && (!mask || !ind || !target)) return RR_E_BADARG;
    if ((long long)B * max_n * c > 0x7fffffffLL) return RR_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (grad) RR_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)B * c * H * W, st), rc);
    regl1_kernel<<<1, kRegThreads, 0, st>>>(output, mask, ind, target, B, c, H * W, max_n, grad_scale, loss, grad);
    RR_LAUNCHED_K(rc, "regl1_kernel", st);
    return rc;
}
