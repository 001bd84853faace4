// True-positive matching of the evaluation for sm_100a: utils/metrics/metrics.py:51-136 (`get_tp`,
// with `bbox_iou` :10-47) for a batch of images, one CTA per image.  The reference does this on the CPU with an
// [m, n, T] IoU tensor and a Python loop over classes and detections; here the IoUs are recomputed on the fly:
//   1. ground truth inside an ignore region (class 0, overlap >= 0.5 of the box's own area) is dropped (:74-79),
//   2. detections are ranked by score (rank sort), those inside an ignore region are dropped (:82-88),
//   3. per (class, IoU threshold) pair one warp walks the detections best-first and gives each the unused ground-truth box
//      of its class with the largest IoU >= threshold (first maximum, like torch.max) (:105-134),
//   4. a detection is emitted only if the image holds ground truth of its class (:113-114).
// IoU arithmetic is the reference's, op for op in fp32 (built with --fmad=false), so the flags are bit-exact.
#include "rr_common.cuh"

namespace rr {

constexpr int kApThreads = 1024;
constexpr int kApMax = 768;             // detections / ground-truth boxes per image (the reference caps both at 500)
constexpr int kApMaxT = 16, kApMaxCls = 32;

struct ApBox { float x1, y1, x2, y2, area; };

__device__ __forceinline__ ApBox ap_box(const float* r) {          // bbox_iou, x1y1x2y2=False: :22-29
    ApBox b;
    b.x1 = r[0]; b.y1 = r[1];
    b.x2 = __fadd_rn(r[2], r[0]); b.y2 = __fadd_rn(r[3], r[1]);
    b.area = __fmul_rn(__fsub_rn(b.x2, b.x1), __fsub_rn(b.y2, b.y1));
    return b;
}
__device__ __forceinline__ float ap_inter(const ApBox& a, const ApBox& b) {   // :31-35
    const float iw = fmaxf(__fsub_rn(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.f);
    return __fmul_rn(iw, ih);
}

__global__ void __launch_bounds__(kApThreads)
ap_match_kernel(const float* __restrict__ pred, const int* __restrict__ n_pred, const float* __restrict__ target,
                const int* __restrict__ n_tgt, const float* __restrict__ thresholds, int M, int N, int T, int cls_num,
                int* __restrict__ order, float* __restrict__ tp, int* __restrict__ out_cls,
                float* __restrict__ target_count, float* __restrict__ in_img) {
    __shared__ ApBox s_g[kApMax];                  // ground truth
    __shared__ int s_gcls[kApMax];                 // class, -1 = dropped
    __shared__ ApBox s_d[kApMax];                  // detections in rank order
    __shared__ int s_dcls[kApMax];                 // class, -1 = dropped / not emitted
    __shared__ float s_score[kApMax];
    __shared__ int s_cnt[kApMaxCls];
    __shared__ int s_any_ignore;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = min(n_pred[b], M), n = min(n_tgt[b], N);
    const float* P = pred + (size_t)b * M * 6;
    const float* G = target + (size_t)b * N * 6;
    if (tid < kApMaxCls) s_cnt[tid] = 0;
    if (tid == 0) s_any_ignore = 0;
    __syncthreads();
    for (int j = tid; j < n; j += kApThreads) {
        s_g[j] = ap_box(G + (size_t)j * 6);
        const int c = (int)G[(size_t)j * 6 + 5];
        s_gcls[j] = c;
        if (c == 0) s_any_ignore = 1;
    }
    for (int i = tid; i < m; i += kApThreads) s_score[i] = P[(size_t)i * 6 + 4];
    __syncthreads();
    const bool any_ignore = s_any_ignore != 0;

    // ---- 1. ground truth inside ignore regions (uses the unfiltered list, like the reference) ----
    int drop_g[(kApMax + kApThreads - 1) / kApThreads];
#pragma unroll
    for (int k = 0; k < (kApMax + kApThreads - 1) / kApThreads; ++k) {
        const int j = tid + k * kApThreads;
        drop_g[k] = 0;
        if (any_ignore && j < n && s_gcls[j] != 0) {
            float best = -INFINITY;
            bool nan = false;
            for (int q = 0; q < n; ++q) {
                if (s_gcls[q] != 0) continue;
                const float ov = __fdiv_rn(ap_inter(s_g[j], s_g[q]), s_g[j].area);     // overlap of the box in the region
                nan |= ov != ov;
                best = fmaxf(best, ov);
            }
            drop_g[k] = nan || !(best < 0.5f);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < (kApMax + kApThreads - 1) / kApThreads; ++k) {
        const int j = tid + k * kApThreads;
        if (j < n) {
            if (drop_g[k]) s_gcls[j] = -1;
            else if (s_gcls[j] > 0 && s_gcls[j] < kApMaxCls) atomicAdd(&s_cnt[s_gcls[j]], 1);
        }
    }

    // ---- 2. detections: rank by score (descending, ties by index), drop those inside ignore regions ----
    for (int i = tid; i < m; i += kApThreads) {
        const float sc = s_score[i];
        int rank = 0;
        for (int q = 0; q < m; ++q) {
            const float o = s_score[q];
            rank += (o > sc) || (o == sc && q < i);
        }
        const ApBox d = ap_box(P + (size_t)i * 6);
        int c = (int)P[(size_t)i * 6 + 5];
        if (any_ignore) {
            float best = -INFINITY;
            bool nan = false;
            for (int q = 0; q < n; ++q) {
                if (s_gcls[q] != 0) continue;          // ignore regions are never dropped in step 1
                const float ov = __fdiv_rn(ap_inter(d, s_g[q]), d.area);
                nan |= ov != ov;
                best = fmaxf(best, ov);
            }
            if (nan || !(best < 0.5f)) c = -1;
        }
        s_d[rank] = d;
        s_dcls[rank] = c;
        order[(size_t)b * M + rank] = i;
    }
    for (int i = m + tid; i < M; i += kApThreads) { order[(size_t)b * M + i] = -1; out_cls[(size_t)b * M + i] = -1; }
    for (int i = tid; i < M * T; i += kApThreads) tp[(size_t)b * M * T + i] = 0.f;
    __syncthreads();
    if (tid >= 1 && tid < cls_num) {
        target_count[(size_t)b * (cls_num - 1) + tid - 1] = (float)s_cnt[tid];
        in_img[(size_t)b * (cls_num - 1) + tid - 1] = s_cnt[tid] != 0 ? 1.f : 0.f;
    }
    // ---- 4. (first, it only needs the counts) a detection is emitted only with ground truth of its class in the image ----
    for (int p = tid; p < m; p += kApThreads) {
        const int c = s_dcls[p];
        out_cls[(size_t)b * M + p] = (c >= 1 && c < cls_num && s_cnt[c] > 0) ? c : -1;
    }

    // ---- 3. greedy matching: one warp per (class, threshold) pair ----
    for (int pair = warp; pair < (cls_num - 1) * T; pair += kApThreads / 32) {
        const int c = 1 + pair / T, t = pair - (c - 1) * T;
        if (s_cnt[c] == 0) continue;
        const float thr = thresholds[t];
        unsigned used = 0;                             // bit k: ground-truth box lane + 32 k is taken at this threshold
        for (int p = 0; p < m; ++p) {
            if (s_dcls[p] != c) continue;              // warp-uniform
            const ApBox d = s_d[p];
            float best = 0.f;
            int arg = 0x7fffffff;
            for (int k = 0; lane + 32 * k < n; ++k) {
                const int j = lane + 32 * k;
                if (s_gcls[j] != c || ((used >> k) & 1u)) continue;
                const ApBox g = s_g[j];
                const float inter = ap_inter(d, g);
                const float ua = fmaxf(__fsub_rn(__fadd_rn(d.area, g.area), inter), 1e-8f);      // :37-39
                const float iou = __fdiv_rn(inter, ua);
                if (__fsub_rn(iou, thr) >= 0.f && iou > best) { best = iou; arg = j; }           // :97; first maximum per lane
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {         // warp arg-max, ties to the smaller index
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (best > 0.f) {                          // max_iou.nonzero() :124
                if ((arg & 31) == lane) used |= 1u << (arg >> 5);
                if (lane == 0) tp[((size_t)b * M + p) * T + t] = 1.f;
            }
        }
    }
}

}  // namespace rr

using namespace rr;

RR_API int rr_ap_match(const float* pred, const int32_t* n_pred, const float* target, const int32_t* n_tgt,
                       const float* thresholds, int B, int M, int N, int T, int cls_num,
                       int32_t* order, float* tp, int32_t* out_cls, float* target_count, float* in_img, void* stream) {
    if (B <= 0 || M <= 0 || N < 0 || T <= 0 || cls_num < 2) return RR_E_BADARG;
    if (M > kApMax || N > kApMax || T > kApMaxT || cls_num > kApMaxCls) return RR_E_RANGE;
    if (!pred || !n_pred || !n_tgt || !thresholds || !order || !tp || !out_cls || !target_count || !in_img) return RR_E_BADARG;
    if (N > 0 && !target) return RR_E_BADARG;
    int rc = 0;
    ap_match_kernel<<<B, kApThreads, 0, (cudaStream_t)stream>>>(pred, n_pred, target, n_tgt, thresholds, M, N, T, cls_num,
                                                               order, tp, out_cls, target_count, in_img);
    RR_LAUNCHED_K(rc, "ap_match_kernel", (cudaStream_t)stream);
    return rc;
}
