// Gaussian / linear / hard Soft-NMS on the device for sm_100a, batched over segments.
//
// Replaces ext.nms.nms_wrapper.soft_nms -> cpu_soft_nms (ext/nms/nms_wrapper.py:13-19,
// ext/nms/nms/cpu_nms.pyx:17-120), the O(n^2) serial host loop that
// RRNetOperator._ext_nms (operators/rrnet_operator.py:211-232) runs per class after a D2H copy.
//
// The reference's loop is sequential in i (pick the current max, decay everything after it),
// but inside one i every later box is decayed independently, and the "overwrite with the last
// row, shrink N, re-examine" removal (:108-115) is a two-pointer compaction whose outcome is
// fully determined by the remove flags: survivors below the new N stay put and the k-th hole
// (ascending) receives the k-th survivor from the top (descending).  One CTA per segment does,
// per i:  block arg-max (first maximum wins, :46-50) -> swap (:53-64) -> parallel decay (:74-106)
// -> ordered ballot/scan compaction.  Each warp owns a contiguous slice of the live range, so
// position order is a running count inside the warp plus a 32-entry scan across warps.
// Rows live in shared memory (SoA) for segments up to kSoftCap boxes, otherwise in place in
// global memory (scratch from the caller's workspace).  fp32 arithmetic in the reference's
// order; the gaussian weight is exp() in double of the fp32 quotient, rounded to fp32 (:97).
// Built with --fmad=false.
#include "rr_common.cuh"

namespace rr {

constexpr int kSoftThreads = 1024;
constexpr int kSoftWarps = kSoftThreads / 32;
constexpr int kSoftCap = 6144;        // boxes per segment held in shared memory

template <bool kSmem>
struct Rows {
    float* f;     // kSmem: SoA [5][kSoftCap]; else AoS rows [n][5] in global memory (segment base)
    int* src;     // original global row id now at each position (may be null in global mode)
    __device__ __forceinline__ float get(int i, int c) const {
        return kSmem ? f[c * kSoftCap + i] : f[(size_t)i * 5 + c];
    }
    __device__ __forceinline__ void set(int i, int c, float v) const {
        if (kSmem) f[c * kSoftCap + i] = v; else f[(size_t)i * 5 + c] = v;
    }
};

template <bool kSmem>
__global__ void __launch_bounds__(kSoftThreads)
soft_nms_kernel(float* __restrict__ boxes, const int* __restrict__ seg_offsets, float sigma, float Nt,
                float threshold, int method, int* __restrict__ src_idx, int* __restrict__ ws_flag,
                int* __restrict__ ws_list, int* __restrict__ keep_count) {
    extern __shared__ float s_dyn[];
    __shared__ float s_red_v[kSoftWarps];
    __shared__ int s_red_i[kSoftWarps];
    __shared__ int s_wcnt[kSoftWarps];     // removed count per warp slice -> exclusive prefix
    __shared__ int s_removed_total;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = seg_offsets[s], n = seg_offsets[s + 1] - s0;
    if (n <= 0) { if (tid == 0 && kSmem) keep_count[s] = 0; return; }
    if (kSmem != (n <= kSoftCap)) return;                 // the other instantiation owns this segment

    Rows<kSmem> R;
    int *flag, *list;
    float* gb = boxes + (size_t)s0 * 5;
    if (kSmem) {
        R.f = s_dyn;
        R.src = reinterpret_cast<int*>(s_dyn + 5 * kSoftCap);
        flag = R.src + kSoftCap;
        list = flag + kSoftCap;
        for (int i = tid; i < n; i += kSoftThreads) {
#pragma unroll
            for (int c = 0; c < 5; ++c) R.f[c * kSoftCap + i] = gb[(size_t)i * 5 + c];
            R.src[i] = s0 + i;
        }
    } else {
        R.f = gb;
        R.src = src_idx ? src_idx + s0 : nullptr;
        flag = ws_flag + s0;
        list = ws_list + s0;
        if (R.src) for (int i = tid; i < n; i += kSoftThreads) R.src[i] = s0 + i;
    }
    __syncthreads();

    int N = n;
    for (int i = 0; i < N; ++i) {
        // ---- arg-max over [i, N): first maximum (cpu_nms.pyx:46-50 compares with strict '<') ----
        float bv = -INFINITY;
        int bp = 0x7fffffff;
        for (int p = i + tid; p < N; p += kSoftThreads) {
            const float v = R.get(p, 4);
            if (v > bv || (v == bv && p < bp)) { bv = v; bp = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
        }
        if (lane == 0) { s_red_v[warp] = bv; s_red_i[warp] = bp; }
        __syncthreads();
        if (warp == 0) {
            bv = s_red_v[lane]; bp = s_red_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
            }
            if (lane == 0) {
                const int mp = (bp == 0x7fffffff) ? i : bp;
                if (mp != i) {                             // swap rows i <-> maxpos (:53-64)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const float t = R.get(i, c);
                        R.set(i, c, R.get(mp, c));
                        R.set(mp, c, t);
                    }
                    if (R.src) { const int t = R.src[i]; R.src[i] = R.src[mp]; R.src[mp] = t; }
                }
            }
        }
        __syncthreads();
        const float tx1 = R.get(i, 0), ty1 = R.get(i, 1), tx2 = R.get(i, 2), ty2 = R.get(i, 3);
        const float tarea = __fmul_rn(__fadd_rn(__fsub_rn(tx2, tx1), 1.f), __fadd_rn(__fsub_rn(ty2, ty1), 1.f));

        // ---- decay every later row (:74-106); warp w owns the contiguous slice [a, b) ----
        const int L = N - (i + 1);
        const int Lw = (((L + kSoftWarps - 1) / kSoftWarps) + 31) & ~31;
        const int a = i + 1 + warp * Lw, b = min(a + Lw, N);
        int wcnt = 0;
        for (int p0 = a; p0 < b; p0 += 32) {
            const int p = p0 + lane;
            bool rem = false;
            if (p < b) {
                const float x1 = R.get(p, 0), y1 = R.get(p, 1), x2 = R.get(p, 2), y2 = R.get(p, 3);
                const float area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
                const float iw = __fadd_rn(__fsub_rn(fminf(tx2, x2), fmaxf(tx1, x1)), 1.f);
                if (iw > 0.f) {
                    const float ih = __fadd_rn(__fsub_rn(fminf(ty2, y2), fmaxf(ty1, y1)), 1.f);
                    if (ih > 0.f) {
                        const float inter = __fmul_rn(iw, ih);
                        const float ua = __fsub_rn(__fadd_rn(tarea, area), inter);
                        const float ov = __fdiv_rn(inter, ua);
                        float weight;
                        if (method == 1) weight = (ov > Nt) ? __fsub_rn(1.f, ov) : 1.f;
                        else if (method == 2) weight = (float)exp((double)(-__fdiv_rn(__fmul_rn(ov, ov), sigma)));
                        else weight = (ov > Nt) ? 0.f : 1.f;
                        const float ns = __fmul_rn(weight, R.get(p, 4));
                        R.set(p, 4, ns);
                        rem = ns < threshold;              // (:108) only tested inside the overlap branch
                    }
                }
                flag[p] = rem ? 1 : 0;
            }
            wcnt += __popc(__ballot_sync(0xffffffffu, rem));
        }
        if (lane == 0) s_wcnt[warp] = wcnt;
        __syncthreads();
        if (warp == 0) {                                   // exclusive scan of the 32 slice counts
            const int v = s_wcnt[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_wcnt[lane] = incl - v;
            if (lane == 31) s_removed_total = incl;
        }
        __syncthreads();
        const int removed = s_removed_total;
        if (removed == 0) continue;                        // uniform

        // ---- ordered compaction: the k-th hole below the new N takes the k-th survivor from the top ----
        const int Nn = N - removed, surv_total = L - removed;
        int r = s_wcnt[warp];                              // removed rows before this slice
        for (int p0 = a; p0 < b; p0 += 32) {
            const int p = p0 + lane;
            const bool in = p < b;
            const bool rem = in && (flag[p] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, rem);
            const int rem_before = r + __popc(m & ((1u << lane) - 1u));
            if (in) {
                if (rem) {
                    if (p < Nn) flag[p] = rem_before + 1;  // hole rank (1-based), consumed below
                } else if (p >= Nn) {
                    const int surv_before = (p - (i + 1)) - rem_before;
                    list[surv_total - 1 - surv_before] = p;
                }
            }
            r += __popc(m);
        }
        __syncthreads();
        for (int p = i + 1 + tid; p < Nn; p += kSoftThreads) {
            const int h = flag[p];
            if (h > 0) {
                const int q = list[h - 1];
#pragma unroll
                for (int c = 0; c < 5; ++c) R.set(p, c, R.get(q, c));
                if (R.src) R.src[p] = R.src[q];
            }
        }
        __syncthreads();
        N = Nn;
    }
    if (kSmem) {
        for (int i = tid; i < N; i += kSoftThreads) {
#pragma unroll
            for (int c = 0; c < 5; ++c) gb[(size_t)i * 5 + c] = R.f[c * kSoftCap + i];
            if (src_idx) src_idx[s0 + i] = R.src[i];
        }
    }
    if (tid == 0) keep_count[s] = N;
}

}  // namespace rr

using namespace rr;

RR_API size_t rr_soft_nms_workspace_bytes(int M) {
    if (M <= 0) return 256;
    return 2 * align_up((size_t)M * sizeof(int));
}

RR_API int rr_soft_nms_batched(float* boxes, const int32_t* seg_offsets, int M, int S,
                               float sigma, float Nt, float threshold, int method,
                               int32_t* src_idx, int32_t* keep_count,
                               void* ws, size_t ws_bytes, void* stream) {
    if (M < 0 || S <= 0 || !seg_offsets || !keep_count) return RR_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (M == 0) {
        RR_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * S, st), rc);
        return rc;
    }
    if (!boxes) return RR_E_BADARG;
    const size_t smem = (size_t)kSoftCap * (5 * sizeof(float) + 3 * sizeof(int));    // 192 KB
    RR_CUDA(cudaFuncSetAttribute(soft_nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
    soft_nms_kernel<true><<<S, kSoftThreads, smem, st>>>(boxes, seg_offsets, sigma, Nt, threshold, method,
                                                        src_idx, nullptr, nullptr, keep_count);
    RR_LAUNCHED_K(rc, "soft_nms_kernel", st);
    if (M > kSoftCap) {                                    // some segment may exceed the smem capacity
        if (!ws) return RR_E_BADARG;
        if (ws_bytes < rr_soft_nms_workspace_bytes(M) || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
        Carver cv(ws);
        int* flag = cv.take<int>((size_t)M);
        int* list = cv.take<int>((size_t)M);
        soft_nms_kernel<false><<<S, kSoftThreads, 0, st>>>(boxes, seg_offsets, sigma, Nt, threshold, method,
                                                          src_idx, flag, list, keep_count);
        RR_LAUNCHED_K(rc, "soft_nms_kernel", st);
    }
    return rc;
}
