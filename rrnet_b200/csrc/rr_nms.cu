// Bit-mask NMS for sm_100a: stage-1 per-class NMS of the decoded top-K boxes, generic
// segmented hard NMS, and the reference's legacy `_nms` host ABI.
//
// Replaces RRNet.nms (models/rrnet.py:56-72, torchvision.ops.nms per class) and ext/nms
// (nms/nms_kernel.cu:34-144, nms/cpu_nms.pyx:122-173, nms/py_cpu_nms.py).  One kernel body
// covers the three semantics in the reference: pixel_offset in {0,1} and '>' vs '>='.
//
// Layout: every group g (an image) owns a list of up to Kmax boxes sorted by (label asc,
// score desc); a segment is the run of one label.  The suppression matrix of a group is
// Kmax x W 64-bit words (W = ceil(Kmax/64)); bit j of word (i, w) says "box w*64+j (> i,
// same label) overlaps box i".  Only upper-triangle 64x64 tiles whose label ranges intersect
// are computed.  The greedy reduce runs on the device, one CTA per segment: the 64x64 diagonal
// tile is resolved serially by one warp on 64-bit words, the off-diagonal words of the kept
// rows are OR-ed in parallel.  Nothing is copied to the host (the reference copies the whole
// mask back and scans it on the CPU, nms_kernel.cu:113-139).
//
// Built with --fmad=false: IoU = inter / (area_a + area_b - inter) must round exactly like
// the CPU implementations (one IEEE divide, no contraction).
#include "rr_common.cuh"

#include <mutex>

namespace rr {

// --------------------------------------------------------------------------------------------
// Stage-1 partition: stable counting sort of one image's K score-sorted rows by class.
// dets [B,K,6] -> sbox/slab/ssrc/sscore [B,K] and seg_off [B,C+1].
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
stage1_partition_kernel(const float* __restrict__ dets, int K, int C,
                        float4* __restrict__ sbox, int* __restrict__ slab, int* __restrict__ ssrc,
                        float* __restrict__ sscore, int* __restrict__ seg_off) {
    extern __shared__ unsigned short s_cnt[];       // [nch][C] per-chunk class counts -> prefixes
    __shared__ int s_base[RR_MAX_CLASSES + 1];
    const int b = blockIdx.x;
    const int nch = (K + 31) >> 5;
    const int lane = lane_id(), warp = warp_id(), nwarp = blockDim.x >> 5;
    const float* d = dets + (size_t)b * K * 6;

    for (int ch = warp; ch < nch; ch += nwarp) {
        int r = ch * 32 + lane;
        int c = -1;
        if (r < K) { c = (int)d[(size_t)r * 6 + 5]; c = min(max(c, 0), C - 1); }
        for (int cc = 0; cc < C; ++cc) {
            unsigned m = __ballot_sync(0xffffffffu, c == cc);
            if (lane == 0) s_cnt[ch * C + cc] = (unsigned short)__popc(m);
        }
    }
    __syncthreads();
    if (threadIdx.x < C) {                           // exclusive scan over chunks, per class
        int cc = threadIdx.x, run = 0;
        for (int ch = 0; ch < nch; ++ch) {
            int v = s_cnt[ch * C + cc];
            s_cnt[ch * C + cc] = (unsigned short)run;
            run += v;
        }
        s_base[cc + 1] = run;                        // class totals, scanned below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_base[0] = 0;
        for (int cc = 0; cc < C; ++cc) s_base[cc + 1] += s_base[cc];
        for (int cc = 0; cc <= C; ++cc) seg_off[b * (C + 1) + cc] = s_base[cc];
    }
    __syncthreads();
    for (int ch = warp; ch < nch; ch += nwarp) {
        int r = ch * 32 + lane;
        int c = -1;
        const float* row = d + (size_t)r * 6;
        if (r < K) { c = (int)row[5]; c = min(max(c, 0), C - 1); }
        int rank = 0;
        for (int cc = 0; cc < C; ++cc) {
            unsigned m = __ballot_sync(0xffffffffu, c == cc);
            if (c == cc) rank = __popc(m & ((1u << lane) - 1u));
        }
        if (r < K) {
            int pos = s_base[c] + s_cnt[ch * C + c] + rank;
            size_t o = (size_t)b * K + pos;
            sbox[o] = make_float4(row[0], row[1], row[2], row[3]);
            slab[o] = c;
            ssrc[o] = r;
            sscore[o] = row[4];
        }
    }
}

// --------------------------------------------------------------------------------------------
// Generic API: stable sort by score (desc) inside each segment by rank counting
// (rank_i = #{j in segment : s_j > s_i or (s_j == s_i and j < i)}), O(n^2) compares, no library.
// --------------------------------------------------------------------------------------------
constexpr int kRankSplit = 8;                     // threads that share one element's rank count

__global__ void __launch_bounds__(256)
rank_sort_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                 const int* __restrict__ seg_offsets, int M, int S,
                 float4* __restrict__ sbox, int* __restrict__ slab, int* __restrict__ ssrc) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt / kRankSplit, part = gt % kRankSplit;     // the kRankSplit threads of an element sit in one warp
    const bool on = i < M;
    int lo = 0, s0 = 0, s1 = 0;
    float si = 0.f;
    if (on) {
        int hi = S;                                  // largest s with seg_offsets[s] <= i
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (seg_offsets[mid] <= i) lo = mid; else hi = mid;
        }
        s0 = seg_offsets[lo]; s1 = seg_offsets[lo + 1];
        si = scores[i];
    }
    int rank = 0;
    for (int j = s0 + part; j < s1; j += kRankSplit) {
        float sj = __ldg(scores + j);
        rank += (sj > si) || (sj == si && j < i);
    }
#pragma unroll
    for (int o = kRankSplit / 2; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (!on || part != 0) return;
    int pos = s0 + rank;
    const float* bx = boxes + (size_t)i * 4;
    sbox[pos] = make_float4(bx[0], bx[1], bx[2], bx[3]);
    slab[pos] = lo;
    ssrc[pos] = i;
}

// Legacy ABI: rows [n,dim] already sorted; repack to float4, single label.
__global__ void repack_rows_kernel(const float* __restrict__ rows, int n, int dim,
                                   float4* __restrict__ sbox, int* __restrict__ slab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rows + (size_t)i * dim;
    sbox[i] = make_float4(r[0], r[1], r[2], r[3]);
    slab[i] = 0;
}

// --------------------------------------------------------------------------------------------
// Suppression mask, one 64x64 upper-triangle tile per CTA (64 threads, one row each).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ sbox, const int* __restrict__ slab,
                int n, int Kmax, int W, int T, double thr, float o, int ge,
                unsigned long long* __restrict__ mask) {
    const int g = blockIdx.y;
    // blockIdx.x -> (rt, ct), ct >= rt, row-major over the upper triangle of a T x T grid
    const long long p = blockIdx.x;
    int rt = (int)(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)p)) * 0.5);
    rt = max(0, min(rt, T - 1));
    while (rt > 0 && (long long)rt * T - (long long)rt * (rt - 1) / 2 > p) --rt;
    while ((long long)(rt + 1) * T - (long long)(rt + 1) * rt / 2 <= p) ++rt;
    const int ct = rt + (int)(p - ((long long)rt * T - (long long)rt * (rt - 1) / 2));
    if (rt * 64 >= n || ct * 64 >= n) return;

    const float4* bx = sbox + (size_t)g * Kmax;
    const int* lb = slab + (size_t)g * Kmax;
    // labels ascend along the list: the tiles share a label iff last(row tile) >= first(col tile)
    if (lb[min(rt * 64 + 63, n - 1)] < lb[ct * 64]) return;

    __shared__ float4 c_box[64];
    __shared__ float c_area[64];
    __shared__ int c_lab[64];
    const int tid = threadIdx.x;
    const int col_n = min(64, n - ct * 64);
    if (tid < col_n) {
        float4 q = bx[ct * 64 + tid];
        c_box[tid] = q;
        c_area[tid] = __fmul_rn(__fadd_rn(__fsub_rn(q.z, q.x), o), __fadd_rn(__fsub_rn(q.w, q.y), o));
        c_lab[tid] = lb[ct * 64 + tid];
    }
    __syncthreads();
    const int i = rt * 64 + tid;
    if (i >= n) return;
    const float4 a = bx[i];
    const int la = lb[i];
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), o), __fadd_rn(__fsub_rn(a.w, a.y), o));
    // IoU of a pair with empty intersection is 0/u: suppresses only if 0 cmp thr holds
    const bool zero_hits = ge ? (0.0 >= thr) : (0.0 > thr);
    unsigned long long bits = 0;
    const int start = (rt == ct) ? tid + 1 : 0;
    for (int j = start; j < col_n; ++j) {
        if (c_lab[j] != la) continue;
        const float4 q = c_box[j];
        float w = __fadd_rn(__fsub_rn(fminf(a.z, q.z), fmaxf(a.x, q.x)), o);
        float h = __fadd_rn(__fsub_rn(fminf(a.w, q.w), fmaxf(a.y, q.y)), o);
        w = fmaxf(w, 0.0f);
        h = fmaxf(h, 0.0f);
        const float inter = __fmul_rn(w, h);
        if (inter > 0.0f || zero_hits) {
            const float uni = __fsub_rn(__fadd_rn(area_a, c_area[j]), inter);
            const double iou = (double)__fdiv_rn(inter, uni);
            if (ge ? (iou >= thr) : (iou > thr)) bits |= 1ull << j;
        }
    }
    mask[((size_t)g * Kmax + i) * W + ct] = bits;
}

// --------------------------------------------------------------------------------------------
// Greedy reduce of one segment (one CTA).  keep_out[g*Kmax + s0 + t] = t-th accepted box of
// the segment: its list position, or map[position] when `map` is given.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ seg_off,
                int segs_per_group, int Kmax, int W, const int* __restrict__ map,
                int* __restrict__ keep_out, int* __restrict__ keep_cnt) {
    extern __shared__ unsigned long long s_remv[];       // [W]
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_kept;
    const int g = blockIdx.y, s = blockIdx.x, tid = threadIdx.x;
    const int s0 = seg_off[g * (segs_per_group + 1) + s];
    const int s1 = seg_off[g * (segs_per_group + 1) + s + 1];
    if (s1 <= s0) {
        if (tid == 0) keep_cnt[g * segs_per_group + s] = 0;
        return;
    }
    const unsigned long long* m = mask + (size_t)g * Kmax * W;
    const int t0 = s0 >> 6, t1 = (s1 - 1) >> 6;
    for (int w = t0 + tid; w <= t1; w += blockDim.x) s_remv[w] = 0ull;
    int nk = 0;                                           // meaningful in warp 0
    for (int t = t0; t <= t1; ++t) {
        const int row0 = t << 6;
        if (tid < 64) {
            int r = row0 + tid;
            s_diag[tid] = (r >= s0 && r < s1) ? m[(size_t)r * W + t] : 0ull;
        }
        __syncthreads();
        if (tid < 32) {
            const int lo = max(s0, row0) - row0, hi = min(s1, row0 + 64) - row0;   // live bit range
            unsigned long long valid = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
            unsigned long long cur = s_remv[t] | ~valid;
            unsigned long long kept = 0ull;
#pragma unroll 8
            for (int bit = 0; bit < 64; ++bit) {
                if (!((cur >> bit) & 1ull)) {
                    kept |= 1ull << bit;
                    cur |= s_diag[bit];
                }
            }
            if (tid == 0) s_kept = kept;
            int* out = keep_out + (size_t)g * Kmax + s0 + nk;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                int bit = tid + 32 * half;
                if ((kept >> bit) & 1ull) {
                    int idx = __popcll(kept & ((1ull << bit) - 1ull));
                    int pos = row0 + bit;
                    out[idx] = map ? map[(size_t)g * Kmax + pos] : pos;
                }
            }
            nk += __popcll(kept);
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        // OR the kept rows' mask words into remv for every later tile.  The 64 rows are split in 4 groups of
        // 16 bits handled by different threads (up to 16 independent loads in flight each) and merged with a
        // shared-memory atomicOr, so a tile costs one or two L2 round trips instead of eight.
        const int nw = t1 - t;                               // words t+1 .. t1
        for (int it = tid; it < nw * 4; it += blockDim.x) {
            const int w = t + 1 + (it >> 2), grp = it & 3;
            unsigned long long kk = (kept >> (16 * grp)) & 0xffffull;
            if (!kk) continue;
            unsigned long long v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                v[u] = 0ull;
                if (kk) {
                    int bit = __ffsll((long long)kk) - 1;
                    kk &= kk - 1ull;
                    v[u] = m[(size_t)(row0 + 16 * grp + bit) * W + w];
                }
            }
            unsigned long long acc = 0ull;
#pragma unroll
            for (int u = 0; u < 16; ++u) acc |= v[u];
            if (acc) atomicOr(&s_remv[w], acc);
        }
        // the next iteration's first barrier orders these s_remv writes before they are read
    }
    if (tid == 0) keep_cnt[g * segs_per_group + s] = nk;
}

// --------------------------------------------------------------------------------------------
// Stage-1 compaction: image-major, class-ascending, score-descending rows
// (models/rrnet.py:44-49,60-72) + counts [B+1].
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
stage1_compact_kernel(const float4* __restrict__ sbox, const float* __restrict__ sscore,
                      const int* __restrict__ slab, const int* __restrict__ seg_off,
                      const int* __restrict__ keep_pos, const int* __restrict__ keep_cnt,
                      int B, int K, int C,
                      float* __restrict__ out_bxyxy, float* __restrict__ out_scores,
                      float* __restrict__ out_clses, int* __restrict__ out_counts) {
    __shared__ int s_cls_base[RR_MAX_CLASSES + 1];
    __shared__ int s_img_base;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {                                 // warp 0: rows of the images before this one, class offsets inside it
        int before = 0;
        for (int i = tid; i < b * C; i += 32) before += keep_cnt[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        if (tid == 0) {
            s_img_base = before;
            int run = 0;
            for (int c = 0; c < C; ++c) { s_cls_base[c] = run; run += keep_cnt[b * C + c]; }
            s_cls_base[C] = run;
            out_counts[b] = run;
            if (b == B - 1) out_counts[B] = before + run;
        }
    }
    __syncthreads();
    const int base = s_img_base, total = s_cls_base[C];
    for (int t = tid; t < total; t += blockDim.x) {       // all classes in one sweep: row t of the image's output
        int c = 0;
        while (t >= s_cls_base[c + 1]) ++c;
        const int pos = keep_pos[(size_t)b * K + seg_off[b * (C + 1) + c] + (t - s_cls_base[c])];
        const size_t src = (size_t)b * K + pos;
        const float4 q = sbox[src];
        const size_t o = (size_t)base + t;
        float* r = out_bxyxy + o * 5;
        r[0] = (float)b; r[1] = q.x; r[2] = q.y; r[3] = q.z; r[4] = q.w;
        out_scores[o] = sscore[src];
        out_clses[o] = (float)slab[src];
    }
}

// --------------------------------------------------------------------------------------------
// host-side launch helpers
// --------------------------------------------------------------------------------------------
static int launch_mask_scan(const float4* sbox, const int* slab, const int* seg_off, int G,
                            int segs_per_group, int n, int Kmax, double thr, int pixel_offset,
                            int ge_cmp, unsigned long long* mask, const int* map, int* keep_out,
                            int* keep_cnt, cudaStream_t st) {
    int rc = 0;
    const int W = (Kmax + 63) / 64, T = W;
    const long long pairs = (long long)T * (T + 1) / 2;
    if (pairs > 0x7fffffffLL) return RR_E_RANGE;
    dim3 gm((unsigned)pairs, (unsigned)G);
    nms_mask_kernel<<<gm, 64, 0, st>>>(sbox, slab, n, Kmax, W, T, thr, pixel_offset ? 1.0f : 0.0f,
                                       ge_cmp ? 1 : 0, mask);
    RR_LAUNCHED_K(rc, "nms_mask_kernel", st);
    dim3 gs((unsigned)segs_per_group, (unsigned)G);
    size_t smem = (size_t)W * sizeof(unsigned long long);
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return RR_E_RANGE;
        RR_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
    }
    nms_scan_kernel<<<gs, 1024, smem, st>>>(mask, seg_off, segs_per_group, Kmax, W, map, keep_out, keep_cnt);
    RR_LAUNCHED_K(rc, "nms_scan_kernel", st);
    return rc;
}

struct Stage1Ws {
    float4* sbox; int* slab; int* ssrc; float* sscore; int* seg_off; int* keep_pos; int* keep_cnt;
    unsigned long long* mask;
    size_t bytes;
};
static Stage1Ws carve_stage1(void* ws, int B, int K, int C) {
    Carver cv(ws);
    Stage1Ws w;
    w.sbox = cv.take<float4>((size_t)B * K);
    w.slab = cv.take<int>((size_t)B * K);
    w.ssrc = cv.take<int>((size_t)B * K);
    w.sscore = cv.take<float>((size_t)B * K);
    w.seg_off = cv.take<int>((size_t)B * (C + 1));
    w.keep_pos = cv.take<int>((size_t)B * K);
    w.keep_cnt = cv.take<int>((size_t)B * C);
    w.mask = cv.take<unsigned long long>((size_t)B * K * ((K + 63) / 64));
    w.bytes = cv.off;
    return w;
}

int stage1_nms_launch(const float* dets, int B, int K, int C, double thr, float* out_bxyxy,
                      float* out_scores, float* out_clses, int32_t* out_counts, void* ws,
                      cudaStream_t st) {
    int rc = 0;
    Stage1Ws w = carve_stage1(ws, B, K, C);
    const int nch = (K + 31) / 32;
    size_t smem = (size_t)nch * C * sizeof(unsigned short);
    if (smem > 48 * 1024)
        RR_CUDA(cudaFuncSetAttribute(stage1_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
    stage1_partition_kernel<<<B, 1024, smem, st>>>(dets, K, C, w.sbox, w.slab, w.ssrc, w.sscore, w.seg_off);
    RR_LAUNCHED_K(rc, "stage1_partition_kernel", st);
    int r2 = launch_mask_scan(w.sbox, w.slab, w.seg_off, B, C, K, K, thr, 0, 0, w.mask, nullptr,
                              w.keep_pos, w.keep_cnt, st);
    if (rc == 0) rc = r2;
    stage1_compact_kernel<<<B, 1024, 0, st>>>(w.sbox, w.sscore, w.slab, w.seg_off, w.keep_pos, w.keep_cnt,
                                             B, K, C, out_bxyxy, out_scores, out_clses, out_counts);
    RR_LAUNCHED_K(rc, "stage1_compact_kernel", st);
    return rc;
}

size_t stage1_nms_ws_bytes(int B, int K, int C) { return carve_stage1(nullptr, B, K, C).bytes; }

struct GenericWs {
    float4* sbox; int* slab; int* ssrc; unsigned long long* mask; size_t bytes;
};
static GenericWs carve_generic(void* ws, int M) {
    Carver cv(ws);
    GenericWs w;
    w.sbox = cv.take<float4>((size_t)M);
    w.slab = cv.take<int>((size_t)M);
    w.ssrc = cv.take<int>((size_t)M);
    w.mask = cv.take<unsigned long long>((size_t)M * ((M + 63) / 64));
    w.bytes = cv.off;
    return w;
}

}  // namespace rr

using namespace rr;

RR_API size_t rr_stage1_nms_workspace_bytes(int B, int K, int num_classes) {
    if (B <= 0 || K <= 0 || num_classes <= 0) return 0;
    return stage1_nms_ws_bytes(B, K, num_classes);
}

RR_API int rr_stage1_nms(const float* dets, int B, int K, int num_classes, double thr,
                         float* out_bxyxy, float* out_scores, float* out_clses, int32_t* out_counts,
                         void* ws, size_t ws_bytes, void* stream) {
    if (!dets || !out_bxyxy || !out_scores || !out_clses || !out_counts || !ws) return RR_E_BADARG;
    if (B <= 0 || K <= 0 || num_classes <= 0) return RR_E_BADARG;
    if (num_classes > RR_MAX_CLASSES || K > 65535) return RR_E_RANGE;
    if (ws_bytes < stage1_nms_ws_bytes(B, K, num_classes) || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    return stage1_nms_launch(dets, B, K, num_classes, thr, out_bxyxy, out_scores, out_clses, out_counts,
                             ws, (cudaStream_t)stream);
}

RR_API size_t rr_nms_workspace_bytes(int M, int S) {
    (void)S;
    if (M <= 0) return 256;
    return carve_generic(nullptr, M).bytes;
}

RR_API int rr_nms_batched(const float* boxes, const float* scores, const int32_t* seg_offsets,
                          int M, int S, double thr, int pixel_offset, int ge_cmp,
                          int32_t* keep_idx, int32_t* keep_count,
                          void* ws, size_t ws_bytes, void* stream) {
    if (M < 0 || S <= 0 || !seg_offsets || !keep_count) return RR_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (M == 0) {                                   // empty input: all counts zero
        RR_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * S, st), rc);
        return rc;
    }
    if (!boxes || !scores || !keep_idx || !ws) return RR_E_BADARG;
    if (ws_bytes < carve_generic(nullptr, M).bytes || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    GenericWs w = carve_generic(ws, M);
    rank_sort_kernel<<<(int)(((long long)M * kRankSplit + 255) / 256), 256, 0, st>>>(boxes, scores, seg_offsets, M, S, w.sbox, w.slab, w.ssrc);
    RR_LAUNCHED_K(rc, "rank_sort_kernel", st);
    int r2 = launch_mask_scan(w.sbox, w.slab, seg_offsets, 1, S, M, M, thr, pixel_offset, ge_cmp, w.mask,
                              w.ssrc, keep_idx, keep_count, st);
    return rc ? rc : r2;
}

// ---- legacy `_nms` ABI (ext/nms/nms/gpu_nms.hpp:1-2) -----------------------------------------
namespace {
struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
};
std::mutex g_mu;
Scratch g_scratch[64];

int scratch_get(int dev, size_t bytes, void** out) {
    Scratch& s = g_scratch[dev & 63];
    if (s.bytes < bytes) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.bytes = 0;
        cudaError_t e = cudaMalloc(&s.p, bytes);
        if (e != cudaSuccess) return (int)e;
        s.bytes = bytes;
    }
    *out = s.p;
    return 0;
}
}  // namespace

RR_API int rr_nms_legacy_host(int* keep_out_host, int* num_out_host, const float* boxes_host,
                              int boxes_num, int boxes_dim, float nms_overlap_thresh, int device_id) {
    if (!num_out_host) return RR_E_BADARG;
    if (boxes_num <= 0) { *num_out_host = 0; return boxes_num == 0 ? 0 : RR_E_BADARG; }
    if (!keep_out_host || !boxes_host || boxes_dim < 4) return RR_E_BADARG;
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = 0;
    RR_CUDA(cudaSetDevice(device_id), rc);          // the reference does the same (nms_kernel.cu:80-89)
    if (rc) return rc;
    const int n = boxes_num;
    GenericWs lay = carve_generic(nullptr, n);
    const size_t rows_b = align_up((size_t)n * boxes_dim * sizeof(float));
    const size_t tail_b = align_up((size_t)n * sizeof(int)) + 512;
    void* base = nullptr;
    rc = scratch_get(device_id, lay.bytes + rows_b + tail_b, &base);
    if (rc) return rc;
    GenericWs w = carve_generic(base, n);
    char* q = (char*)base + lay.bytes;
    float* rows = (float*)q;
    int* keep = (int*)(q + rows_b);
    int* seg_off = (int*)(q + rows_b + align_up((size_t)n * sizeof(int)));
    int* cnt = seg_off + 2;
    cudaStream_t st = 0;
    int segs[2] = {0, n};
    RR_CUDA(cudaMemcpyAsync(rows, boxes_host, (size_t)n * boxes_dim * sizeof(float), cudaMemcpyHostToDevice, st), rc);
    RR_CUDA(cudaMemcpyAsync(seg_off, segs, sizeof(segs), cudaMemcpyHostToDevice, st), rc);
    repack_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(rows, n, boxes_dim, w.sbox, w.slab);
    RR_LAUNCHED_K(rc, "repack_rows_kernel", st);
    int r2 = launch_mask_scan(w.sbox, w.slab, seg_off, 1, 1, n, n, (double)nms_overlap_thresh, 1, 0, w.mask,
                              nullptr, keep, cnt, st);
    if (rc == 0) rc = r2;
    int num = 0;
    RR_CUDA(cudaMemcpyAsync(&num, cnt, sizeof(int), cudaMemcpyDeviceToHost, st), rc);
    RR_CUDA(cudaStreamSynchronize(st), rc);
    if (rc) return rc;
    RR_CUDA(cudaMemcpy(keep_out_host, keep, (size_t)num * sizeof(int), cudaMemcpyDeviceToHost), rc);
    *num_out_host = num;
    return rc;
}
