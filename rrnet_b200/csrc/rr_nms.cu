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
    RR_PDL_PROLOGUE();
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
    RR_PDL_PROLOGUE();
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
// Suppression mask, one 64x64 upper-triangle tile per CTA of 4 warps, by warp ballot: a lane owns two COLUMN
// boxes of the tile (columns lane and lane + 32: box, area and label stay in registers for the whole tile), a warp
// walks 16 ROW boxes (one broadcast shared-memory read per row, not per pair), every lane decides its two pairs
// and two __ballot_sync calls assemble the row's 64-bit word, which lane (row & 15) keeps and stores.  No
// shared-memory traffic per pair and no serial per-thread column loop (the reference's mapping, nms_kernel.cu:34-78,
// is one thread per row looping over 64 columns).  Decision arithmetic: one IEEE divide, no contraction (--fmad=false).
// --------------------------------------------------------------------------------------------
constexpr int kMaskThreads = 128;
// words per mask row: one per 64-box tile, padded to a multiple of 4 so that a row's words [4b, 4b+3] are one aligned
// 32-byte sector (the pullers of the greedy scan fetch four columns of a row with one sector)
static inline int mask_row_words(int kmax) { return ((kmax + 63) / 64 + 3) & ~3; }

__device__ __forceinline__ bool nms_pair(const float4& a, float area_a, const float4& q, float area_q, float o,
                                         double thr, int ge, bool zero_hits) {
    float w = __fadd_rn(__fsub_rn(fminf(a.z, q.z), fmaxf(a.x, q.x)), o);
    float h = __fadd_rn(__fsub_rn(fminf(a.w, q.w), fmaxf(a.y, q.y)), o);
    w = fmaxf(w, 0.0f);
    h = fmaxf(h, 0.0f);
    const float inter = __fmul_rn(w, h);
    if (!(inter > 0.0f || zero_hits)) return false;
    const float uni = __fsub_rn(__fadd_rn(area_a, area_q), inter);
    const double iou = (double)__fdiv_rn(inter, uni);
    return ge ? (iou >= thr) : (iou > thr);
}

__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(const float4* __restrict__ sbox, const int* __restrict__ slab,
                int n, int Kmax, int W, int T, double thr, float o, int ge,
                unsigned long long* __restrict__ mask) {
    RR_PDL_PROLOGUE();
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

    __shared__ float4 r_box[64];
    __shared__ float2 r_al[64];                       // (area, label bits) of the row boxes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) {
        const int i = rt * 64 + tid;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int la = -1;                                   // rows past the list never match a label
        if (i < n) { a = bx[i]; la = lb[i]; }
        r_box[tid] = a;
        r_al[tid] = make_float2(__fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), o), __fadd_rn(__fsub_rn(a.w, a.y), o)),
                                __int_as_float(la));
    }
    // this lane's two column boxes
    float4 q[2];
    float area_q[2];
    int lq[2], jq[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        jq[h] = ct * 64 + 32 * h + lane;
        q[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        lq[h] = -2;                                    // columns past the list never match either
        if (jq[h] < n) { q[h] = bx[jq[h]]; lq[h] = lb[jq[h]]; }
        area_q[h] = __fmul_rn(__fadd_rn(__fsub_rn(q[h].z, q[h].x), o), __fadd_rn(__fsub_rn(q[h].w, q[h].y), o));
    }
    __syncthreads();
    // IoU of a pair with empty intersection is 0/u: suppresses only if 0 cmp thr holds
    const bool zero_hits = ge ? (0.0 >= thr) : (0.0 > thr);
    unsigned long long mine = 0ull;
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
        const int rl = warp * 16 + r, i = rt * 64 + rl;
        const float4 a = r_box[rl];                    // warp-uniform address: broadcast
        const float2 al = r_al[rl];
        const int la = __float_as_int(al.y);
        const bool hit0 = lq[0] == la && jq[0] > i && nms_pair(a, al.x, q[0], area_q[0], o, thr, ge, zero_hits);
        const bool hit1 = lq[1] == la && jq[1] > i && nms_pair(a, al.x, q[1], area_q[1], o, thr, ge, zero_hits);
        const unsigned m0 = __ballot_sync(0xffffffffu, hit0), m1 = __ballot_sync(0xffffffffu, hit1);
        if (lane == r) mine = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    }
    const int i = rt * 64 + warp * 16 + lane;
    if (lane < 16 && i < n) mask[((size_t)g * Kmax + i) * W + ct] = mine;
}

// --------------------------------------------------------------------------------------------
// Greedy reduce of one segment (one CTA).  keep_out[g*Kmax + s0 + t] = t-th accepted box of
// the segment: its list position, or map[position] when `map` is given.
//
// The chain over the segment's 64-box tiles is serial by nature; what it must NOT contain is a memory round trip
// per tile.  Three roles, hand-over through flags in shared memory:
//   * warp 0, the RESOLVER: for tile t it takes the tile's diagonal words and the words of the next kNear columns
//     from a shared-memory ring, walks the not-yet-suppressed bits in order (one shared-memory read per KEPT box)
//     and folds the kept rows' near words into the removal words of the next kNear tiles with warp OR-reductions.
//   * warps 1..7, the LOADERS: run up to kRing tiles ahead and copy those (1 + kNear) x 64 words per tile from the
//     mask into the ring, so the L2 latency never sits on the chain.
//   * warps 8..31, the PULLERS: a warp owns a column word w (round robin) and ORs, over all tiles t <= w - kNear - 1
//     as they get resolved, the words [row][w] of the kept rows - many independent loads in flight, one writer per
//     removal word, no atomics; the resolver needs the word kNear + 1 steps after the last tile it depends on.
// --------------------------------------------------------------------------------------------
#ifdef RR_SCAN_TRACE
__device__ long long g_scan_trace[8];
#define SCAN_LAP(i) do { const long long n_ = clock64(); tr_[i] += n_ - lap_; lap_ = n_; } while (0)
#else
#define SCAN_LAP(i) do { } while (0)
#endif
constexpr int kScanThreads2 = 640;                      // 102 registers per thread: the resolver keeps a tile's diagonal words in registers
constexpr int kNear = 3;                 // columns t+1 .. t+kNear come from the ring
constexpr int kRing = 16;                // tiles the loaders may run ahead
constexpr int kRingWords = 64 * (1 + kNear);
// Warp w issues from scheduler w % 4.  The resolver's chain is latency bound, so it gets scheduler 0 to itself: warps
// 4, 8, .. exit at once and the polling roles live on the other three schedulers.
constexpr int kWorkerWarps = (kScanThreads2 / 32) * 3 / 4;   // 15 warps with w % 4 != 0
constexpr int kLoaderWarps = 4;
constexpr int kPullerWarps = kWorkerWarps - kLoaderWarps;

// Hand-over between the roles: data first, then the flag, both as volatile shared-memory stores of ONE thread (or plain
// stores of a warp, __syncwarp, then the flag).  Shared-memory accesses of a warp are performed in issue order and there
// is no cache in between, so a reader that sees the flag sees the data; the compiler keeps volatile accesses in order.
// A memory fence here (__threadfence_block = MEMBAR.SC.CTA, or fence.acq_rel.cta) also waits for the thread's GLOBAL
// stores in flight (the keep list): measured 5-6 k cycles per tile, i.e. the whole chain.  -DRR_SCAN_FENCE restores it.
__device__ __forceinline__ void scan_release() {
#ifdef RR_SCAN_FENCE
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#else
    asm volatile("" ::: "memory");
#endif
}

// back-off of a polling worker lane: an ALU-only delay.  NANOSLEEP stalls the resolver (see the kernel comment).
__device__ __forceinline__ void scan_backoff() {
#ifdef RR_SCAN_NANOSLEEP
    __nanosleep(100);
#else
    const long long t = clock64();
    while (clock64() - t < 150) { }
#endif
}

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(kScanThreads2, 1)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ seg_off,
                int segs_per_group, int Kmax, int W,
                int* __restrict__ keep_out, int* __restrict__ keep_cnt) {
    RR_PDL_PROLOGUE();
    extern __shared__ unsigned long long s_dyn[];        // remv [W] | kept [W] | colflag [W] (int)
    __shared__ unsigned long long s_ring[kRing][kRingWords];      // [slot][row][0 = diagonal, 1.. = near columns]
    __shared__ int s_ready[kRing];                       // tile index + 1 whose words sit in the slot
    __shared__ int s_done;                               // tiles resolved
    unsigned long long* s_remv = s_dyn;
    unsigned long long* s_kept = s_dyn + W;
    int* s_colflag = reinterpret_cast<int*>(s_dyn + 2 * W);
    const int g = blockIdx.y, s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = seg_off[g * (segs_per_group + 1) + s];
    const int s1 = seg_off[g * (segs_per_group + 1) + s + 1];
    if (s1 <= s0) {
        if (tid == 0) keep_cnt[g * segs_per_group + s] = 0;
        return;
    }
    const unsigned long long* m = mask + (size_t)g * Kmax * W;
    const int t0 = s0 >> 6, t1 = (s1 - 1) >> 6, ntile = t1 - t0 + 1;
    for (int w = t0 + tid; w <= t1; w += blockDim.x) { s_remv[w] = 0ull; s_colflag[w] = 0; }
    if (tid < kRing) s_ready[tid] = 0;
    if (tid == 0) s_done = 0;
    __syncthreads();
    volatile int* v_done = &s_done;
    volatile int* v_ready = s_ready;
    volatile int* v_colflag = s_colflag;

    if (warp == 0) {
        // ------------------------------ resolver ------------------------------
        // Everything below is executed by all 32 lanes in lock step (uniform polls, no `if (lane == 0)` branches): a warp
        // that reaches a collective in a diverged state takes the slow BRA.DIV path (measured: ~2 k cycles per collective).
        unsigned long long carry[kNear];                 // near contributions to tiles t, t+1, .. (from earlier tiles)
#pragma unroll
        for (int j = 0; j < kNear; ++j) carry[j] = 0ull;
        int nk = 0;
#ifdef RR_SCAN_TRACE
        long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lap_ = clock64();
#endif
        for (int k = 0; k < ntile; ++k) {
            const int t = t0 + k, row0 = t << 6, slot = k % kRing;
            SCAN_LAP(5);
            while (v_ready[slot] != k + 1) { }
            SCAN_LAP(0);
            if (k > kNear) { while (v_colflag[t] == 0) { } }
            asm volatile("" ::: "memory");
            SCAN_LAP(1);
            const int lo = max(s0, row0) - row0, hi = min(s1, row0 + 64) - row0;   // live bit range
            const unsigned long long valid = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&s_remv[t]) | carry[0] | ~valid;
            // the 64 diagonal words -> registers (broadcast reads, all issued before the chain), then 64 steps with
            // compile-time bit positions on registers only: test / predicated OR, no memory access on the chain
            unsigned d_lo[64], d_hi[64];
            const unsigned long long* ringp = s_ring[slot];
#pragma unroll
            for (int bit = 0; bit < 64; ++bit) {
                const unsigned long long dd = ringp[bit * (1 + kNear)];
                d_lo[bit] = (unsigned)dd; d_hi[bit] = (unsigned)(dd >> 32);
                asm volatile("" : "+r"(d_lo[bit]), "+r"(d_hi[bit]));          // materialise now: not sunk into the chain below
            }
            unsigned c_lo = (unsigned)cur, c_hi = (unsigned)(cur >> 32), k_lo = 0u, k_hi = 0u;
#pragma unroll
            for (int bit = 0; bit < 32; ++bit) {                              // two dependent operations per step
                if (!(c_lo & (1u << bit))) { k_lo |= 1u << bit; c_lo |= d_lo[bit]; c_hi |= d_hi[bit]; }
            }
#pragma unroll
            for (int bit = 0; bit < 32; ++bit) {                              // rows 32..63 only suppress boxes 33..63
                if (!(c_hi & (1u << bit))) { k_hi |= 1u << bit; c_hi |= d_hi[32 + bit]; }
            }
            const unsigned long long kept = ((unsigned long long)k_hi << 32) | k_lo;
            SCAN_LAP(2);
            // near columns: the kept rows' words of columns t+1 .. t+kNear (lane l owns rows l and l + 32), OR-reduced over
            // the warp with a shuffle butterfly (the three reductions interleave)
            const bool k0 = (kept >> lane) & 1ull, k1 = (kept >> (lane + 32)) & 1ull;
            unsigned long long nearw[kNear];
#pragma unroll
            for (int j = 0; j < kNear; ++j) {
                const unsigned long long a = ringp[lane * (1 + kNear) + 1 + j], b2 = ringp[(lane + 32) * (1 + kNear) + 1 + j];
                nearw[j] = (k0 ? a : 0ull) | (k1 ? b2 : 0ull);
            }
            SCAN_LAP(6);
            asm volatile("" ::: "memory");
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&s_kept[t]) = kept;
            if (lane == 0) *v_done = k + 1;                           // pullers may use the tile, loaders may reuse the slot
            SCAN_LAP(7);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < kNear; ++j) nearw[j] |= __shfl_xor_sync(0xffffffffu, nearw[j], o);
#pragma unroll
            for (int j = 0; j < kNear; ++j) carry[j] = (j + 1 < kNear ? carry[j + 1] : 0ull) | nearw[j];
            SCAN_LAP(3);
            int* out = keep_out + (size_t)g * Kmax + s0 + nk;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int bit = lane + 32 * half;
                if ((kept >> bit) & 1ull) {
                    const int idx = __popcll(kept & ((1ull << bit) - 1ull));
                    out[idx] = row0 + bit;       // list position; nms_map_kernel translates afterwards (a global LOAD here sat on the chain)
                }
            }
            nk += __popcll(kept);
            SCAN_LAP(4);
        }
#ifdef RR_SCAN_TRACE
        if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0) { for (int i = 0; i < 8; ++i) g_scan_trace[i] = tr_[i]; g_scan_trace[5] = ntile; }
#endif
        if (lane == 0) keep_cnt[g * segs_per_group + s] = nk;
        return;
    }

    if ((warp & 3) == 0) return;                          // scheduler 0 belongs to the resolver
    const int wk = warp - 1 - (warp >> 2);                // 0 .. kWorkerWarps-1 over the warps with w % 4 != 0
    if (wk < kLoaderWarps) {
        // ------------------------------ loaders ------------------------------
        for (int k = wk; k < ntile; k += kLoaderWarps) {
            const int t = t0 + k, slot = k % kRing;
            if (k >= kRing) {                                         // the slot's previous tile must be resolved
                while (*v_done < k - kRing + 1) scan_backoff();
            }
            unsigned long long v[2][1 + kNear];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = (t << 6) + 32 * h + lane;
                const bool on = r >= s0 && r < s1;
                const unsigned long long* row = m + (size_t)r * W + t;
#pragma unroll
                for (int j = 0; j <= kNear; ++j) v[h][j] = (on && t + j <= t1) ? __ldcg(row + j) : 0ull;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j <= kNear; ++j)
                    *reinterpret_cast<volatile unsigned long long*>(&s_ring[slot][(32 * h + lane) * (1 + kNear) + j]) = v[h][j];
            __syncwarp();
            if (lane == 0) {
                scan_release();
                v_ready[slot] = k + 1;
            }
        }
        return;
    }

    // ------------------------------ pullers ------------------------------
    // A warp owns a block of four consecutive column words [4b, 4b+3]: the four words of a mask row are one aligned
    // 32-byte sector.  Column w takes its far contributions from the tiles t <= w - kNear - 1, so the block first
    // accumulates the tiles that all four columns need, then one more tile per column, publishing each column as soon
    // as its own range is complete.
    auto pull_tiles = [&](int ta, int tb, int w4, unsigned long long (&acc)[4]) {      // tiles ta..tb (<= 4 of them)
        ulonglong2 v[4][2][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = ta + u;
            unsigned long long kept = 0ull;
            if (t <= tb) kept = *reinterpret_cast<volatile unsigned long long*>(&s_kept[t]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool on = (kept >> (lane + 32 * h)) & 1ull;
                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(m + ((size_t)(t << 6) + lane + 32 * h) * W + w4);
                v[u][h][0] = on ? __ldcg(src) : make_ulonglong2(0ull, 0ull);
                v[u][h][1] = on ? __ldcg(src + 1) : make_ulonglong2(0ull, 0ull);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                acc[0] |= v[u][h][0].x; acc[1] |= v[u][h][0].y; acc[2] |= v[u][h][1].x; acc[3] |= v[u][h][1].y;
            }
    };
    auto publish = [&](int w, unsigned long long a) {
        a |= __shfl_xor_sync(0xffffffffu, a, 16); a |= __shfl_xor_sync(0xffffffffu, a, 8); a |= __shfl_xor_sync(0xffffffffu, a, 4);
        a |= __shfl_xor_sync(0xffffffffu, a, 2); a |= __shfl_xor_sync(0xffffffffu, a, 1);
        if (lane == 0 && w <= t1 && w >= t0 + kNear + 1) {
            *reinterpret_cast<volatile unsigned long long*>(&s_remv[w]) = a;
            scan_release();
            v_colflag[w] = 1;
        }
    };
    for (int blk = ((t0 + kNear + 1) >> 2) + (wk - kLoaderWarps); 4 * blk <= t1; blk += kPullerWarps) {
        const int w4 = 4 * blk;
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        const int common = w4 - kNear - 1;                            // last tile that all four columns need
        for (int ta = t0; ta <= common; ta += 4) {
            const int tb = min(ta + 3, common);
            while (*v_done < tb - t0 + 1) scan_backoff();
            pull_tiles(ta, tb, w4, acc);
        }
        publish(w4, acc[0]);
#pragma unroll
        for (int j = 1; j < 4; ++j) {                                 // column w4 + j also needs tile common + j
            const int t = common + j;
            if (t >= t0 && t <= t1) {
                while (*v_done < t - t0 + 1) scan_backoff();
                pull_tiles(t, t, w4, acc);
            }
            publish(w4 + j, acc[j]);
        }
    }
}

// keep_out[g*Kmax + s0 + i] (i < keep_cnt of the segment): list position -> map[position]
__global__ void __launch_bounds__(256)
nms_map_kernel(const int* __restrict__ seg_off, int segs_per_group, int Kmax, const int* __restrict__ map,
               int* __restrict__ keep_out, const int* __restrict__ keep_cnt) {
    RR_PDL_PROLOGUE();
    const int g = blockIdx.z, s = blockIdx.y;
    const int s0 = seg_off[g * (segs_per_group + 1) + s];
    const int n = keep_cnt[g * segs_per_group + s];
    int* out = keep_out + (size_t)g * Kmax + s0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = map[(size_t)g * Kmax + out[i]];
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
    RR_PDL_PROLOGUE();
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
    const int T = (Kmax + 63) / 64, W = mask_row_words(Kmax);
    const long long pairs = (long long)T * (T + 1) / 2;
    if (pairs > 0x7fffffffLL) return RR_E_RANGE;
    dim3 gm((unsigned)pairs, (unsigned)G);
    launch_pdl(nms_mask_kernel, dim3(gm), dim3(kMaskThreads), 0, st, sbox, slab, n, Kmax, W, T, thr, pixel_offset ? 1.0f : 0.0f,
                                       ge_cmp ? 1 : 0, mask);
    RR_LAUNCHED_K(rc, "nms_mask_kernel", st);
    dim3 gs((unsigned)segs_per_group, (unsigned)G);
    size_t smem = (size_t)W * (2 * sizeof(unsigned long long) + sizeof(int));
    if (smem > 16 * 1024) {
        if (smem > 200 * 1024) return RR_E_RANGE;
        RR_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
    }
    launch_pdl(nms_scan_kernel, dim3(gs), dim3(kScanThreads2), smem, st, mask, seg_off, segs_per_group, Kmax, W, keep_out, keep_cnt);
    RR_LAUNCHED_K(rc, "nms_scan_kernel", st);
    if (map) {                                       // kept list positions -> the caller's row indices
        dim3 gmap((unsigned)min((n + 255) / 256, 64), (unsigned)segs_per_group, (unsigned)G);
        launch_pdl(nms_map_kernel, dim3(gmap), dim3(256), 0, st, seg_off, segs_per_group, Kmax, map, keep_out, keep_cnt);
        RR_LAUNCHED_K(rc, "nms_map_kernel", st);
    }
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
    w.mask = cv.take<unsigned long long>((size_t)B * K * mask_row_words(K));
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
    launch_pdl(stage1_partition_kernel, dim3(B), dim3(1024), smem, st, dets, K, C, w.sbox, w.slab, w.ssrc, w.sscore, w.seg_off);
    RR_LAUNCHED_K(rc, "stage1_partition_kernel", st);
    int r2 = launch_mask_scan(w.sbox, w.slab, w.seg_off, B, C, K, K, thr, 0, 0, w.mask, nullptr,
                              w.keep_pos, w.keep_cnt, st);
    if (rc == 0) rc = r2;
    launch_pdl(stage1_compact_kernel, dim3(B), dim3(1024), 0, st, w.sbox, w.sscore, w.slab, w.seg_off, w.keep_pos, w.keep_cnt,
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
    w.mask = cv.take<unsigned long long>((size_t)M * mask_row_words(M));
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
    launch_pdl(rank_sort_kernel, dim3((int)(((long long)M * kRankSplit + 255) / 256)), dim3(256), 0, st, boxes, scores, seg_offsets, M, S, w.sbox, w.slab, w.ssrc);
    RR_LAUNCHED_K(rc, "rank_sort_kernel", st);
    int r2 = launch_mask_scan(w.sbox, w.slab, seg_offsets, 1, S, M, M, thr, pixel_offset, ge_cmp, w.mask,
                              w.ssrc, keep_idx, keep_count, st);
    return rc ? rc : r2;
}

#ifdef RR_SCAN_TRACE
RR_API int rr_debug_scan_trace(long long* host) { return (int)cudaMemcpyFromSymbol(host, rr::g_scan_trace, sizeof(long long) * 8); }
#endif

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
