// Heat-map decode for sm_100a: per-image top-K over C*H*W logits, wh/offset gather, box assembly.
//
// Replaces RRNet.transform_bbox / _topk / _gather_feat / _transpose_and_gather_feat
// (models/rrnet.py:83-138): sigmoid over the full map, two torch.topk calls (16-pass radix
// select over 80 slices), two full NCHW->NHWC permute copies and five gathers become
//
//   1. decode_sample_kernel   one CTA per image reads ~16K pseudo-random samples and picks a
//                             threshold key whose expected population count is ~3K;
//   2. decode_collect_kernel  ONE streaming pass over the logits (128-bit loads, all SMs):
//                             elements with key >= threshold are appended to a per-image
//                             candidate list (shared-memory staging, one global atomic per CTA);
//   3. decode_select_kernel   one CTA per image sorts the candidates (bitonic, shared memory),
//                             takes the first K, gathers wh/offset with direct strided loads
//                             (no permute copy) and writes [B,K,6] + inds.
//
// The ranking key is the LOGIT (monotone bit transform); sigmoid is evaluated for the K
// survivors only.  Ties are ordered by ascending flat index.  The sample threshold is only a
// speed-up: if the candidate count falls outside [K, CAP] (adversarial data, massive ties,
// pool=3 on smooth maps) step 3 falls back to an exact in-CTA radix descent over the whole
// image, so the result is exact for every input.
//
// HBM traffic: the logits once (B*C*H*W*4 bytes) + K*4 gathers + outputs.
// Built with --fmad=false (box arithmetic must round like the reference's separate torch ops).
#include "rr_decode.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>
namespace cg = cooperative_groups;

namespace rr {

// value used for ranking: the logit, or -inf when pool=3 and the element is not a 3x3 peak
// (operators/centernet_operator.py:204-210: keep = (maxpool3x3(heat) == heat)).
__device__ __forceinline__ float pooled_value(const float* __restrict__ img, int H, int W, int HW,
                                              unsigned flat, float v) {
    int c = flat / HW, ind = flat - c * HW;
    int y = ind / W, x = ind - y * W;
    const float* plane = img + (size_t)c * HW;
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            if (__ldg(plane + yy * W + xx) > v) return -CUDART_INF_F;
        }
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// 1. threshold estimate from a pseudo-random sample (one CTA per image)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampleThreads)
decode_sample_kernel(const float* __restrict__ hm, int H, int W, int N, int K, int pool,
                     unsigned int* __restrict__ thr_key, unsigned int* __restrict__ count) {
    RR_PDL_PROLOGUE();
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) count[b] = 0;
    if (N <= kCap) {                   // everything fits in the candidate list: no threshold
        if (tid == 0) thr_key[b] = 0u;
        return;
    }
    const float* img = hm + (size_t)b * N;
    const int HW = H * W;
    const unsigned nlines = (unsigned)N / 32u;
    const int warp = tid >> 5, lane = tid & 31;
    // target population count T in (K, kCap]: centred so that the candidate count lands in (K, 4096] with
    // ~3 sigma on either side at config 2 (the select kernel sorts the next power of two)
    int target = max(2 * K, K + 1300);
    target = min(target, (K + kCap) / 2);
    // samples per thread so that the sample rank of the threshold, r = T * S / N, is about 128
    int spt = (int)((128ll * N + (long long)target * kSampleThreads - 1) / ((long long)target * kSampleThreads));
    spt = min(max(spt, 1), kSamplesPerThread);
    const int S = kSampleThreads * spt;
    int r = (int)(((long long)target * S + N - 1) / N);
    r = min(max(r, 8), kSampleThreads / 2);
    // Every thread keeps the MAXIMUM of its samples; the r-th largest of the 1024 maxima estimates the
    // r-th largest sample (the top r << 1024 samples almost surely sit in different threads; a collision
    // only lowers the threshold slightly, i.e. a few more candidates).  The estimate need not be exact:
    // the select kernel falls back to an exact search when the candidate count leaves [K, kCap].
    unsigned m = 0u;
    if (pool != 3) {                                   // all loads of a thread are issued before the first use
        float v[kSamplesPerThread];
#pragma unroll
        for (int it = 0; it < kSamplesPerThread; ++it) {
            const unsigned L = (unsigned)(warp * spt + min(it, spt - 1));
            const unsigned line = __umulhi(L * 2654435761u + 12345u, nlines);     // hash -> [0, nlines), no 64-bit modulo
            v[it] = __ldg(img + line * 32u + (unsigned)lane);
        }
#pragma unroll
        for (int it = 0; it < kSamplesPerThread; ++it) m = max(m, f2key(v[it]));  // it >= spt repeats the last sample
    } else {
        for (int it = 0; it < spt; ++it) {
            const unsigned L = (unsigned)(warp * spt + it);
            const unsigned line = __umulhi(L * 2654435761u + 12345u, nlines);
            const unsigned flat = line * 32u + (unsigned)lane;
            m = max(m, f2key(pooled_value(img, H, W, HW, flat, __ldg(img + flat))));
        }
    }
    // largest t with #{threads : m >= t} >= r: the 1024 maxima go to shared memory and ONE warp runs the
    // MSB-first descent on 32 values per lane (no block barrier per bit)
    __shared__ unsigned s_max[kSampleThreads];
    s_max[tid] = m;
    __syncthreads();
    if (warp != 0) return;
    unsigned mine[kSampleThreads / 32];
#pragma unroll
    for (int q = 0; q < kSampleThreads / 32; ++q) mine[q] = s_max[q * 32 + lane];
    unsigned t = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned trial = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int q = 0; q < kSampleThreads / 32; ++q) c += (mine[q] >= trial);
        if (__reduce_add_sync(0xffffffffu, c) >= r) t = trial;
    }
    if (tid == 0) thr_key[b] = t;
}

// ---------------------------------------------------------------------------------------------
// 2. streaming collect: grid (ctas_per_image, B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCollectThreads)
decode_collect_kernel(const float* __restrict__ hm, int H, int W, int N, int pool,
                      const unsigned int* __restrict__ thr_key, unsigned int* __restrict__ count,
                      unsigned long long* __restrict__ cand) {
    RR_PDL_PROLOGUE();
    __shared__ unsigned long long s_stage[kStage];
    __shared__ int s_n;
    __shared__ unsigned s_base;
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* img = hm + (size_t)b * N;
    const unsigned thr = thr_key[b];
    unsigned int* g_count = count + b;
    unsigned long long* g_cand = cand + (size_t)b * kCap;
    const int HW = H * W;
    if (tid == 0) s_n = 0;
    __syncthreads();

    const bool vec_ok = ((N & 3) == 0) && ((((uintptr_t)img) & 15) == 0);
    const int stride = gridDim.x * blockDim.x;
    const int gtid = blockIdx.x * blockDim.x + tid;
    if (vec_ok) {
        const float4* p4 = reinterpret_cast<const float4*>(img);
        const int n4 = N >> 2;
        int i = gtid;
        // 4 independent 128-bit loads in flight per thread
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_stream_f4(p4 + i + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unsigned key = f2key(e[q]);
                    if (key >= thr) {
                        unsigned flat = (unsigned)(i + u * stride) * 4u + q;
                        if (pool == 3) {
                            key = f2key(pooled_value(img, H, W, HW, flat, e[q]));
                            if (key < thr) continue;
                        }
                        push_candidate(key, flat, s_stage, &s_n, g_count, g_cand);
                    }
                }
            }
        }
        for (; i < n4; i += stride) {
            float4 v = ld_stream_f4(p4 + i);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned key = f2key(e[q]);
                if (key >= thr) {
                    unsigned flat = (unsigned)i * 4u + q;
                    if (pool == 3) {
                        key = f2key(pooled_value(img, H, W, HW, flat, e[q]));
                        if (key < thr) continue;
                    }
                    push_candidate(key, flat, s_stage, &s_n, g_count, g_cand);
                }
            }
        }
    } else {
        for (int i = gtid; i < N; i += stride) {
            float e = __ldg(img + i);
            unsigned key = f2key(e);
            if (key >= thr) {
                if (pool == 3) {
                    key = f2key(pooled_value(img, H, W, HW, (unsigned)i, e));
                    if (key < thr) continue;
                }
                push_candidate(key, (unsigned)i, s_stage, &s_n, g_count, g_cand);
            }
        }
    }
    __syncthreads();
    const int n = min(s_n, kStage);
    if (n == 0) return;
    if (tid == 0) s_base = atomicAdd(g_count, (unsigned)n);
    __syncthreads();
    const unsigned base = s_base;
    for (int k = tid; k < n; k += blockDim.x) {
        unsigned pos = base + k;
        if (pos < (unsigned)kCap) g_cand[pos] = s_stage[k];
    }
}

// ---------------------------------------------------------------------------------------------
// 3. select + sort + gather + box assembly (one CTA per image)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum_int(int v, int* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                     // protect s_red from the previous use
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int tot = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    return tot;
}

// Exact selection over the whole image by one CTA: K-th largest key by MSB-first descent, then
// an index-ordered collection (all keys > kth, and the lowest-index keys == kth).  Fills s_e[0..K).
__device__ void exact_select(const float* __restrict__ img, int H, int W, int N, int K, int pool,
                             unsigned long long* s_e, int* s_red) {
    const int HW = H * W, tid = threadIdx.x, nthr = blockDim.x;
    unsigned kth = 0;
    for (int bit = 31; bit >= 0; --bit) {
        unsigned trial = kth | (1u << bit);
        int c = 0;
        for (int i = tid; i < N; i += nthr) {
            float v = __ldg(img + i);
            unsigned key = f2key(v);
            if (key >= trial && pool == 3) key = f2key(pooled_value(img, H, W, HW, (unsigned)i, v));
            c += (key >= trial);
        }
        if (block_sum_int(c, s_red) >= K) kth = trial;
    }
    int c_gt = 0;
    for (int i = tid; i < N; i += nthr) {
        float v = __ldg(img + i);
        unsigned key = f2key(v);
        if (key > kth && pool == 3) key = f2key(pooled_value(img, H, W, HW, (unsigned)i, v));
        c_gt += (key > kth);
    }
    const int n_gt = block_sum_int(c_gt, s_red);
    const int need_eq = K - n_gt;        // >= 1 by definition of kth
    __shared__ int s_fill, s_eq_seen;
    if (tid == 0) { s_fill = 0; s_eq_seen = 0; }
    __syncthreads();
    // index-ordered sweep in chunks of nthr so that "first need_eq equal keys" is deterministic
    for (int base = 0; base < N; base += nthr) {
        int i = base + tid;
        unsigned key = 0;
        bool gt = false, eq = false;
        if (i < N) {
            float v = __ldg(img + i);
            key = f2key(v);
            if (key >= kth && pool == 3) key = f2key(pooled_value(img, H, W, HW, (unsigned)i, v));
            gt = key > kth;
            eq = key == kth;
        }
        // rank of this thread among the chunk's equal keys
        unsigned m = __ballot_sync(0xffffffffu, eq);
        const int lane = tid & 31, warp = tid >> 5;
        int my = __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (lane == 0) s_red[warp] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int w2 = 0; w2 < warp; ++w2) before += s_red[w2];
        int chunk_eq = 0;
        for (int w2 = 0; w2 < (nthr >> 5); ++w2) chunk_eq += s_red[w2];
        const int seen = s_eq_seen;
        if (gt || (eq && seen + before + my < need_eq)) {
            int slot = atomicAdd(&s_fill, 1);
            s_e[slot] = ((unsigned long long)key << 32) | (unsigned long long)(~(unsigned)i);
        }
        __syncthreads();
        if (tid == 0) s_eq_seen = seen + chunk_eq;
        __syncthreads();
    }
}

// In-place cut of s_e[0..n) to its K largest entries (all entries distinct): MSB-first radix descent, one
// byte per round (256-bin histogram of the entries that match the prefix so far), then a stable compaction
// through `tmp`.  Returns K.  Block-wide; s_e and tmp live in shared memory.
__device__ int select_top_k(unsigned long long* s_e, int n, int K, unsigned long long* tmp) {
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_fill, s_done;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_prefix = 0ull; s_need = K; s_fill = 0; s_done = 0; }
    for (int round = 0; round < 8; ++round) {
        const int shift = 56 - 8 * round;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = round == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i0 = tid - lane; i0 < n; i0 += nthr) {        // warp-aggregated: one atomic per distinct bin per warp
            const int i = i0 + lane;
            const unsigned long long e = i < n ? s_e[i] : 0ull;
            const bool act = i < n && (e & himask) == prefix;
            const unsigned bin = act ? (unsigned)((e >> shift) & 255ull) : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (act && lane == __ffs(peers) - 1) atomicAdd(&s_hist[bin], __popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            int loc[8], tot = 0;                               // lane owns bins [8*lane, 8*lane+8)
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = s_hist[8 * lane + q]; tot += loc[q]; }
            int suf = tot;                                     // inclusive suffix scan over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += v;
            }
            const int need = s_need;
            int run = suf - tot, pick = -1, need_in = 0;       // entries in the bins of higher lanes
#pragma unroll
            for (int q = 7; q >= 0; --q) {
                if (pick < 0 && run < need && run + loc[q] >= need) { pick = 8 * lane + q; need_in = need - run; }
                run += loc[q];
            }
            if (pick >= 0) {                                   // exactly one lane (1 <= need <= matching entries)
                s_prefix = prefix | ((unsigned long long)pick << shift);
                s_need = need_in;
                // the whole bin is wanted (always so once it holds a single entry): every entry >= the prefix padded
                // with zero bits is in the cut, the remaining bytes need not be looked at
                if (need_in == s_hist[pick]) s_done = 1;
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    const unsigned long long kth = s_prefix;                   // the K-th largest entry itself
    for (int i0 = tid - lane; i0 < n; i0 += nthr) {            // order is irrelevant: sorted next
        const int i = i0 + lane;
        const unsigned long long e = i < n ? s_e[i] : 0ull;
        const bool keep = i < n && e >= kth;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_fill, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) tmp[base + __popc(m & ((1u << lane) - 1u))] = e;
    }
    __syncthreads();
    for (int i = tid; i < K; i += nthr) s_e[i] = tmp[i];
    __syncthreads();
    return K;
}

// Bitonic sort, descending, P a power of two.  Stages whose partner distance j is <= 32 only touch one
// aligned 64-element block per warp iteration, so a warp runs them back to back with __syncwarp only; block
// barriers are needed for the j >= 64 stages (21 instead of 78 barriers at P = 4096).
__device__ __forceinline__ void bitonic_ce(unsigned long long* s_e, int t, int j, int k) {
    const int i = 2 * t - (t & (j - 1));                // lower index of the pair
    const int ixj = i + j;
    const bool desc = ((i & k) == 0);
    const unsigned long long a = s_e[i], c = s_e[ixj];
    if (desc ? (a < c) : (a > c)) { s_e[i] = c; s_e[ixj] = a; }
}

__device__ void bitonic_sort_desc(unsigned long long* s_e, int P) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        int j = k >> 1;
        for (; j > 32; j >>= 1) {
            for (int t = tid; t < half; t += nthr) bitonic_ce(s_e, t, j, k);
            __syncthreads();
        }
        for (int t0 = tid - lane; t0 < half; t0 += nthr) {       // pairs t0..t0+31 <-> elements [2*t0, 2*t0+64)
            for (int jj = j; jj > 0; jj >>= 1) {
                if (t0 + lane < half) bitonic_ce(s_e, t0 + lane, jj, k);
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// one output row: candidate entry e at rank k of image b -> box assembly (models/rrnet.py:117-138) + index
__device__ __forceinline__ void decode_emit(unsigned long long e, int k, int b, const float* __restrict__ whb,
                                            const float* __restrict__ ofb, int HW, int W, int K, int raw,
                                            float* __restrict__ out_dets, long long* __restrict__ out_inds) {
    const unsigned key = (unsigned)(e >> 32);
    const unsigned flat = ~(unsigned)(e & 0xffffffffull);
    const int cls = flat / HW, ind = flat - cls * HW;
    const int yi = ind / W, xi = ind - yi * W;                  // models/rrnet.py:99-100
    const float logit = key2f(key);
    // raw: the map already holds scores (RRNet._topk on its own, :93-109); a pooled-out entry is heat*0
    const float score = (logit == -CUDART_INF_F) ? 0.0f : (raw ? logit : sigmoid_f32(logit));   // :119
    const float xs = ofb ? __fadd_rn((float)xi, __ldg(ofb + ind)) : (float)xi;         // :126
    const float ys = ofb ? __fadd_rn((float)yi, __ldg(ofb + HW + ind)) : (float)yi;    // :127
    float w = whb ? __ldg(whb + ind) : 0.f, h = whb ? __ldg(whb + HW + ind) : 0.f;
    w = (w < 0.0f) ? 0.0f : w;                                   // :128 clamp(min=0), NaN passes
    h = (h < 0.0f) ? 0.0f : h;
    const float px = __fsub_rn(xs, __fmul_rn(w, 0.5f));          // :133 (w/2 is exact)
    const float py = __fsub_rn(ys, __fmul_rn(h, 0.5f));          // :134
    float* o = out_dets + ((size_t)b * K + k) * 6;
    o[0] = px; o[1] = py;
    o[2] = __fadd_rn(w, px);                                     // :137 pred_w + pred_x
    o[3] = __fadd_rn(h, py);
    o[4] = score;
    o[5] = (float)cls;
    if (out_inds) out_inds[(size_t)b * K + k] = ind;
}

__global__ void __launch_bounds__(kSelectThreads)
decode_select_kernel(const float* __restrict__ hm, const float* __restrict__ wh,
                     const float* __restrict__ off, int C, int H, int W, int K, int pool, int raw,
                     const unsigned int* __restrict__ count, const unsigned long long* __restrict__ cand,
                     float* __restrict__ out_dets, long long* __restrict__ out_inds) {
    RR_PDL_PROLOGUE();
    extern __shared__ unsigned long long s_e[];          // [P], P = pow2 >= candidates
    __shared__ int s_red[kSelectThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = H * W, N = C * HW;
    const float* img = hm + (size_t)b * N;
    const unsigned cnt = count[b];
    int n;
    if (cnt >= (unsigned)K && cnt <= (unsigned)kCap) {
        n = (int)cnt;
        const unsigned long long* src = cand + (size_t)b * kCap;
        for (int i = tid; i < n; i += blockDim.x) s_e[i] = src[i];
    } else {
        exact_select(img, H, W, N, K, pool, s_e, s_red);
        n = K;
    }
    // more than 2048 candidates: cut to exactly the K largest first (entries are unique 64-bit values, so an
    // 8-round byte-wise radix descent finds the K-th one exactly), then sort half as many elements
    if (n > 2048 && n <= kCap / 2 && K < n) {
        __syncthreads();
        n = select_top_k(s_e, n, K, s_e + kCap / 2);
    }
    int P = 1;
    while (P < n) P <<= 1;
    __syncthreads();
    for (int i = n + tid; i < P; i += blockDim.x) s_e[i] = 0ull;   // below every real entry
    __syncthreads();
    bitonic_sort_desc(s_e, P);

    const float* whb = wh ? wh + (size_t)b * 2 * HW : nullptr;
    const float* ofb = off ? off + (size_t)b * 2 * HW : nullptr;
    for (int k = tid; k < K; k += blockDim.x) decode_emit(s_e[k], k, b, whb, ofb, HW, W, K, raw, out_dets, out_inds);
}

// ---------------------------------------------------------------------------------------------
// 3b. the same selection by a CLUSTER of kSelCtas CTAs per image (default).  One CTA per image is a latency chain
// (load, radix cut, 2048-element bitonic sort, gather: 38 us at config 2 on 8 of 148 SMs).  Here every CTA of the
// cluster takes an eighth of the image's candidates, sorts it in shared memory, the CTAs exchange their sorted runs
// through distributed shared memory, and every element finds its global rank = its own position + the number of larger
// elements in each of the other runs (bisection: entries are unique 64-bit values, so ranks are a permutation).  Rank
// < K is the output row.  No cut to K, no merge passes.  A candidate count outside [K, kCap] (the sample threshold
// misfired) is handled by rank 0 alone with the exact single-CTA path above.
// ---------------------------------------------------------------------------------------------
constexpr int kSelCtas = 8;
constexpr int kSelThreads = 512;
constexpr int kSelRun = kCap / kSelCtas;             // largest run per CTA (2048 entries)

__global__ void __launch_bounds__(kSelThreads, 1)
decode_select_cluster_kernel(const float* __restrict__ hm, const float* __restrict__ wh,
                             const float* __restrict__ off, int C, int H, int W, int K, int pool, int raw,
                             const unsigned int* __restrict__ count, const unsigned long long* __restrict__ cand,
                             float* __restrict__ out_dets, long long* __restrict__ out_inds) {
    RR_PDL_PROLOGUE();
    extern __shared__ unsigned long long s_all[];        // [kCap] all runs after the exchange (also the fallback's buffer)
    __shared__ int s_red[kSelectThreads / 32];
    unsigned long long* s_run = s_all + kCap;            // [kSelRun] this CTA's sorted run
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / kSelCtas, tid = threadIdx.x;
    const int HW = H * W, N = C * HW;
    const unsigned cnt = count[b];
    const float* whb = wh ? wh + (size_t)b * 2 * HW : nullptr;
    const float* ofb = off ? off + (size_t)b * 2 * HW : nullptr;
    if (!(cnt >= (unsigned)K && cnt <= (unsigned)kCap)) {          // cluster-uniform
        if (rank == 0) {                                           // exact selection over the whole image by one CTA
            exact_select(hm + (size_t)b * N, H, W, N, K, pool, s_all, s_red);
            int P = 1;
            while (P < K) P <<= 1;
            __syncthreads();
            for (int i = K + tid; i < P; i += blockDim.x) s_all[i] = 0ull;
            __syncthreads();
            bitonic_sort_desc(s_all, P);
            for (int k = tid; k < K; k += blockDim.x) decode_emit(s_all[k], k, b, whb, ofb, HW, W, K, raw, out_dets, out_inds);
        }
        return;                                                    // no CTA touches another one's shared memory on this path
    }
    const int n = (int)cnt;
    const int per = (n + kSelCtas - 1) / kSelCtas;                 // run length (the last run may be shorter)
    const int lo = min(rank * per, n), mine = min(per, n - lo);
    int P = 32;
    while (P < per) P <<= 1;
    const unsigned long long* src = cand + (size_t)b * kCap + lo;
    for (int i = tid; i < P; i += blockDim.x) s_run[i] = i < mine ? src[i] : 0ull;     // 0 is below every real entry
    __syncthreads();
    bitonic_sort_desc(s_run, P);
    cluster.sync();                                                // every run is sorted
    for (int r = 0; r < kSelCtas; ++r) {                           // gather the runs (own one included) into s_all
        const unsigned long long* remote = cluster.map_shared_rank(s_run, r);
        const int nr = min(per, max(n - r * per, 0));
        for (int i = tid; i < nr; i += blockDim.x) s_all[r * per + i] = remote[i];
    }
    cluster.sync();                                                // all remote reads are done: CTAs may run ahead / exit
    for (int j = tid; j < mine; j += blockDim.x) {
        const unsigned long long e = s_all[rank * per + j];
        int rk = j;
#pragma unroll 1
        for (int r = 0; r < kSelCtas; ++r) {
            if (r == rank) continue;
            const unsigned long long* run = s_all + r * per;
            int a = 0, z = min(per, max(n - r * per, 0));          // first index with run[idx] < e  (= entries larger than e)
            while (a < z) {
                const int mid = (a + z) >> 1;
                if (run[mid] > e) a = mid + 1; else z = mid;
            }
            rk += a;
        }
        if (rk < K) decode_emit(e, rk, b, whb, ofb, HW, W, K, raw, out_dets, out_inds);
    }
}

int g_select_single_cta = 0;     // rr_set_option("select_single_cta", 1): the one-CTA-per-image selection kernel

int decode_launch(const float* hm, const float* wh, const float* off, int B, int C, int H, int W,
                  int K, int mode, float* out_dets, int64_t* out_inds, void* ws, cudaStream_t st) {
    int rc = 0;
    const int pool = mode & 0xff, raw = (mode & RR_DECODE_RAW_SCORES) ? 1 : 0;
    DecodeWs w = carve_decode(ws, B);
    const int N = C * H * W;
    if (!(mode & RR_DECODE_PRECOLLECTED)) {        // else: rr_hm_tail_collect already left count / cand in the workspace
        launch_pdl(decode_sample_kernel, dim3(B), dim3(kSampleThreads), 0, st, hm, H, W, N, K, pool, w.thr_key, w.count);
        RR_LAUNCHED_K(rc, "decode_sample_kernel", st);
        // fill the machine: ~8 CTAs of 256 threads per SM in total, at least one per image
        int per_img = max(1, (kSMs * 8 + B - 1) / B);
        int max_useful = max(1, (N / 4 + kCollectThreads - 1) / kCollectThreads);
        per_img = min(per_img, max_useful);
        dim3 gc((unsigned)per_img, (unsigned)B);
        launch_pdl(decode_collect_kernel, dim3(gc), dim3(kCollectThreads), 0, st, hm, H, W, N, pool, w.thr_key, w.count, w.cand);
        RR_LAUNCHED_K(rc, "decode_collect_kernel", st);
    }
    static OncePerDevice attr_once; int attr_dev;
    const size_t smem = (size_t)kCap * sizeof(unsigned long long);
    const size_t smem_c = (size_t)(kCap + kSelRun) * sizeof(unsigned long long);
    if (attr_once.need(&attr_dev)) {
        RR_CUDA(cudaFuncSetAttribute(decode_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
        RR_CUDA(cudaFuncSetAttribute(decode_select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c), rc);
        if (rc == 0) attr_once.mark(attr_dev);
    }
    if (g_select_single_cta) {
        launch_pdl(decode_select_kernel, dim3(B), dim3(kSelectThreads), smem, st, hm, wh, off, C, H, W, K, pool, raw, w.count, w.cand,
                   out_dets, (long long*)out_inds);
        RR_LAUNCHED_K(rc, "decode_select_kernel", st);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * kSelCtas)); cfg.blockDim = dim3(kSelThreads); cfg.dynamicSmemBytes = smem_c; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kSelCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const unsigned int* cnt_p = w.count;
        const unsigned long long* cand_p = w.cand;
        long long* inds_p = (long long*)out_inds;
        (void)cudaLaunchKernelEx(&cfg, decode_select_cluster_kernel, hm, wh, off, C, H, W, K, pool, raw, cnt_p, cand_p, out_dets, inds_p);
        RR_LAUNCHED_K(rc, "decode_select_cluster_kernel", st);
    }
    return rc;
}

size_t decode_ws_bytes(int B) { return carve_decode(nullptr, B).bytes; }

}  // namespace rr

using namespace rr;

RR_API size_t rr_decode_workspace_bytes(int B, int C, int H, int W, int K) {
    (void)C; (void)H; (void)W; (void)K;
    if (B <= 0) return 0;
    return decode_ws_bytes(B);
}

RR_API int rr_decode_topk(const float* hm, const float* wh, const float* off,
                          int B, int C, int H, int W, int K, int pool,
                          float* out_dets, int64_t* out_inds,
                          void* ws, size_t ws_bytes, void* stream) {
    if (!hm || !out_dets || !ws) return RR_E_BADARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0) return RR_E_BADARG;
    const int pl = pool & ~(RR_DECODE_RAW_SCORES | RR_DECODE_PRECOLLECTED);
    if (pl != 0 && pl != 3) return RR_E_BADARG;
    if ((pool & RR_DECODE_PRECOLLECTED) && pl != 0) return RR_E_BADARG;      // the fused tail has no 3x3 peak pooling
    if (!(pool & RR_DECODE_RAW_SCORES) && (!wh || !off)) return RR_E_BADARG;   // wh/off optional only for plain top-K
    if ((long long)C * H * W >= (1LL << 31)) return RR_E_RANGE;
    if (K > RR_MAX_TOPK || (long long)K > (long long)H * W) return RR_E_RANGE;   // torch.topk raises (:96)
    if (ws_bytes < decode_ws_bytes(B) || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    return decode_launch(hm, wh, off, B, C, H, W, K, pool, out_dets, out_inds, ws, (cudaStream_t)stream);
}
