// Re-regression head on the 5th-generation tensor cores (tcgen05, sm_100a), fp32 accuracy through a
// 3xTF32 split.
//
// Same contract as head_forward_kernel (rr_head.cu): Bottleneck(256,64) -> avg-pool -> 1x1 conv 256->4 on
// the RoIs' [256,3,3] features (models/rrnet.py:155-157, detectors/fasterrcnn_detector.py:13-18,
// backbones/resnet.py:33-53), BatchNorms folded.  The three convolutions are GEMMs over M = 128 rows
// (14 RoIs x 9 pixels + 2 zero rows) per CTA:
//     conv1  [128 x 256] x [256 x  64]                      8 K-chunks of 32
//     conv2  9 taps x [128 x 64] x [64 x 64]                the A rows of a tap are the 3x3-shifted rows of t1
//     conv3  [128 x  64] x [64 x 256]                       4 N-quarters x 2 K-chunks
// Every MMA is tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 64, K = 8, issued by one thread; the
// accumulators live in TMEM (conv1: columns 0-63, conv2: 64-127, conv3: 128-383) and come back with
// tcgen05.ld for the bias / ReLU / pooling epilogues.  The residual x is added by the tensor core too: while
// a K chunk of x sits in shared memory for conv1, one more MMA against a 32 x 32 identity tile deposits it in
// conv3's accumulator columns, so x is read from memory exactly once.  An fp32 product a*b is evaluated as
// a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with hi = tf32(v), lo = tf32(v - hi): three MMAs per K-step,
// error ~2^-21 relative per product, so the result stays within the 1e-5 parity budget.
//
// Operands are K-major tiles of 32 tf32 per row (128 bytes) with the 128-byte swizzle, written by ordinary
// stores: A tiles (activations, split on the fly) by all threads into two alternating stages, B tiles
// (weights) copied with cp.async from a pre-split, pre-swizzled image made once by rr_head_fold into a
// ring of four slots, two steps ahead of their use.  While the tensor core works on one step the threads
// fill the next; a tcgen05.commit on a per-stage mbarrier tells when a stage / slot may be overwritten.
#include "rr_head.cuh"

namespace rr {

constexpr int kTcRois = 14;                      // RoIs per CTA: 126 of the 128 MMA rows
constexpr int kTcThreads = 512;
constexpr int kTcBlock = kTcThreads;
constexpr int kTcItems = 1024 / kTcThreads;      // (row, 16-byte chunk) items of an A tile per thread
constexpr int kTcATile = 128 * 128;              // bytes: 128 rows x 32 tf32
constexpr int kTcBTile = 64 * 128;               // bytes:  64 rows x 32 tf32
constexpr int kTcAStage = 2 * kTcATile;          // A_hi | A_lo = 32 KB, two stages
constexpr int kTcBSlot = 2 * kTcBTile;           // B_hi | B_lo = 16 KB, ring of four (weights are prefetched 2 steps ahead)
constexpr int kTcBRing = 4;
constexpr int kT1Stride = 68;                    // floats per row of the plain t1 / t2 array (conflict-free 128-bit rows)
constexpr int kTcT1Bytes = 2 * 128 * kT1Stride * 4;   // t1 / t2 as (hi, lo) tf32 planes
constexpr int kTcEyeBytes = 32 * 128;                // 32 x 32 identity tile (residual through the tensor core)
constexpr int kTcPoolBytes = kTcRois * 256 * 4;
constexpr int kTcSmem = 2 * kTcAStage + kTcBRing * kTcBSlot + kTcEyeBytes + kTcT1Bytes + kTcPoolBytes + 1024;   // + slack for the 1024-byte alignment
constexpr int kTmemCols = 512;
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
//                          D = f32      A = tf32     B = tf32      N = 64               M = 128      (both K-major)
constexpr uint32_t kIdescN32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate,
                                          uint32_t idesc = kIdesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
    hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
    lo.x = to_tf32(v.x - hi.x); lo.y = to_tf32(v.y - hi.y); lo.z = to_tf32(v.z - hi.z); lo.w = to_tf32(v.w - hi.w);
}
// 16-byte stores of an already split value into the swizzled A tiles: row r, 16-byte chunk ch
__device__ __forceinline__ void store_tiles(uint8_t* a_hi, uint8_t* a_lo, int r, int ch, float4 hi, float4 lo) {
    const int off = r * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(a_hi + off) = hi;
    *reinterpret_cast<float4*>(a_lo + off) = lo;
}
// (hi, lo) split of four values and their stores into the swizzled A tiles
__device__ __forceinline__ void store_split(uint8_t* a_hi, uint8_t* a_lo, int r, int ch, float4 v) {
    float4 hi, lo;
    split4(v, hi, lo);
    const int off = r * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(a_hi + off) = hi;
    *reinterpret_cast<float4*>(a_lo + off) = lo;
}

// ---- weight image: (hi, lo) tf32 tiles in the swizzled shared-memory layout, made once per fold ----
__global__ void head_fold_tc_kernel(float* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kTcSteps * kTcTileFloats) return;
    const int step = i / kTcTileFloats, e = i - step * kTcTileFloats;
    const int n = e >> 5, k = e & 31;                               // tile row (output channel), K index inside the chunk
    float w;
    if (step < 8) {
        w = f[kOffW1 + (32 * step + k) * 64 + n];
    } else if (step < 26) {
        const int t = step - 8, tap = t >> 1, kc = t & 1;
        w = f[kOffW2 + ((32 * kc + k) * 9 + tap) * 64 + n];
    } else {
        const int t = step - 26, q = t >> 1, kc = t & 1;
        w = f[kOffW3 + (32 * kc + k) * 256 + 64 * q + n];
    }
    const float hi = to_tf32(w), lo = to_tf32(w - hi);
    const int pos = n * 32 + ((((k >> 2) ^ (n & 7))) << 2) + (k & 3);
    float* dst = f + kOffTc + step * kTcStepFloats;
    dst[pos] = hi;
    dst[kTcTileFloats + pos] = lo;
}

int head_fold_tc_launch(float* folded, cudaStream_t st) {
    int rc = 0;
    head_fold_tc_kernel<<<(kTcSteps * kTcTileFloats + 255) / 256, 256, 0, st>>>(folded);
    RR_LAUNCHED(rc);
    return rc;
}

// x[row][c0 .. c0+3] of the CTA's RoI tile: row = roi_local * 9 + pixel
__device__ __forceinline__ float4 tc_load_x4(const HeadSrc& src, int n, int p, int sb, int pieces, float inv, int c0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sb < 0) {
        const float* xn = src.roi_feat + (size_t)n * 2304 + p;
        v.x = __ldg(xn + (c0 + 0) * 9); v.y = __ldg(xn + (c0 + 1) * 9);
        v.z = __ldg(xn + (c0 + 2) * 9); v.w = __ldg(xn + (c0 + 3) * 9);
    } else {
        for (int k = 0; k < pieces; ++k) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src.partial + (size_t)(sb + k) * 2304 + p * 256 + c0));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
    }
    return v;
}

#ifdef RR_HEAD_TC_TRACE      // tools/head_trace.py: per-CTA phase time stamps (never defined in the shipped build)
__device__ unsigned long long g_tc_trace[1024 * 32];
#define TC_TRACE(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tc_trace[blockIdx.x * 32 + (k)] = t_; } } while (0)
#else
#define TC_TRACE(k) do { } while (0)
#endif

__global__ void __launch_bounds__(kTcBlock, 1)
head_tc_kernel(HeadSrc src, const int* __restrict__ n_rois_dev, int n_cap,
               const float* __restrict__ f, float* __restrict__ reg) {
    extern __shared__ uint8_t s_dyn[];
    TC_TRACE(0);
    __shared__ __align__(8) unsigned long long s_free[2];      // stage may be overwritten (its MMAs are done)
    __shared__ __align__(8) unsigned long long s_phase;        // all MMAs of a phase are done
    __shared__ uint32_t s_tmem;
    __shared__ int s_sb[kTcRois], s_pc[kTcRois];
    __shared__ float s_inv[kTcRois];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    const int roi0 = blockIdx.x * kTcRois;
    if (roi0 >= live) return;
    const int nroi = min(kTcRois, live - roi0);

    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)s_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t* stage[2] = {base, base + kTcAStage};                               // A_hi | A_lo
    uint8_t* ring = base + 2 * kTcAStage;                                       // kTcBRing x (B_hi | B_lo)
    float* s_eye = reinterpret_cast<float*>(ring + kTcBRing * kTcBSlot);        // identity B tile, 32 x 32
    float* s_thi = s_eye + kTcEyeBytes / 4;                                     // t1, then t2: tf32 hi plane [128][kT1Stride]
    float* s_tlo = s_thi + 128 * kT1Stride;                                     //              tf32 lo plane
    float* s_pool = s_tlo + 128 * kT1Stride;                                    // [14][256]
    for (int i = tid; i < 32 * 32; i += kTcThreads) {       // I[n][k] in the swizzled tile layout
        const int n = i >> 5, k = i & 31;
        s_eye[n * 32 + ((((k >> 2) ^ (n & 7))) << 2) + (k & 3)] = (n == k) ? 1.0f : 0.0f;
    }

    if (tid < kTcRois) {
        int sb = 0, pc = 0;
        float cnt = 1.f;
        if (tid < nroi) {
            const int n = roi0 + tid;
            if (src.partial) { sb = src.slot[n]; pc = src.pieces[n]; cnt = src.count[n]; }
            else sb = -1;
        }
        s_sb[tid] = sb; s_pc[tid] = pc; s_inv[tid] = 1.0f / cnt;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_free[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_free[1])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_phase)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_TRACE(1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const float* ftc = f + kOffTc;

    int step = 0;                                   // ring position over all 34 weight steps
    auto acquire = [&](int s) {                     // stage s&1 was last used by step s-2
        if (s >= 2) bar_wait(&s_free[s & 1], (uint32_t)(((s >> 1) - 1) & 1));
    };
    // weight tile pair of step s -> ring slot s % 4 (16 KB, straight copy); always commits a group so that the
    // group count stays in step with s even past the last step
    auto prefetch_b = [&](int s) {
#ifdef RR_TC_EXP_NOB
        if (s < 4) {
#else
        if (s < kTcSteps) {
#endif
            const float4* g = reinterpret_cast<const float4*>(ftc + (size_t)s * kTcStepFloats);
            uint8_t* d = ring + (s & (kTcBRing - 1)) * kTcBSlot;
#pragma unroll
            for (int q = 0; q < kTcItems; ++q) {
                const int i = tid + q * kTcThreads;     // 1024 x 16 bytes
                const unsigned da = smem_addr(d + i * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(g + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Every step: acquire(s) (the MMAs of step s-2 are done: its A stage and its B slot are free), prefetch the
    // weights of step s+2 into that slot, fill the A stage, then publish(s): barrier, and thread 0 issues.
    auto publish = [&](int s) {
        asm volatile("cp.async.wait_group 2;" ::: "memory");               // groups s+1, s+2 may still be in flight
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_col = s < 8 ? 0u : (s < 26 ? 64u : 128u + 64u * (uint32_t)((s - 26) >> 1));
            const bool first = s == 0 || s == 8;      // conv3 accumulates onto the residual placed during conv1
            const uint32_t sa = smem_addr(stage[s & 1]), sb = smem_addr(ring + (s & (kTcBRing - 1)) * kTcBSlot);
            const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + kTcATile);
            const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + kTcBTile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {        // K = 32 per stage: four K = 8 instructions, 32 bytes apart
                umma_tf32(tmem + d_col, a_hi + 2 * ks, b_hi + 2 * ks, (first && ks == 0) ? 0u : 1u);
                umma_tf32(tmem + d_col, a_hi + 2 * ks, b_lo + 2 * ks, 1u);
                umma_tf32(tmem + d_col, a_lo + 2 * ks, b_hi + 2 * ks, 1u);
            }
#ifdef RR_TC_EXP_NOEYE
            if (s < 0) {
#else
            if (s < 8) {                            // conv3's accumulator starts as x itself: D3[:, 32s .. 32s+32) = A . I
#endif
                const uint64_t eye = umma_desc(smem_addr(s_eye));
                const uint32_t d3 = tmem + 128u + 32u * (uint32_t)s;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_tf32(d3, a_hi + 2 * ks, eye + 2 * ks, ks == 0 ? 0u : 1u, kIdescN32);
                    umma_tf32(d3, a_lo + 2 * ks, eye + 2 * ks, 1u, kIdescN32);
                }
            }
            umma_commit(&s_free[s & 1]);
            if (s == 7 || s == 25 || s == kTcSteps - 1) umma_commit(&s_phase);
        }
    };
    auto stagers_sync = [&]() { __syncthreads(); };
    prefetch_b(0);
    prefetch_b(1);

    // ============================== conv1: x [128 x 256] . W1 ==============================
    // x is prefetched two K chunks ahead into registers.  A row's value is the sum of its RoI's partial slots:
    // the first four slots of all four items are loaded back to back (16 independent 128-bit loads in flight
    // per thread; a load-add-load-add loop would serialise on the in-order issue), the adds happen at use.
    float4 xpa[kTcItems][4], xpb[kTcItems][4];      // two chunks in flight (even / odd K chunk)
    auto load_x_chunk = [&](int kc, float4 (&xp)[kTcItems][4]) {
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {        // 128 rows x 8 chunks of 4 channels
            const int i = tid + q * kTcThreads, r = i >> 3, ch = i & 7;
            const int rl = r / 9, p = r - rl * 9, c0 = 32 * kc + 4 * ch;
#pragma unroll
            for (int k = 0; k < 4; ++k) xp[q][k] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef RR_TC_EXP_NOX
            if (rl < nroi && kc < 2) {
#else
            if (rl < nroi) {
#endif
                const int sb = s_sb[rl], pc = s_pc[rl];
                if (sb < 0) {
                    xp[q][0] = tc_load_x4(src, roi0 + rl, p, sb, 0, 1.f, c0);
                } else {
                    const float* pp = src.partial + (size_t)sb * 2304 + p * 256 + c0;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < pc) xp[q][k] = __ldg(reinterpret_cast<const float4*>(pp + (size_t)k * 2304));
                }
            }
        }
    };
    auto x_value = [&](int kc, int q, const float4 (&xp)[kTcItems][4]) {   // same summation order as roi_combine_kernel: slot 0, 1, 2, ...
        const int i = tid + q * kTcThreads, r = i >> 3, ch = i & 7;
        const int rl = r / 9, p = r - rl * 9;
        float4 v = xp[q][0];
        if (rl < nroi && s_sb[rl] >= 0) {
            const int sb = s_sb[rl], pc = s_pc[rl];
#pragma unroll
            for (int k = 1; k < 4; ++k) { v.x += xp[q][k].x; v.y += xp[q][k].y; v.z += xp[q][k].z; v.w += xp[q][k].w; }
#ifdef RR_TC_EXP_NOTAIL
            for (int k = 4; k < 0; ++k) {
#else
            for (int k = 4; k < pc; ++k) {          // RoIs cut into more than four pieces are rare
#endif
                const float4 t = __ldg(reinterpret_cast<const float4*>(src.partial + (size_t)(sb + k) * 2304 + p * 256 + 32 * kc + 4 * ch));
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            const float inv = s_inv[rl];
            v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
        }
        return v;
    };
    auto conv1_step = [&](int kc, float4 (&xp)[kTcItems][4]) {
        const int s = step++;
        acquire(s);
        prefetch_b(s + 2);
        uint8_t* a_hi = stage[s & 1];
        uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {
            const int i = tid + q * kTcThreads;
            store_split(a_hi, a_lo, i >> 3, i & 7, x_value(kc, q, xp));
        }
        if (kc + 2 < 8) load_x_chunk(kc + 2, xp);   // in flight across two steps
        publish(s);
        TC_TRACE(16 + kc);
    };
    load_x_chunk(0, xpa);
    TC_TRACE(2);
    load_x_chunk(1, xpb);
#pragma unroll 1
    for (int kc = 0; kc < 8; kc += 2) {
        conv1_step(kc, xpa);
        conv1_step(kc + 1, xpb);
    }
    TC_TRACE(3);
    bar_wait(&s_phase, 0u);
    TC_TRACE(4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 8) {   // t1 = relu(D1 + b1) -> (hi, lo) planes [row][64]; eight warps cover 4 lane quarters x 2 column halves
        const int q = warp & 3, h = warp >> 2, row = 32 * q + lane;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * h), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaxf(v[j] + __ldg(f + kOffB1 + 32 * h + j), 0.f);
            o.y = fmaxf(v[j + 1] + __ldg(f + kOffB1 + 32 * h + j + 1), 0.f);
            o.z = fmaxf(v[j + 2] + __ldg(f + kOffB1 + 32 * h + j + 2), 0.f);
            o.w = fmaxf(v[j + 3] + __ldg(f + kOffB1 + 32 * h + j + 3), 0.f);
            float4 hi, lo;
            split4(o, hi, lo);
            *reinterpret_cast<float4*>(s_thi + row * kT1Stride + 32 * h + j) = hi;
            *reinterpret_cast<float4*>(s_tlo + row * kT1Stride + 32 * h + j) = lo;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    stagers_sync();
    TC_TRACE(5);

    // ============================== conv2: 9 taps, A rows = shifted rows of t1 ==============================
    for (int t = 0; t < 18; ++t) {
        const int s = step++;
        const int tap = t >> 1, kc = t & 1, dy = tap / 3 - 1, dx = tap % 3 - 1;
        acquire(s);
        prefetch_b(s + 2);
        uint8_t* a_hi = stage[s & 1];
        uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {
            const int i = tid + q * kTcThreads, r = i >> 3, ch = i & 7;
            float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
            const int rl = r / 9, p = r - rl * 9, py = p / 3 + dy, px = p % 3 + dx;
            if (rl < kTcRois && py >= 0 && py < 3 && px >= 0 && px < 3) {     // zero padding of the 3x3 map
                const int o = (r + dy * 3 + dx) * kT1Stride + 32 * kc + 4 * ch;
                hi = *reinterpret_cast<const float4*>(s_thi + o);
                lo = *reinterpret_cast<const float4*>(s_tlo + o);
            }
            store_tiles(a_hi, a_lo, r, ch, hi, lo);
        }
        publish(s);
    }
    TC_TRACE(6);
    bar_wait(&s_phase, 1u);
    TC_TRACE(7);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 8) {   // t2 = relu(D2 + b2) -> overwrites t1 (every tap has been staged and consumed)
        const int q = warp & 3, h = warp >> 2, row = 32 * q + lane;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + 64u + (uint32_t)(32 * h), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaxf(v[j] + __ldg(f + kOffB2 + 32 * h + j), 0.f);
            o.y = fmaxf(v[j + 1] + __ldg(f + kOffB2 + 32 * h + j + 1), 0.f);
            o.z = fmaxf(v[j + 2] + __ldg(f + kOffB2 + 32 * h + j + 2), 0.f);
            o.w = fmaxf(v[j + 3] + __ldg(f + kOffB2 + 32 * h + j + 3), 0.f);
            float4 hi, lo;
            split4(o, hi, lo);
            *reinterpret_cast<float4*>(s_thi + row * kT1Stride + 32 * h + j) = hi;
            *reinterpret_cast<float4*>(s_tlo + row * kT1Stride + 32 * h + j) = lo;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    stagers_sync();
    TC_TRACE(8);

    // ============================== conv3: t2 [128 x 64] . W3, four N quarters ==============================
    for (int t = 0; t < 8; ++t) {
        const int s = step++;                       // s & 1 == kc: the A chunk kc stays in stage kc for all quarters
        const int q4 = t >> 1, kc = t & 1;
        acquire(s);
        prefetch_b(s + 2);
        if (q4 == 0) {
            uint8_t* a_hi = stage[s & 1];
            uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
            for (int q = 0; q < kTcItems; ++q) {
                const int i = tid + q * kTcThreads, r = i >> 3, ch = i & 7;
                const int o = r * kT1Stride + 32 * kc + 4 * ch;
                store_tiles(a_hi, a_lo, r, ch, *reinterpret_cast<const float4*>(s_thi + o),
                            *reinterpret_cast<const float4*>(s_tlo + o));
            }
        }
        publish(s);      // accumulates onto the residual placed by conv1
        if (t == 0) TC_TRACE(24); if (t == 6) TC_TRACE(25);
    }
    TC_TRACE(9);
    bar_wait(&s_phase, 0u);
    TC_TRACE(10);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ============================== + b3 + residual, relu, avg-pool over the 9 rows, regressor ==============================
    asm volatile("cp.async.wait_all;" ::: "memory");
    float* s_relu = reinterpret_cast<float*>(stage[0]);            // [2][128][33] = 33 KB over both A stages, idle now
    {
        const int q = warp & 3, h = warp >> 2, row = 32 * q + lane;
        for (int cb = 0; cb < 4; ++cb) {            // warps 0-3: column blocks 2cb, warps 4-7: 2cb+1 (32 columns each)
            if (warp < 8) {
                const int c0 = 32 * (2 * cb + h);
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + 128u + (uint32_t)c0, v);    // W3.t2 + x (residual already inside)
                float* dst = s_relu + (h * 128 + row) * 33;
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[j] = fmaxf(v[j] + __ldg(f + kOffB3 + c0 + j), 0.f);   // resnet.py:49-50
            }
            stagers_sync();
            for (int i = tid; i < 2 * kTcRois * 32; i += kTcThreads) {      // (half, roi, column): mean of 9 rows
                const int hh = i / (kTcRois * 32), rem = i - hh * (kTcRois * 32), r2 = rem >> 5, c = rem & 31;
                const float* sp = s_relu + (hh * 128 + r2 * 9) * 33 + c;
                float acc = 0.f;
#pragma unroll
                for (int pp = 0; pp < 9; ++pp) acc += sp[pp * 33];
                s_pool[r2 * 256 + 32 * (2 * cb + hh) + c] = acc / 9.0f;
            }
            stagers_sync();
        }
    }
    for (int rl = warp; rl < nroi; rl += kTcThreads / 32) {        // regressor 256 -> 4 (+ bias)
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        for (int o = lane; o < 256; o += 32) {
            const float pv = s_pool[rl * 256 + o];
            r0 = fmaf(__ldg(f + kOffWr + o), pv, r0);
            r1 = fmaf(__ldg(f + kOffWr + 256 + o), pv, r1);
            r2 = fmaf(__ldg(f + kOffWr + 512 + o), pv, r2);
            r3 = fmaf(__ldg(f + kOffWr + 768 + o), pv, r3);
        }
        r0 = warp_sum(r0); r1 = warp_sum(r1); r2 = warp_sum(r2); r3 = warp_sum(r3);
        if (lane == 0)
            reinterpret_cast<float4*>(reg)[roi0 + rl] = make_float4(r0 + __ldg(f + kOffBr), r1 + __ldg(f + kOffBr + 1),
                                                                    r2 + __ldg(f + kOffBr + 2), r3 + __ldg(f + kOffBr + 3));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    stagers_sync();
    TC_TRACE(12);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

int head_tc_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded, float* reg, cudaStream_t st) {
    int rc = 0;
    static bool attr_set = false;
    if (!attr_set) {
        RR_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem), rc);
        attr_set = true;
    }
    const int grid = (n_cap + kTcRois - 1) / kTcRois;
    head_tc_kernel<<<grid, kTcBlock, kTcSmem, st>>>(src, n_rois_dev, n_cap, folded, reg);
    RR_LAUNCHED(rc);
    return rc;
}

}  // namespace rr

#ifdef RR_HEAD_TC_TRACE
RR_API int rr_debug_head_trace(unsigned long long* host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, rr::g_tc_trace, sizeof(unsigned long long) * (size_t)n_words);
}
#endif
