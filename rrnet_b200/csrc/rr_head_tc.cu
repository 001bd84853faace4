// Re-regression head on the 5th-generation tensor cores (tcgen05, sm_100a), fp32 accuracy through a
// 3xTF32 split.
//
// Same contract as head_forward_kernel (rr_head.cu): Bottleneck(256,64) -> avg-pool -> 1x1 conv 256->4 on
// the RoIs' [256,3,3] features (models/rrnet.py:155-157, detectors/fasterrcnn_detector.py:13-18,
// backbones/resnet.py:33-53), BatchNorms folded.  A CTA takes 8 RoIs.  Their pixels are the rows of one
// M = 128 tile in a padded list of 16 rows per RoI,
//     row(roi, py, px) = 16 roi + 4 + 4 py + px          (rows 16 roi + 0..3 and every px = 3 row stay zero)
// so that the neighbour (py + dy, px + dx) of a pixel is the row 4 dy + dx further down and every neighbour
// outside the 3x3 map is a zero row.  The three convolutions are GEMMs over that tile:
//     conv1  [128 x 256] x [256 x  64]              8 K-chunks of 32
//     conv2  9 taps x [128 x 64] x [64 x 64]        a tap is the SAME t1 tile read through a shared-memory
//                                                   descriptor whose start address is shifted by 4 dy + dx rows
//                                                   (the 128-byte swizzle is a function of the address, so a
//                                                   128-byte-aligned start is legal: tools/tc_shift_probe.cu)
//     conv3  [128 x  64] x [64 x 256]               2 halves of 2 N-quarters x 2 K-chunks
// Nothing is restaged between the taps: after conv1 the only data movement is the weight stream.
// Every MMA is tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 64, K = 8, both operands in shared memory:
// 48 cycles each, bound by the 6 KB of operands it reads (tools/tc_rate_probe.cu), not by the 32 cycles of
// math.  An fp32 product a*b is evaluated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with hi = tf32(v),
// lo = tf32(v - hi): three MMAs per K-step, error ~2^-21 relative per product, inside the 1e-5 parity budget.
//
// The tensor pipe idles while a tile is loaded and during its epilogues, and the workers idle during conv2 /
// conv3, so TWO CTAs share an SM and fill each other's gaps.  That sets the budget of one CTA: 256 TMEM columns
// (conv1 and conv2 share columns 0-63, a conv3 half has 64-191, the two halves run one after the other),
// <= 113 KB of shared memory (the t1 / t2 tiles reuse the conv1 stages, the weight ring has two slots) and
// <= 88 registers per thread.  The residual x is added in the last epilogue from a per-RoI fp32 copy the workers
// leave in global memory (L2) while they stage conv1.
//
// Warp roles (no block-wide barrier after the prologue; mbarriers connect the roles):
//   warps 0-8   workers: sum the RoIs' partial slots (x), split, fill the two conv1 A stages; epilogues
//   warp  9     one elected lane issues every tcgen05.mma and the tcgen05.commit that frees a stage / ring slot
//   warp  10    lane 0 streams the 34 pre-split, pre-swizzled weight tile pairs (16 KB each, made once by
//               rr_head_fold) with cp.async.bulk (TMA) + complete_tx; the warp also prefetches into L2 the x
//               slots of the CTA that will follow this one on the SM
// The regressor is applied per row before the pooling (both are linear): reg = (sum_p Wr.relu(y_p)) / 9 + br.
#include "rr_head.cuh"

namespace rr {

constexpr int kTcRois = 8;                       // RoIs per CTA
constexpr int kRoiRows = 16;                     // tile rows per RoI: 3 x (3 pixels + 1 zero) + 4 zero rows in front
constexpr int kWorkerWarps = 9;
constexpr int kWorkers = 32 * kWorkerWarps;      // 288 threads x 2 items = the 576 live (row, 16-byte chunk) items of an x tile
constexpr int kTcItems = 2;
constexpr int kTcSlotsInReg = 6;                 // partial slots of an item held in registers per K chunk
constexpr int kTcTiles = 2;                      // M = 128 tiles per CTA, in lockstep: every weight slot serves both
constexpr int kTcBlock = kTcTiles * kWorkers + 64;   // + the MMA issuer warp + the weight stream warp
constexpr int kTcCtasPerSm = 1;
constexpr int kTcATile = 128 * 128;              // bytes: 128 rows x 32 tf32
constexpr int kTcBTile = 64 * 128;               // bytes:  64 rows x 32 tf32
constexpr int kTcAStage = 2 * kTcATile;          // A_hi | A_lo = 32 KB
constexpr int kAStages = 2;
constexpr int kTcBSlot = 2 * kTcBTile;           // B_hi | B_lo = 16 KB = one step of the weight image
constexpr int kTcBRing = 4;
constexpr int kMargin = 8;                       // zero rows above and below the 128 tile rows of a t1 / t2 plane
constexpr int kPlaneBytes = (128 + 2 * kMargin) * 128;
constexpr int kRegion0 = 4 * kPlaneBytes;        // 72 KB: the conv1 stages (64 KB), then the four t planes
static_assert(kRegion0 >= kAStages * kTcAStage, "the t planes reuse the conv1 stages");
static_assert(kTcBSlot == kTcStepFloats * 4, "a ring slot is one step of the folded image");
constexpr int kTcSmem = kTcTiles * kRegion0 + kTcBRing * kTcBSlot + 1024;   // + slack for the 1024-byte alignment
constexpr int kTmemCols = 512;
constexpr uint32_t kTileCols = 192;               // TMEM columns of one tile
constexpr uint32_t kColD12 = 0, kColD3 = 64;     // conv1 / conv2 accumulator, conv3 half accumulator (128 columns)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
//                          D = f32      A = tf32     B = tf32      N = 64               M = 128      (both K-major)

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
// A whole warp waits, one lane polls (with back-off): 300 threads spinning on try_wait would compete with the
// tensor core for the shared-memory port its operands come through.
__device__ __forceinline__ void bar_wait_warp(unsigned long long* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        uint32_t ok;
        for (;;) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
            if (ok) break;
#ifndef RR_TC_EXP_NOSLEEP
            __nanosleep(64);
#endif
        }
    }
    __syncwarp();
}
__device__ __forceinline__ bool elect_one() {        // one lane of a converged warp
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
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


#ifdef RR_HEAD_TC_TRACE      // tools/head_trace.py: per-CTA phase time stamps (never defined in the shipped build)
__device__ unsigned long long g_tc_trace[2048 * 32];
#define TC_TRACE(k) do { if (threadIdx.x == 0 && blockIdx.x < 2048) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tc_trace[blockIdx.x * 32 + (k)] = t_; } } while (0)
#define TC_TRACE_VAL(k, v) do { if (blockIdx.x < 2048) g_tc_trace[blockIdx.x * 32 + (k)] = (unsigned long long)(v); } while (0)
#define TC_CLOCK() clock64()
#else
#define TC_TRACE(k) do { } while (0)
#define TC_TRACE_VAL(k, v) do { } while (0)
#define TC_CLOCK() 0ll
#endif

__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ float4 ld_global_f4(const float* p) {      // coherent load: the data was written by this kernel
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kTcBlock, kTcCtasPerSm)
head_tc_kernel(HeadSrc src, const int* __restrict__ n_rois_dev, int n_cap,
               const float* __restrict__ f, float* __restrict__ reg, int wave) {
    extern __shared__ uint8_t s_dyn[];
    __shared__ __align__(8) unsigned long long s_full_a[kAStages];   // conv1 A stage filled (one arrival per worker warp)
    __shared__ __align__(8) unsigned long long s_free_a[kAStages];   // ... consumed (tcgen05.commit)
    __shared__ __align__(8) unsigned long long s_full_b[kTcBRing];   // weight slot landed (complete_tx)
    __shared__ __align__(8) unsigned long long s_free_b[kTcBRing];   // ... consumed (tcgen05.commit)
    __shared__ __align__(8) unsigned long long s_phase[4];           // all MMAs of conv1 / conv2 / conv3 half 0 / half 1 are done
    __shared__ __align__(8) unsigned long long s_tready;             // t1 (then t2) tiles written by the epilogue
    __shared__ __align__(8) unsigned long long s_d3free;             // conv3 half 0 has been read out of TMEM
    __shared__ uint32_t s_tmem;
    __shared__ int s_sb[kTcTiles][kTcRois];
    __shared__ float s_b1[64], s_b2[64], s_b3[256];
    __shared__ float4 s_wr[256];                                     // regressor weights, one float4 per channel
    __shared__ float4 s_part[kTcTiles][2][kTcRois];
    TC_TRACE(0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    if (blockIdx.x * (kTcTiles * kTcRois) >= live) return;

    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)s_dyn + 1023) & ~(uintptr_t)1023);
    // [0, 72 KB): two conv1 A stages (A_hi | A_lo); after conv1 the same bytes hold the t1 / t2 tiles:
    // plane (hi|lo, kc) at base + (2*lo + kc) * kPlaneBytes, 8 zero rows, the 128 tile rows, 8 zero rows
    uint8_t* ring = base + kTcTiles * kRegion0;                                            // kTcBRing x (B_hi | B_lo)

    // ------------------------------ prologue (all warps) ------------------------------
    for (int i = tid; i < kTcTiles * kRegion0 / 16; i += kTcBlock)              // pad rows of the A stages stay zero
        reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 64; i += kTcBlock) { s_b1[i] = __ldg(f + kOffB1 + i); s_b2[i] = __ldg(f + kOffB2 + i); }
    for (int i = tid; i < 256; i += kTcBlock) {
        s_b3[i] = __ldg(f + kOffB3 + i);
        s_wr[i] = make_float4(__ldg(f + kOffWr + i), __ldg(f + kOffWr + 256 + i), __ldg(f + kOffWr + 512 + i),
                              __ldg(f + kOffWr + 768 + i));
    }
    if (tid >= 64 && tid < 64 + kTcTiles * kTcRois) {
        const int n = blockIdx.x * (kTcTiles * kTcRois) + (tid - 64);
        (&s_sb[0][0])[tid - 64] = (n < live && src.partial) ? __ldg(src.slot + n) : -1;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        auto init = [](unsigned long long* b, int count) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
        };
        for (int i = 0; i < kAStages; ++i) { init(&s_full_a[i], kTcTiles * kWorkerWarps); init(&s_free_a[i], 1); }
        for (int i = 0; i < kTcBRing; ++i) { init(&s_full_b[i], 1); init(&s_free_b[i], 1); }
        for (int i = 0; i < 4; ++i) init(&s_phase[i], 1);
        init(&s_tready, kTcTiles * kWorkerWarps);
        init(&s_d3free, kTcTiles * 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // zero fill -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TRACE(1);
    const uint32_t tmem = s_tmem;
    const float* ftc = f + kOffTc;

    // ------------------------------ warp 10: the weight stream ------------------------------
    if (warp == kTcTiles * kWorkerWarps + 1) {
        auto load_b = [&](int s) {
            const uint32_t bar = smem_addr(&s_full_b[s % kTcBRing]);
            const uint32_t dst = smem_addr(ring + (s % kTcBRing) * kTcBSlot);
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(kTcBSlot) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(ftc + (size_t)s * kTcStepFloats), "r"(kTcBSlot), "r"(bar) : "memory");
        };
        if (lane == 0)
            for (int s = 0; s < kTcBRing; ++s) load_b(s);
        __syncwarp();
        if (lane == 0)
            for (int s = kTcBRing; s < kTcSteps; ++s) {
                bar_wait(&s_free_b[s % kTcBRing], (uint32_t)((s / kTcBRing - 1) & 1));
                load_b(s);
            }
        return;
    }

    // ------------------------------ warp 9: the MMA issuer ------------------------------
    // The warp stays converged and one elected lane issues: the compiler then keeps the descriptors in uniform
    // registers and emits the MMAs back to back (a `lane == 0` branch wraps every tcgen05.mma in an
    // ELECT / BRA.U.ANY loop, and with the per-step index arithmetic the single issuing thread, not the tensor
    // core, set the pace).  The 34 steps are unrolled: ring slots, parities, columns and tap shifts are immediates.
    if (warp == kTcTiles * kWorkerWarps) {
        const uint32_t sbase = smem_addr(base), sring = smem_addr(ring);
#pragma unroll
        for (int s = 0; s < kTcSteps; ++s) {
            uint32_t sa_hi, sa_lo, d_col;
            if (s < 8) {                                    // conv1: A = stage s % 2
                bar_wait(&s_full_a[s % kAStages], (uint32_t)((s / kAStages) & 1));
                sa_hi = sbase + (uint32_t)((s % kAStages) * kTcAStage);
                sa_lo = sa_hi + kTcATile;
                d_col = kColD12;
            } else {
                const int kc = s & 1;                       // steps 8.. are (tap, kc) then (quarter, kc): kc = parity of s
                int shift = 0;
                if (s < 26) {
                    if (s == 8) bar_wait(&s_tready, 0u);    // t1 is in place (and D1 has been read: conv2 reuses its columns)
                    const int tap = (s - 8) >> 1;
                    shift = 4 * (tap / 3 - 1) + (tap % 3 - 1);
                    d_col = kColD12;
                } else {
                    if (s == 26) bar_wait(&s_tready, 1u);
                    if (s == 30) bar_wait(&s_d3free, 0u);   // the first half has left TMEM
                    d_col = kColD3 + 64u * (uint32_t)(((s - 26) >> 1) & 1);
                }
                sa_hi = sbase + (uint32_t)(kc * kPlaneBytes + (kMargin + shift) * 128);
                sa_lo = sa_hi + 2 * kPlaneBytes;
            }
            bar_wait(&s_full_b[s % kTcBRing], (uint32_t)((s / kTcBRing) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const bool first = s == 0 || s == 8 || (s >= 26 && (s & 1) == 0);   // first K chunk of an accumulator
                const uint32_t sb = sring + (uint32_t)((s % kTcBRing) * kTcBSlot);
                const uint64_t a_hi = umma_desc(sa_hi), a_lo = umma_desc(sa_lo);
                const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + kTcBTile);
#pragma unroll
                for (int t = 0; t < kTcTiles; ++t) {
                    const uint64_t ta_hi = a_hi + (uint64_t)(t * (kRegion0 >> 4)), ta_lo = a_lo + (uint64_t)(t * (kRegion0 >> 4));
                    const uint32_t d = tmem + (uint32_t)t * kTileCols + d_col;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {        // K = 32 per step: four K = 8 instructions, 32 bytes apart
                        umma_tf32(d, ta_hi + 2 * ks, b_hi + 2 * ks, (first && ks == 0) ? 0u : 1u);
                        umma_tf32(d, ta_hi + 2 * ks, b_lo + 2 * ks, 1u);
                        umma_tf32(d, ta_lo + 2 * ks, b_hi + 2 * ks, 1u);
                    }
                }
                umma_commit(&s_free_b[s % kTcBRing]);
                if (s < 8) umma_commit(&s_free_a[s % kAStages]);
                if (s == 7) umma_commit(&s_phase[0]);
                if (s == 25) umma_commit(&s_phase[1]);
                if (s == 29) umma_commit(&s_phase[2]);
                if (s == kTcSteps - 1) umma_commit(&s_phase[3]);
            }
            __syncwarp();
        }
        return;
    }

    // ------------------------------ warps 0-17: two groups of nine worker warps, one tile each ------------------------------
    const int grp = warp / kWorkerWarps;
    const int roi0 = (blockIdx.x * kTcTiles + grp) * kTcRois;
    const int nroi = max(0, min(kTcRois, live - roi0));        // 0: the last CTA's second tile may be empty (it still keeps step)
    const int wtid = tid - grp * kWorkers, wwarp = warp - grp * kWorkerWarps;
    base += grp * kRegion0;
    const uint32_t tmem_t = tmem + (uint32_t)grp * kTileCols;
    // conv1 A tile: 72 live rows x eight 16-byte chunks = 576 items, two per thread, the same two for all 8 K chunks
    const float* xptr[kTcItems];
    float* xscr[kTcItems];                                      // fp32 copy of x for the residual (tile-path RoIs)
    int xpc[kTcItems], xoff[kTcItems];
    float xinv[kTcItems];
#pragma unroll
    for (int q = 0; q < kTcItems; ++q) {
        const int i = wtid + q * kWorkers, r72 = i >> 3, ch = i & 7;
        const int rl = r72 / 9, p = r72 - rl * 9, m = kRoiRows * rl + 4 + 4 * (p / 3) + p % 3;
        xoff[q] = m * 128 + ((ch ^ (m & 7)) << 4);
        xptr[q] = src.roi_feat; xscr[q] = nullptr; xpc[q] = 0; xinv[q] = 1.f;
        if (rl < nroi) {
            const int n = roi0 + rl;
            int sb = -1, pcs = 0;
            float cnt = 1.f;
            if (src.partial) { sb = __ldg(src.slot + n); pcs = __ldg(src.pieces + n); cnt = __ldg(src.count + n); }
            if (sb < 0) {                                       // finished feature [256][9] (direct RoIAlign path / plain API)
                xpc[q] = -1;
                xptr[q] = src.roi_feat + (size_t)n * 2304 + p + 36 * ch;
            } else {                                            // partial slots [pieces][9][256], to be summed and scaled
                xptr[q] = src.partial + (size_t)sb * 2304 + p * 256 + 4 * ch;
                xscr[q] = src.scratch + (size_t)n * 2304 + p * 256 + 4 * ch;
                xpc[q] = pcs;
                xinv[q] = 1.0f / cnt;
            }
        }
    }
    // x is prefetched one K chunk ahead into registers, up to six slots per item, every load issued before the
    // first add (a load-add-load-add loop would serialise on the in-order issue); the slots are L2 hits after
    // the first wave thanks to the previous CTA's bulk prefetch, and the other CTA of the SM covers the rest.
    float4 xp[kTcItems][kTcSlotsInReg];
    auto load_x_chunk = [&](int kc) {
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {
            xp[q][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xpc[q] < 0) {
                const float* g = xptr[q] + 288 * kc;
                xp[q][0] = make_float4(__ldg(g), __ldg(g + 9), __ldg(g + 18), __ldg(g + 27));
            } else {
#pragma unroll
                for (int k = 0; k < kTcSlotsInReg; ++k)
#ifdef RR_TC_EXP_NOX
                    if (k < xpc[q] && kc < 1)
#else
                    if (k < xpc[q])
#endif
                        xp[q][k] = __ldg(reinterpret_cast<const float4*>(xptr[q] + (size_t)k * 2304 + 32 * kc));
            }
        }
    };
    auto x_value = [&](int kc, int q) {             // same summation order as roi_combine_kernel: slot 0, 1, 2, ...
        float4 v = xp[q][0];
        if (xpc[q] >= 0) {
#pragma unroll
            for (int k = 1; k < kTcSlotsInReg; ++k)
                if (k < xpc[q]) { v.x += xp[q][k].x; v.y += xp[q][k].y; v.z += xp[q][k].z; v.w += xp[q][k].w; }
            for (int k = kTcSlotsInReg; k < xpc[q]; ++k) {      // RoIs cut into more than six pieces are rare
                const float4 t = __ldg(reinterpret_cast<const float4*>(xptr[q] + (size_t)k * 2304 + 32 * kc));
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            const float inv = xinv[q];
            v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
        }
        return v;
    };
    load_x_chunk(0);
    TC_TRACE(2);
#pragma unroll 1
    for (int kc = 0; kc < 8; ++kc) {
        const int st = kc % kAStages;
        if (kc >= kAStages) bar_wait_warp(&s_free_a[st], (uint32_t)((kc / kAStages - 1) & 1));
        uint8_t* a_hi = base + st * kTcAStage;
        uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {
            const float4 v = x_value(kc, q);
            float4 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<float4*>(a_hi + xoff[q]) = hi;
            *reinterpret_cast<float4*>(a_lo + xoff[q]) = lo;
            if (xscr[q]) *reinterpret_cast<float4*>(xscr[q] + 32 * kc) = v;
        }
        if (kc + 1 < 8) load_x_chunk(kc + 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA
        __syncwarp();
        if (lane == 0) bar_arrive(&s_full_a[st]);
        TC_TRACE(16 + kc);
    }
    TC_TRACE(3);

    // t = relu(D + b) of conv1 / conv2 -> (hi, lo) tf32 planes in the MMA tile layout.  Eight warps cover the
    // 4 TMEM lane quarters x 2 column halves (= K chunks of the next GEMM); the ninth zeroes the plane margins.
    auto epilogue_t = [&](const float* bias, bool mask_pad) {
        if (wwarp < 8) {                                    // a warp reaches the TMEM lane quarter warp % 4 of the CTA
            const int q = warp & 3, h = wwarp >> 2, m = 32 * q + lane;
            const bool keep = !mask_pad || ((m & 15) >= 4 && (m & 3) != 3);     // conv2 reads the pad rows as zeros
            float v[32];
            tmem_ld32(tmem_t + ((uint32_t)(32 * q) << 16) + kColD12 + (uint32_t)(32 * h), v);
            uint8_t* p_hi = base + h * kPlaneBytes + (kMargin + m) * 128;
            uint8_t* p_lo = p_hi + 2 * kPlaneBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 o;
                o.x = keep ? fmaxf(v[4 * c] + bias[32 * h + 4 * c], 0.f) : 0.f;
                o.y = keep ? fmaxf(v[4 * c + 1] + bias[32 * h + 4 * c + 1], 0.f) : 0.f;
                o.z = keep ? fmaxf(v[4 * c + 2] + bias[32 * h + 4 * c + 2], 0.f) : 0.f;
                o.w = keep ? fmaxf(v[4 * c + 3] + bias[32 * h + 4 * c + 3], 0.f) : 0.f;
                float4 hi, lo;
                split4(o, hi, lo);
                const int off = (c ^ (m & 7)) << 4;
                *reinterpret_cast<float4*>(p_hi + off) = hi;
                *reinterpret_cast<float4*>(p_lo + off) = lo;
            }
        } else if (mask_pad) {
            for (int i = lane; i < 4 * 16 * 8; i += 32) {       // 4 planes x (8 + 8) margin rows x 8 chunks
                const int pl = i >> 7, r = (i >> 3) & 15, c = i & 7;
                const int row = r < 8 ? r : kMargin + 128 + (r - 8);
                *reinterpret_cast<float4*>(base + pl * kPlaneBytes + row * 128 + c * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&s_tready);
    };
    bar_wait_warp(&s_phase[0], 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TRACE(4);
    epilogue_t(s_b1, true);
    TC_TRACE(5);
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kWorkers) : "memory");   // every worker's copy of x is visible to the others
    TC_TRACE(6);
    bar_wait_warp(&s_phase[1], 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TRACE(7);
    epilogue_t(s_b2, false);
    TC_TRACE(8);

    // ======== y = D3 + b3 + x, relu (resnet.py:49-50), regressor per row, mean over the 9 rows; two halves of 128 channels ========
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int eq = warp & 3, eg = (wwarp >> 2) & 1, em = 32 * eq + lane;
    const bool ekeep = wwarp < 8 && (em & 15) >= 4 && (em & 3) != 3 && (em >> 4) < nroi;
    const float* res = nullptr;                                  // residual source of this thread's row
    bool res_rows = false;                                       // true: [9][256] copy, false: [256][9] feature
    if (ekeep) {
        const int rl = em >> 4, rr = em & 15, p = 3 * ((rr - 4) >> 2) + (rr & 3), n = roi0 + rl;
        res_rows = s_sb[grp][rl] >= 0;
        res = res_rows ? src.scratch + (size_t)n * 2304 + p * 256 : src.roi_feat + (size_t)n * 2304 + p;
    }
    for (int half = 0; half < 2; ++half) {
        bar_wait_warp(&s_phase[2 + half], 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (half == 0) TC_TRACE(9); else TC_TRACE(10);
        if (wwarp < 8) {
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                const int c0 = 128 * half + 64 * eg + 32 * j;
                float x[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) x[e] = 0.f;
                if (ekeep) {
                    if (res_rows) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 t = ld_global_f4(res + c0 + 4 * e);
                            x[4 * e] = t.x; x[4 * e + 1] = t.y; x[4 * e + 2] = t.z; x[4 * e + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) x[e] = __ldg(res + (c0 + e) * 9);
                    }
                }
                float v[32];
                tmem_ld32(tmem_t + ((uint32_t)(32 * eq) << 16) + kColD3 + (uint32_t)(64 * eg + 32 * j), v);
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float y = fmaxf(v[e] + s_b3[c0 + e] + x[e], 0.f);
                    const float4 w = s_wr[c0 + e];
                    acc.x = fmaf(y, w.x, acc.x); acc.y = fmaf(y, w.y, acc.y); acc.z = fmaf(y, w.z, acc.z); acc.w = fmaf(y, w.w, acc.w);
                }
            }
            if (half == 0) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(&s_d3free);
            }
        }
    }
    if (wwarp < 8) {
        if (!ekeep) acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {                       // the 16 rows of a RoI sit in one half warp
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if ((lane & 15) == 0) s_part[grp][eg][2 * eq + (lane >> 4)] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kWorkers) : "memory");
    if (wtid < nroi) {
        const float4 a = s_part[grp][0][wtid], b = s_part[grp][1][wtid];
        reinterpret_cast<float4*>(reg)[roi0 + wtid] =
            make_float4((a.x + b.x) / 9.0f + __ldg(f + kOffBr), (a.y + b.y) / 9.0f + __ldg(f + kOffBr + 1),
                        (a.z + b.z) / 9.0f + __ldg(f + kOffBr + 2), (a.w + b.w) / 9.0f + __ldg(f + kOffBr + 3));
    }
    TC_TRACE(12);
    asm volatile("bar.sync 3, %0;" ::"n"(kTcTiles * kWorkers) : "memory");      // both tiles are out of TMEM
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

int head_tc_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded, float* reg, cudaStream_t st) {
    int rc = 0;
    static bool attr_set = false;
    if (!attr_set) {
        RR_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem), rc);
        attr_set = true;
    }
    if (((uintptr_t)folded & 15) != 0) return RR_E_BADARG;     // the weight stream is copied in 16-byte units
    if (src.partial && !src.scratch) return RR_E_BADARG;
    const int grid = (n_cap + kTcTiles * kTcRois - 1) / (kTcTiles * kTcRois);
    head_tc_kernel<<<grid, kTcBlock, kTcSmem, st>>>(src, n_rois_dev, n_cap, folded, reg, kTcCtasPerSm * kSMs);
    RR_LAUNCHED(rc);
    return rc;
}

}  // namespace rr

#ifdef RR_HEAD_TC_TRACE
RR_API int rr_debug_head_trace(unsigned long long* host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, rr::g_tc_trace, sizeof(unsigned long long) * (size_t)n_words);
}
#endif
