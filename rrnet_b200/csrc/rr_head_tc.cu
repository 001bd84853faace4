// Re-regression head on the 5th-generation tensor cores (tcgen05, sm_100a), fp32 accuracy through a
// 3xTF32 split.
//
// Same contract as head_forward_kernel (rr_head.cu): Bottleneck(256,64) -> avg-pool -> 1x1 conv 256->4 on
// the RoIs' [256,3,3] features (models/rrnet.py:155-157, detectors/fasterrcnn_detector.py:13-18,
// backbones/resnet.py:33-53), BatchNorms folded.  A tile is 8 RoIs.  Their pixels are the rows of one
// M = 128 MMA tile in a padded list of 16 rows per RoI,
//     row(roi, py, px) = 16 roi + 4 + 4 py + px          (rows 16 roi + 0..3 and every px = 3 row stay zero)
// so that the neighbour (py + dy, px + dx) of a pixel is the row 4 dy + dx further down and every neighbour
// outside the 3x3 map is a zero row.  The three convolutions are GEMMs over that tile:
//     conv1  [128 x 256] x [256 x  64]              8 K-chunks of 32
//     conv2  9 taps x [128 x 64] x [64 x 64]        a tap is the SAME t1 tile read through a shared-memory
//                                                   descriptor whose start address is shifted by 4 dy + dx rows
//                                                   (the 128-byte swizzle is a function of the address, so a
//                                                   128-byte-aligned start is legal: tools/tc_shift_probe.cu)
//     conv3  [128 x  64] x [64 x 256]               4 N-quarters x 2 K-chunks
// Nothing is restaged between the taps: after conv1 the only data movement is the weight stream.
// Every MMA is tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 64, K = 8, both operands in shared memory:
// 48 cycles each, bound by the 6 KB of operands it reads (tools/tc_rate_probe.cu), not by the 32 cycles of
// math.  An fp32 product a*b is evaluated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with hi = tf32(v),
// lo = tf32(v - hi): three MMAs per K-step, error ~2^-21 relative per product, inside the 1e-5 parity budget.
//
// The kernel is persistent (one CTA per SM walks over the tiles) and software-pipelined across tiles, because
// inside one tile the tensor pipe and the ordinary warps only take turns (conv1 -> epilogue -> conv2 ->
// epilogue -> conv3 -> epilogue), and because the 316 MB of RoIAlign partial slots come in 128-byte pieces
// from DRAM (~10 us per tile at about half of the HBM bandwidth).  Twenty warps, five concurrent roles:
//   warps 0-3    E1(k): t1 = relu(conv1 + b1) -> tf32 (hi, lo) tiles in shared memory (conv2's shifted A operand);
//                E2(k): t2 = relu(conv2 + b2) -> tf32 (hi, lo) in TENSOR memory: conv3 takes its A operand there,
//                so the shared-memory tiles are free for t1 of tile k+1 while conv3(k) runs
//   warps 4-7    E3(k): y = conv3 + b3 + x, relu, regressor, mean over the 9 pixels (the long epilogue, off the
//                tensor core's critical path); x comes from an fp32 copy the loaders leave in global memory
//   warps 8-16   x loaders: sum the partial slots of tile k+1.. (slot order = roi_combine_kernel's), scale, split,
//                fill the two conv1 A stages, as far ahead as the stages allow
//   warp  17     issues conv1 of every tile as soon as a stage is published (accumulator double-buffered by
//                tile parity, so conv1(k+1) runs under conv2(k) / conv3(k))
//   warp  18     issues conv2 and conv3 of every tile, two weight steps per round, steps unrolled (immediates)
//   warp  19     lane 0 streams the pre-split, pre-swizzled weight tile pairs (16 KB each, made once by
//                rr_head_fold) with cp.async.bulk (TMA) + complete_tx: one slot for the conv1 stream (paced by
//                the loaders), a ring of four for conv2 / conv3
// In an issuer warp every lane waits on the barriers and one elected lane issues: tcgen05.mma is nearly
// synchronous for the issuing thread (tools/tc_rate_probe.cu; every instruction between two MMAs shows up in
// the rate), so descriptors stay in uniform registers and the MMAs go out back to back.  mbarriers connect the
// roles; there is no block-wide barrier after the prologue.  TMEM (512 columns): conv1 / conv2 accumulator
// 0-63 / 64-127 by tile parity, conv3 accumulator 128-383, t2 (hi | lo) 384-511.
// What sets the pace now (tools/head_trace.py, 17.3 us per tile against 10.3 us of MMAs): two chains of similar
// length.  (1) The conv2/conv3 issuer: ~11.5 us of MMAs (48-54 cycles each with the commits), 3.7 us waiting for
// t1 / t2 (the epilogues between conv1 -> conv2 -> conv3 of one tile), 0.9 us for weights.  (2) The x loaders:
// 15.5 us per tile, 1.7 us per K chunk with one chunk in flight per SM (registers, shared and tensor memory are
// full); without any x load the period is still 15.4 us.  Measured alternatives that lost or changed nothing: L2
// prefetch of the slots (bulk, a tile ahead: -6 %; per line, 1-7 chunks ahead: +-0), two chunks ahead with fewer
// warps left for the epilogues, slot-major x loading, two CTAs per SM with half the resources, two tiles per CTA
// in lockstep, a single epilogue group (E3 then delays t1 of the next tile), conv2(k+1) issued before conv3(k)
// (the tensor core then waits for conv1 of the next tile instead of for t2).
// The regressor is applied per row before the pooling (both are linear): reg = (sum_p Wr.relu(y_p)) / 9 + br.
#include "rr_head.cuh"

#include <cuda_fp16.h>

namespace rr {

constexpr int kTcRois = 8;                       // RoIs per tile
constexpr int kRoiRows = 16;                     // tile rows per RoI: 3 x (3 pixels + 1 zero) + 4 zero rows in front
constexpr int kEpiWarps = 8;                     // epilogue warps: 4 TMEM lane quarters x 2 column halves
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kLoadWarps = 9;
constexpr int kLoaders = 32 * kLoadWarps;        // 288 threads x 2 items = the 576 live (row, 16-byte chunk) items of an x tile
constexpr int kTcItems = 2;
constexpr int kTcSlotsInReg = 6;                 // partial slots of an item held in registers per K chunk
constexpr int kWarpIssuer1 = kEpiWarps + kLoadWarps, kWarpIssuer2 = kWarpIssuer1 + 1, kWarpWeights = kWarpIssuer2 + 1;
constexpr int kTcBlock = 32 * (kWarpWeights + 1);
constexpr int kTcATile = 128 * 128;              // bytes: 128 rows x 32 tf32
constexpr int kTcBTile = 64 * 128;               // bytes:  64 rows x 32 tf32
constexpr int kTcAStage = 2 * kTcATile;          // A_hi | A_lo = 32 KB
constexpr int kAStages = 2;                      // conv1 A stages; ring 1 (conv1 weights) has one slot per stage
constexpr int kTcBSlot = 2 * kTcBTile;           // B_hi | B_lo = 16 KB = one step of the weight image
constexpr int kRing1 = 1;                        // slots of ring 1 (conv1 weights): the conv1 stream is paced by the x loaders, not by its weights
constexpr int kXRing = 8;                        // tiles of loader -> output-epilogue hand-off state (s_xdone, s_sb)
constexpr int kRing2 = 4;                        // slots of ring 2 (conv2 / conv3 weights)
#if RR_HEAD_F16
// fp16 operands (kind::f16): a 128-byte row holds 64 K values instead of 32, so every GEMM takes half as many MMAs for
// the same operand bytes - and the MMA rate here is set by the operand bytes (48 cycles per instruction either way).
// hi = fp16(v), lo = fp16(v - hi): 22 mantissa bits like the tf32 pair; |x|, |t|, |w| must stay below 65 504.
constexpr int kSteps1 = 4, kSteps2 = 13;         // weight steps of stream 1 (conv1: 4 K chunks of 64) and stream 2 (9 taps + 4 quarters)
#else
constexpr int kSteps1 = 8, kSteps2 = 26;         // weight steps of stream 1 (conv1) and stream 2 (conv2, conv3)
#endif
constexpr int kMargin = 8;                       // zero rows above and below the 128 tile rows of a t1 / t2 plane
constexpr int kPlaneBytes = (128 + 2 * kMargin) * 128;
#if RR_HEAD_F16
constexpr int kT2PlaneBytes = 128 * 128;         // t2 (conv3's A operand) lives in shared memory too: no shifted reads, no margins
constexpr int kTBytes = 2 * kPlaneBytes + 2 * kT2PlaneBytes;     // t1 planes (hi | lo) + t2 planes (hi | lo): 68 KB
#else
constexpr int kTBytes = 4 * kPlaneBytes;         // t planes (hi | lo) x (kc 0 | 1): 72 KB
#endif
static_assert(kTcBSlot == kTcStepFloats * 4, "a ring slot is one step of the folded image");
static_assert((kSteps1 / 2) % 2 == 0 && kSteps1 % 2 == 0, "stage / ring-1 phase parities assume an even number of uses per tile");
static_assert(kSteps1 + kSteps2 == kTcSteps, "the two streams cover the folded image");
constexpr int kTcSmem = kTBytes + kAStages * kTcAStage + (kRing1 + kRing2) * kTcBSlot + 1024;   // + slack for the 1024-byte alignment
constexpr int kTmemCols = 512;
constexpr uint32_t kColD12 = 0, kColD3 = 128;    // conv1 / conv2 accumulator (+64 for odd tiles), conv3 accumulator (256 columns)
constexpr uint32_t kColT2 = 384;                 // t2 as conv3's A operand: 64 columns tf32 hi, 64 columns lo
#if RR_HEAD_F16
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
//                          D = f32      A = f16      B = f16       N = 64               M = 128      (both K-major)
#define RR_MMA_KIND "kind::f16"
#else
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
//                          D = f32      A = tf32     B = tf32      N = 64               M = 128      (both K-major)
#define RR_MMA_KIND "kind::tf32"
#endif

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate,
                                          uint32_t idesc = kIdesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1." RR_MMA_KIND " [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand in tensor memory: A[m][k] = lane m, column a_tmem + k (32-bit tf32 words)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1." RR_MMA_KIND " [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                   "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
                   "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
                   "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31]) : "memory");
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
            __nanosleep(32);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
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
// fp16 pair of a value: hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// four values -> 8 bytes of hi and 8 bytes of lo (consecutive K)
__device__ __forceinline__ void split4_h(float4 v, uint2& hi, uint2& lo) {
    __half h[4], l[4];
    split_h(v.x, h[0], l[0]); split_h(v.y, h[1], l[1]); split_h(v.z, h[2], l[2]); split_h(v.w, h[3], l[3]);
    hi.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    hi.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    lo.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    lo.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
}
// ---- weight image: (hi, lo) tiles in the swizzled shared-memory layout, made once per fold ----
#if RR_HEAD_F16
__global__ void head_fold_tc_kernel(float* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // one (step, n, k) element: 64 x 64 per tile
    if (i >= kTcSteps * 64 * 64) return;
    const int step = i / 4096, e = i - step * 4096;
    const int n = e >> 6, k = e & 63;                               // tile row (output channel), K index inside the chunk
    float w;
    if (step < 4) {
        w = f[kOffW1 + (64 * step + k) * 64 + n];
    } else if (step < 13) {
        w = f[kOffW2 + (k * 9 + (step - 4)) * 64 + n];
    } else {
        w = f[kOffW3 + k * 256 + 64 * (step - 13) + n];
    }
    __half hi, lo;
    split_h(w, hi, lo);
    const int pos = n * 64 + ((((k >> 3) ^ (n & 7))) << 3) + (k & 7);          // in halfs: 16-byte chunks of 8 K values, swizzled
    __half* dst = reinterpret_cast<__half*>(f + kOffTc + step * kTcStepFloats);
    dst[pos] = hi;
    dst[2 * kTcTileFloats + pos] = lo;                                          // the lo tile follows 8 KB later
}
constexpr int kFoldThreads = kTcSteps * 64 * 64;
#else
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

constexpr int kFoldThreads = kTcSteps * kTcTileFloats;
#endif

int head_fold_tc_launch(float* folded, cudaStream_t st) {
    int rc = 0;
    head_fold_tc_kernel<<<(kFoldThreads + 255) / 256, 256, 0, st>>>(folded);
    RR_LAUNCHED_K(rc, "head_fold_tc_kernel", st);
    return rc;
}


#ifdef RR_HEAD_TC_TRACE      // tools/head_trace.py: per-CTA time stamps (never defined in the shipped build)
__device__ unsigned long long g_tc_trace[2048 * 32];
#define TC_TRACE(k) do { if (threadIdx.x == 0 && blockIdx.x < 2048) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tc_trace[blockIdx.x * 32 + (k)] = t_; } } while (0)
#define TC_TRACE_VAL(k, v) do { if (blockIdx.x < 2048) g_tc_trace[blockIdx.x * 32 + (k)] = (unsigned long long)(v); } while (0)
__device__ __forceinline__ long long tc_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_LAP(acc) do { const long long n_ = tc_now(); (acc) += n_ - lap_; lap_ = n_; } while (0)
#define TC_TIMED(acc, stmt) do { const long long a_ = tc_now(); stmt; (acc) += tc_now() - a_; } while (0)
#else
#define TC_TIMED(acc, stmt) do { stmt; } while (0)
#define TC_LAP(acc) do { } while (0)
#define TC_TRACE(k) do { } while (0)
#define TC_TRACE_VAL(k, v) do { } while (0)
#endif

__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool bar_test(unsigned long long* bar, uint32_t parity) {      // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ float4 ld_global_f4(const float* p) {      // coherent load: the data was written by this kernel
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kTcBlock, 1)
head_tc_kernel(HeadSrc src, const int* __restrict__ n_rois_dev, int n_cap,
               const float* __restrict__ f, float* __restrict__ reg) {
    RR_PDL_PROLOGUE();
    extern __shared__ uint8_t s_dyn[];
    __shared__ __align__(8) unsigned long long s_full_a[kAStages];   // conv1 A stage filled (one arrival per worker warp)
    __shared__ __align__(8) unsigned long long s_free_a[kAStages];   // ... consumed (tcgen05.commit)
    __shared__ __align__(8) unsigned long long s_full_b1[kRing1];    // ring 1 slot landed (complete_tx)
    __shared__ __align__(8) unsigned long long s_free_b1[kRing1];    // ... consumed (tcgen05.commit)
    __shared__ __align__(8) unsigned long long s_full_b2[kRing2];
    __shared__ __align__(8) unsigned long long s_free_b2[kRing2];
    __shared__ __align__(8) unsigned long long s_phase[4];           // MMAs of conv2 [1] / conv3 first half [2] / second half [3] are done ([0] unused)
    __shared__ __align__(8) unsigned long long s_t1ready, s_t2ready; // t1 / t2 of a tile written by the epilogue (one phase per tile each,
                                                                     // so that the issuer can never miss a phase: see the waits)
    __shared__ __align__(8) unsigned long long s_d3free;             // conv3's accumulator has been read out (one phase per tile)
    __shared__ uint32_t s_tmem;
    // Ring of kXRing tiles for the loaders -> output-epilogue hand-off.  When the loaders start tile j, the dependency chain
    // (stage free <- conv1(j-1) issued <- E2(j-3) done <- conv2(j-3) complete <- conv3(j-4) issued <- E3(j-5) done) only
    // guarantees that the output epilogue of tile j-5 is over, so the ring must hold more than five tiles.
    __shared__ __align__(8) unsigned long long s_xdone[kXRing];      // fp32 copy of x (and s_sb) of a tile written
    __shared__ __align__(8) unsigned long long s_c1done[2];          // MMAs of conv1 of a tile are done, by tile parity (the conv1 issuer runs ahead)
    __shared__ __align__(8) unsigned long long s_d12free[2];         // conv1 / conv2 accumulator of a tile parity read out for good
    __shared__ int s_sb[kXRing][kTcRois];                            // first slot of the tile's RoIs
    __shared__ float s_b1[64], s_b2[64], s_b3[256];
    __shared__ float4 s_wr[256];                                     // regressor weights, one float4 per channel
    TC_TRACE(0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    const int n_tiles = (live + kTcRois - 1) / kTcRois;
    if ((int)blockIdx.x >= n_tiles) return;
    const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;       // tiles blockIdx.x + k * gridDim.x

    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)s_dyn + 1023) & ~(uintptr_t)1023);
    // t planes: plane (hi|lo, kc) at base + (2*lo + kc) * kPlaneBytes: 8 zero rows, the 128 tile rows, 8 zero rows
    uint8_t* stages = base + kTBytes;                           // conv1 A stages (A_hi | A_lo)
    uint8_t* ring1 = stages + kAStages * kTcAStage;             // conv1 weights, slot = stage
    uint8_t* ring2 = ring1 + kRing1 * kTcBSlot;                 // conv2 / conv3 weights

    // ------------------------------ prologue (all warps) ------------------------------
    for (int i = tid; i < (kTBytes + kAStages * kTcAStage) / 16; i += kTcBlock)   // margins and pad rows stay zero for good
        reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 64; i += kTcBlock) { s_b1[i] = __ldg(f + kOffB1 + i); s_b2[i] = __ldg(f + kOffB2 + i); }
    for (int i = tid; i < 256; i += kTcBlock) {
        s_b3[i] = __ldg(f + kOffB3 + i);
        s_wr[i] = make_float4(__ldg(f + kOffWr + i), __ldg(f + kOffWr + 256 + i), __ldg(f + kOffWr + 512 + i),
                              __ldg(f + kOffWr + 768 + i));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        auto init = [](unsigned long long* b, int count) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
        };
        for (int i = 0; i < kAStages; ++i) {
            init(&s_full_a[i], kLoadWarps); init(&s_free_a[i], 1);
            init(&s_d12free[i], 4);
            init(&s_c1done[i], 1);
        }
        for (int i = 0; i < kRing2; ++i) { init(&s_full_b2[i], 1); init(&s_free_b2[i], 1); }
        for (int i = 0; i < kRing1; ++i) { init(&s_full_b1[i], 1); init(&s_free_b1[i], 1); }
        for (int i = 0; i < 4; ++i) init(&s_phase[i], 1);
        init(&s_t1ready, 4);
        init(&s_t2ready, 4);
        init(&s_d3free, 4);
        for (int i = 0; i < kXRing; ++i) init(&s_xdone[i], kLoadWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // zero fill -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TRACE(1);
    const uint32_t tmem = s_tmem;
    const float* ftc = f + kOffTc;

    // ------------------------------ warp 19: the weight streams ------------------------------
    if (warp == kWarpWeights) {
        if (lane == 0) {
            auto load_b = [&](unsigned long long* full, uint8_t* slot, int image_step) {
                const uint32_t bar = smem_addr(full);
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(kTcBSlot) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr(slot)), "l"(ftc + (size_t)image_step * kTcStepFloats), "r"(kTcBSlot), "r"(bar) : "memory");
            };
            const int total1 = n_my * kSteps1, total2 = n_my * kSteps2;
            int c1 = 0, c2 = 0, s2 = 0, slot2 = 0, use2 = 0;           // loads issued; step inside the tile / slot / use count of ring 2
            while (c1 < total1 || c2 < total2) {
                if (c2 < total2 && (use2 == 0 || bar_test(&s_free_b2[slot2], (uint32_t)((use2 - 1) & 1)))) {
                    load_b(&s_full_b2[slot2], ring2 + slot2 * kTcBSlot, kSteps1 + s2);
                    ++c2;
                    if (++s2 == kSteps2) s2 = 0;
                    if (++slot2 == kRing2) { slot2 = 0; ++use2; }
                }
                if (c1 < total1 && (c1 == 0 || bar_test(&s_free_b1[0], (uint32_t)((c1 - 1) & 1)))) {
                    load_b(&s_full_b1[0], ring1, c1 % kSteps1);
                    ++c1;
                }
            }
        }
        return;
    }

    // ------------------------------ warps 17, 18: the MMA issuers ------------------------------
    // Two instruction streams into the one tensor pipe, one warp each: conv1 of every tile (as soon as the loaders
    // publish a stage), and conv2 + conv3 of every tile.  A warp stays converged (every lane waits on the
    // barriers) and one elected lane issues, so the compiler keeps the descriptors in uniform registers and emits
    // the MMAs back to back; tcgen05.mma is nearly synchronous for the issuing warp (tools/tc_rate_probe.cu: every
    // instruction between two MMAs shows up in the rate), hence the unrolled steps with immediate operands.
    if (warp == kWarpIssuer1 || warp == kWarpIssuer2) {
        const uint32_t sT = smem_addr(base), sA = smem_addr(stages), sB1 = smem_addr(ring1), sB2 = smem_addr(ring2);
        auto issue12 = [&](uint32_t sa_hi, uint32_t sa_lo, uint32_t sb, uint32_t d, bool first) {
            const uint64_t a_hi = umma_desc(sa_hi), a_lo = umma_desc(sa_lo);
            const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + kTcBTile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                    // K = 32 per step: four K = 8 instructions, 32 bytes apart
                umma_tf32(d, a_hi + 2 * ks, b_hi + 2 * ks, (first && ks == 0) ? 0u : 1u);
                umma_tf32(d, a_hi + 2 * ks, b_lo + 2 * ks, 1u);
                umma_tf32(d, a_lo + 2 * ks, b_hi + 2 * ks, 1u);
            }
        };
        auto issue12_ts = [&](uint32_t a_hi, uint32_t sb, uint32_t d, bool first) {       // a_lo = a_hi + 64 columns
            const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + kTcBTile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_tf32_ts(d, a_hi + 8u * ks, b_hi + 2 * ks, (first && ks == 0) ? 0u : 1u);
                umma_tf32_ts(d, a_hi + 8u * ks, b_lo + 2 * ks, 1u);
                umma_tf32_ts(d, a_hi + 64u + 8u * ks, b_hi + 2 * ks, 1u);
            }
        };
        if (warp == kWarpIssuer1) {
            long long w_a = 0, w_b1 = 0;                         // trace build: ns blocked on the x stages / on the conv1 weight slot
            for (int j = 0; j < n_my; ++j) {                    // conv1 of tile j -> D12[j & 1]
                if (j >= 2) bar_wait(&s_d12free[j & 1], (uint32_t)(((j >> 1) - 1) & 1));   // tile j - 2 has left that accumulator
                const uint32_t d = tmem + kColD12 + 64u * (uint32_t)(j & 1);
#pragma unroll
                for (int s1 = 0; s1 < kSteps1; ++s1) {          // stage s1 & 1, its use (kSteps1 / 2) j + (s1 >> 1); kSteps1 / 2 is even
                    TC_TIMED(w_a, bar_wait(&s_full_a[s1 & 1], (uint32_t)((s1 >> 1) & 1)));
                    TC_TIMED(w_b1, bar_wait(&s_full_b1[0], (uint32_t)(s1 & 1)));                // ring-1 use kSteps1 j + s1, kSteps1 is even
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t sa = sA + (uint32_t)((s1 & 1) * kTcAStage);
                        issue12(sa, sa + kTcATile, sB1, d, s1 == 0);
                        umma_commit(&s_free_a[s1 & 1]);
                        umma_commit(&s_free_b1[0]);
                        if (s1 == kSteps1 - 1) umma_commit(&s_c1done[j & 1]);
                    }
                }
            }
#ifdef RR_HEAD_TC_TRACE
            if (lane == 0) { TC_TRACE_VAL(25, w_a); TC_TRACE_VAL(26, w_b1); }
#endif
            return;
        }
        uint32_t c2 = 0;                                        // stream-2 steps issued (all tiles): ring slot c2 % kRing2, its use c2 / kRing2
        long long w_b2 = 0, w_t = 0, w_c1 = 0;                   // trace build: ns blocked on ring 2 / on t1, t2
#if RR_HEAD_F16
        const uint32_t sT2 = sT + 2u * (uint32_t)kPlaneBytes;   // t2 planes (hi | lo) behind the t1 planes
        for (int it = 0; it < n_my; ++it) {
            const uint32_t d12_cur = tmem + kColD12 + 64u * (uint32_t)(it & 1);
            uint32_t slot2 = c2 % kRing2, par2 = (c2 / kRing2) & 1u;
#pragma unroll
            for (int s2 = 0; s2 < kSteps2; ++s2) {              // one step = one tap of conv2 (K = 64) or one N quarter of conv3
                if (s2 == 0) TC_TIMED(w_t, bar_wait(&s_t1ready, (uint32_t)(it & 1)));           // t1 in place
                if (s2 == 9) {
                    TC_TIMED(w_t, bar_wait(&s_t2ready, (uint32_t)(it & 1)));                    // t2 in place
                    if (it > 0) bar_wait(&s_d3free, (uint32_t)((it - 1) & 1));  // the previous tile has left conv3's accumulator
                }
                const uint32_t slot_a = slot2, par_a = par2;
                if (++slot2 == kRing2) { slot2 = 0; par2 ^= 1u; }
                TC_TIMED(w_b2, bar_wait(&s_full_b2[slot_a], par_a));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int shift = 0;
                if (s2 < 9) shift = 4 * (s2 / 3 - 1) + (s2 % 3 - 1);
                const uint32_t sa_hi = sT + (uint32_t)((kMargin + shift) * 128);
                if (elect_one()) {
                    if (s2 < 9) {                               // conv2: A = t1 tiles in shared memory, shifted by the tap
                        issue12(sa_hi, sa_hi + kPlaneBytes, sB2 + slot_a * (uint32_t)kTcBSlot, d12_cur, s2 == 0);
                    } else {                                    // conv3: A = t2 tiles in shared memory, one N quarter per step
                        issue12(sT2, sT2 + kT2PlaneBytes, sB2 + slot_a * (uint32_t)kTcBSlot,
                                tmem + kColD3 + 64u * (uint32_t)(s2 - 9), true);
                    }
                    umma_commit(&s_free_b2[slot_a]);
                    if (s2 == 8) umma_commit(&s_phase[1]);
                    if (s2 == 10) umma_commit(&s_phase[2]);
                    if (s2 == 12) umma_commit(&s_phase[3]);
                }
            }
            c2 += kSteps2;
        }
#else
        for (int it = 0; it < n_my; ++it) {
            const uint32_t d12_cur = tmem + kColD12 + 64u * (uint32_t)(it & 1);
            uint32_t slot2 = c2 % kRing2, par2 = (c2 / kRing2) & 1u;
#pragma unroll
            for (int s2 = 0; s2 < kSteps2; s2 += 2) {           // two steps (kc = 0, 1 of a tap / quarter) per round
                if (s2 == 0) TC_TIMED(w_t, bar_wait(&s_t1ready, (uint32_t)(it & 1)));           // t1 in place
                if (s2 == 18) {
                    TC_TIMED(w_t, bar_wait(&s_t2ready, (uint32_t)(it & 1)));                    // t2 in place
                    if (it > 0) bar_wait(&s_d3free, (uint32_t)((it - 1) & 1));  // the previous tile has left conv3's accumulator
                }
                const uint32_t slot_a = slot2, par_a = par2;
                if (++slot2 == kRing2) { slot2 = 0; par2 ^= 1u; }
                const uint32_t slot_b = slot2, par_b = par2;
                if (++slot2 == kRing2) { slot2 = 0; par2 ^= 1u; }
                TC_TIMED(w_b2, bar_wait(&s_full_b2[slot_a], par_a); bar_wait(&s_full_b2[slot_b], par_b));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int shift = 0;
                if (s2 < 18) { const int tap = s2 >> 1; shift = 4 * (tap / 3 - 1) + (tap % 3 - 1); }
                const uint32_t sa_hi = sT + (uint32_t)((kMargin + shift) * 128);
                const uint32_t d = s2 < 18 ? d12_cur : tmem + kColD3 + 64u * (uint32_t)((s2 - 18) >> 1);
                if (elect_one()) {
                    if (s2 < 18) {                              // conv2: A = t1 tiles in shared memory, shifted by the tap
                        issue12(sa_hi, sa_hi + 2 * kPlaneBytes, sB2 + slot_a * (uint32_t)kTcBSlot, d, s2 == 0);
                        umma_commit(&s_free_b2[slot_a]);
                        issue12(sa_hi + kPlaneBytes, sa_hi + 3 * kPlaneBytes, sB2 + slot_b * (uint32_t)kTcBSlot, d, false);
                        umma_commit(&s_free_b2[slot_b]);
                    } else {                                    // conv3: A = t2 in tensor memory
                        issue12_ts(tmem + kColT2, sB2 + slot_a * (uint32_t)kTcBSlot, d, true);
                        umma_commit(&s_free_b2[slot_a]);
                        issue12_ts(tmem + kColT2 + 32u, sB2 + slot_b * (uint32_t)kTcBSlot, d, false);
                        umma_commit(&s_free_b2[slot_b]);
                    }
                    if (s2 == 16) umma_commit(&s_phase[1]);
                    if (s2 == 20) umma_commit(&s_phase[2]);
                    if (s2 == 24) umma_commit(&s_phase[3]);
                }
            }
            c2 += kSteps2;
        }
#endif
#ifdef RR_HEAD_TC_TRACE
        if (lane == 0) { TC_TRACE_VAL(22, w_b2); TC_TRACE_VAL(23, w_t); TC_TRACE_VAL(24, w_c1); }
#endif
        return;
    }

    // ------------------------------ warps 8-16: x loaders ------------------------------
    // x of tile k -> conv1 A stages (tf32 hi | lo) + fp32 copy for the residual, one tile after the other, as far
    // ahead of the tensor core as the two A stages allow.  The 316 MB of partial slots come from DRAM in 128-byte
    // pieces (about half of the HBM bandwidth is reachable that way): ~10 us per tile, which is why this runs in
    // its own warps next to the MMAs and the epilogues instead of in front of them.
    if (warp >= kEpiWarps && warp < kWarpIssuer1) {
        const int ltid = tid - kEpiThreads;
        int it_row[kTcItems], it_p[kTcItems], it_ch[kTcItems], xoff[kTcItems];
#if RR_HEAD_F16
        int it_m7[kTcItems];
#endif
#pragma unroll
        for (int q = 0; q < kTcItems; ++q) {
            const int i = ltid + q * kLoaders, r72 = i >> 3, ch = i & 7;
            const int rl = r72 / 9, p = r72 - rl * 9, m = kRoiRows * rl + 4 + 4 * (p / 3) + p % 3;
            it_row[q] = rl; it_p[q] = p; it_ch[q] = ch;
#if RR_HEAD_F16
            // fp16 stage rows hold 64 K values: the item's four channels of the 32-channel chunk kc are half of the
            // 16-byte chunk 4 (kc & 1) + (ch >> 1) of row m; xoff is the offset for an even kc, odd chunks flip bit 2 of
            // the chunk index before the swizzle (see the store)
            xoff[q] = m * 128;
            it_m7[q] = m & 7;
#else
            xoff[q] = m * 128 + ((ch ^ (m & 7)) << 4);
#endif
        }
        uint32_t cw = 0;                                        // conv1 chunks staged so far (all tiles)
        long long w_fa = 0;                                     // trace build: ns blocked on a free stage
#ifdef RR_HEAD_TC_TRACE
        const long long t_load0 = tc_now();
#endif
        for (int k = 0; k < n_my; ++k) {
            const int roi0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTcRois;
            const int nroi = min(kTcRois, live - roi0);
            const float* xptr[kTcItems];
            float* xscr[kTcItems];
            int xpc[kTcItems];
            float xinv[kTcItems];
#pragma unroll
            for (int q = 0; q < kTcItems; ++q) {
                const int rl = it_row[q], p = it_p[q], ch = it_ch[q];
                xptr[q] = src.roi_feat; xscr[q] = nullptr; xpc[q] = 0; xinv[q] = 1.f;
                if (rl < nroi) {
                    const int n = roi0 + rl;
                    int sb = -1, pcs = 0;
                    float cnt = 1.f;
                    if (src.partial) {
                        sb = __ldg(src.slot + n); pcs = __ldg(src.pieces + n); cnt = __ldg(src.count + n);
                        if (src.combined && sb >= 0) { sb = n; pcs = pcs > 0 ? 1 : 0; if (src.combined == 1) cnt = 1.f; }     // one row per RoI (2: still to be scaled)
                    }
                    if (sb < 0) {                               // finished feature [256][9] (direct RoIAlign path / plain API)
                        xpc[q] = -1;
                        xptr[q] = src.roi_feat + (size_t)n * 2304 + p + 36 * ch;
                    } else {                                    // partial slots [pieces][9][256], to be summed and scaled
                        xptr[q] = src.partial + (size_t)sb * 2304 + p * 256 + 4 * ch;
                        // combined: the row IS the copy - except for a RoI without pieces (all-zero output), whose row nobody wrote
                        xscr[q] = (src.combined == 1 && pcs > 0) ? nullptr : src.scratch + (size_t)n * 2304 + p * 256 + 4 * ch;      // 2: scaled in place
                        xpc[q] = pcs;
                        xinv[q] = 1.0f / cnt;
                    }
                }
            }
            if (ltid < kTcRois) s_sb[k % kXRing][ltid] = (ltid < nroi && src.partial) ? __ldg(src.slot + roi0 + ltid) : -1;
            // one K chunk ahead in registers, up to six slots per item, every load issued before the first add
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
                        for (int j = 0; j < kTcSlotsInReg; ++j)
#ifdef RR_HEAD_PROBE_ONE_SLOT                                      // timing-only experiment (tools/head_trace.py): wrong results
                            if (j < min(xpc[q], 1))
#else
                            if (j < xpc[q])
#endif
                                xp[q][j] = __ldg(reinterpret_cast<const float4*>(xptr[q] + (size_t)j * 2304 + 32 * kc));
                    }
                }
            };
            load_x_chunk(0);
#pragma unroll 1
            for (int kc = 0; kc < 8; ++kc) {
                float4 v[kTcItems];
#pragma unroll
                for (int q = 0; q < kTcItems; ++q) {            // same summation order as roi_combine_kernel: slot 0, 1, 2, ...
                    v[q] = xp[q][0];
                    if (xpc[q] >= 0) {
#pragma unroll
                        for (int j = 1; j < kTcSlotsInReg; ++j)
                            if (j < xpc[q]) { v[q].x += xp[q][j].x; v[q].y += xp[q][j].y; v[q].z += xp[q][j].z; v[q].w += xp[q][j].w; }
                        for (int j = kTcSlotsInReg; j < xpc[q]; ++j) {   // RoIs cut into more than six pieces are rare
                            const float4 t = __ldg(reinterpret_cast<const float4*>(xptr[q] + (size_t)j * 2304 + 32 * kc));
                            v[q].x += t.x; v[q].y += t.y; v[q].z += t.z; v[q].w += t.w;
                        }
                        const float inv = xinv[q];
                        v[q].x = __fmul_rn(v[q].x, inv); v[q].y = __fmul_rn(v[q].y, inv);
                        v[q].z = __fmul_rn(v[q].z, inv); v[q].w = __fmul_rn(v[q].w, inv);
                    }
                }
                if (kc + 1 < 8) load_x_chunk(kc + 1);           // in flight while this chunk is split and stored
#if RR_HEAD_F16
                // a stage holds TWO 32-channel chunks (K = 64): wait for it before the first, publish it after the second
                const uint32_t fill = cw >> 1, st = fill & 1u, use = fill >> 1;
                if ((cw & 1u) == 0 && use > 0) TC_TIMED(w_fa, bar_wait_warp(&s_free_a[st], (use - 1) & 1u));
                uint8_t* a_hi = stages + st * kTcAStage;
                uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
                for (int q = 0; q < kTcItems; ++q) {
                    uint2 hi, lo;
                    split4_h(v[q], hi, lo);
                    const int chunk = (4 * (int)(cw & 1u) + (it_ch[q] >> 1)) ^ it_m7[q];
                    const int o = xoff[q] + (chunk << 4) + ((it_ch[q] & 1) << 3);
                    *reinterpret_cast<uint2*>(a_hi + o) = hi;
                    *reinterpret_cast<uint2*>(a_lo + o) = lo;
#ifndef RR_HEAD_PROBE_NO_SCRATCH
                    if (xscr[q]) *reinterpret_cast<float4*>(xscr[q] + 32 * kc) = v[q];
#endif
                }
                if (cw & 1u) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
                    __syncwarp();
                    if (lane == 0) bar_arrive(&s_full_a[st]);
                }
                ++cw;
#else
                const uint32_t st = cw & 1u, use = cw >> 1;
                if (use > 0) TC_TIMED(w_fa, bar_wait_warp(&s_free_a[st], (use - 1) & 1u));
                uint8_t* a_hi = stages + st * kTcAStage;
                uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
                for (int q = 0; q < kTcItems; ++q) {
                    float4 hi, lo;
                    split4(v[q], hi, lo);
                    *reinterpret_cast<float4*>(a_hi + xoff[q]) = hi;
                    *reinterpret_cast<float4*>(a_lo + xoff[q]) = lo;
                    if (xscr[q]) *reinterpret_cast<float4*>(xscr[q] + 32 * kc) = v[q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA
                __syncwarp();
                if (lane == 0) bar_arrive(&s_full_a[st]);
                ++cw;
#endif
            }
            __threadfence_block();                              // the fp32 copy and s_sb, before the epilogue warps are told
            __syncwarp();
            if (lane == 0) bar_arrive(&s_xdone[k % kXRing]);
        }
#ifdef RR_HEAD_TC_TRACE
        if (ltid == 0) { TC_TRACE_VAL(27, w_fa); TC_TRACE_VAL(28, tc_now() - t_load0); }
#endif
        return;
    }

    // ------------------------------ warps 0-3: t epilogues, warps 4-7: output epilogue ------------------------------
    // A warp reaches the TMEM lane quarter warp % 4, so each group of four covers the 128 rows.  Two groups, because
    // the output epilogue of tile k (the long one) must not sit between the tensor core and t1 of tile k + 1.
    const int eq = warp & 3, em = 32 * eq + lane;
    const bool pixel_row = (em & 15) >= 4 && (em & 3) != 3;
    if (warp < 4) {
        // E1 / E2: t = relu(D + b) of conv1 / conv2 -> (hi, lo) tf32 planes in the MMA tile layout; the two column
        // halves of the accumulator are the two K chunks of the next GEMM
        // E1: t1 = relu(conv1 + b1) -> (hi, lo) tf32 tiles in shared memory (conv2 reads them through shifted
        //     descriptors; pad rows are written as zeros).  The two column halves are the K chunks of conv2.
        // E2: t2 = relu(conv2 + b2) -> (hi, lo) tf32 in tensor memory, conv3's A operand: the shared-memory tiles
        //     are free for t1 of the next tile while conv3 runs.
#if RR_HEAD_F16
        // t = relu(D + b) as fp16 (hi, lo) rows of 64 K values in shared memory: t1 into the margin-padded planes that
        // conv2 reads through shifted descriptors (pad rows written as zeros), t2 into its own planes (conv3's A operand)
        auto epilogue_t = [&](uint32_t d12, const float* bias, bool to_t1) {
            uint8_t* p_hi = to_t1 ? base + (kMargin + em) * 128 : base + 2 * kPlaneBytes + em * 128;
            uint8_t* p_lo = p_hi + (to_t1 ? kPlaneBytes : kT2PlaneBytes);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(d12 + ((uint32_t)(32 * eq) << 16) + (uint32_t)(32 * h), v);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                   // 16-byte chunk 4 h + c = channels 32 h + 8 c .. + 7
                    float4 o0, o1;
                    const float* bb = bias + 32 * h + 8 * c;
                    const bool on = pixel_row || !to_t1;
                    o0.x = on ? fmaxf(v[8 * c] + bb[0], 0.f) : 0.f;     o0.y = on ? fmaxf(v[8 * c + 1] + bb[1], 0.f) : 0.f;
                    o0.z = on ? fmaxf(v[8 * c + 2] + bb[2], 0.f) : 0.f; o0.w = on ? fmaxf(v[8 * c + 3] + bb[3], 0.f) : 0.f;
                    o1.x = on ? fmaxf(v[8 * c + 4] + bb[4], 0.f) : 0.f; o1.y = on ? fmaxf(v[8 * c + 5] + bb[5], 0.f) : 0.f;
                    o1.z = on ? fmaxf(v[8 * c + 6] + bb[6], 0.f) : 0.f; o1.w = on ? fmaxf(v[8 * c + 7] + bb[7], 0.f) : 0.f;
                    uint2 h0, l0, h1, l1;
                    split4_h(o0, h0, l0);
                    split4_h(o1, h1, l1);
                    const int off = ((4 * h + c) ^ (em & 7)) << 4;
                    *reinterpret_cast<uint4*>(p_hi + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                    *reinterpret_cast<uint4*>(p_lo + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(to_t1 ? &s_t1ready : &s_t2ready);
        };
#else
        auto epilogue_t = [&](uint32_t d12, const float* bias, bool to_smem) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(d12 + ((uint32_t)(32 * eq) << 16) + (uint32_t)(32 * h), v);
                if (to_smem) {
                    uint8_t* p_hi = base + h * kPlaneBytes + (kMargin + em) * 128;
                    uint8_t* p_lo = p_hi + 2 * kPlaneBytes;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 o;
                        o.x = pixel_row ? fmaxf(v[4 * c] + bias[32 * h + 4 * c], 0.f) : 0.f;
                        o.y = pixel_row ? fmaxf(v[4 * c + 1] + bias[32 * h + 4 * c + 1], 0.f) : 0.f;
                        o.z = pixel_row ? fmaxf(v[4 * c + 2] + bias[32 * h + 4 * c + 2], 0.f) : 0.f;
                        o.w = pixel_row ? fmaxf(v[4 * c + 3] + bias[32 * h + 4 * c + 3], 0.f) : 0.f;
                        float4 hi, lo;
                        split4(o, hi, lo);
                        const int off = (c ^ (em & 7)) << 4;
                        *reinterpret_cast<float4*>(p_hi + off) = hi;
                        *reinterpret_cast<float4*>(p_lo + off) = lo;
                    }
                } else {
                    float lo[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float o = fmaxf(v[e] + bias[32 * h + e], 0.f);
                        v[e] = to_tf32(o);
                        lo[e] = to_tf32(o - v[e]);
                    }
                    tmem_st32(tmem + ((uint32_t)(32 * eq) << 16) + kColT2 + (uint32_t)(32 * h), v);
                    tmem_st32(tmem + ((uint32_t)(32 * eq) << 16) + kColT2 + 64u + (uint32_t)(32 * h), lo);
                }
            }
            if (to_smem) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(to_smem ? &s_t1ready : &s_t2ready);
        };
#endif
        TC_TRACE(2);
#ifdef RR_HEAD_TC_TRACE
        long long lap_ = tc_now(), w_p0 = 0, w_e1 = 0, w_x = 0, w_p1 = 0, w_e2 = 0, w_e3 = 0;
#endif
        for (int k = 0; k < n_my; ++k) {
            const uint32_t par = (uint32_t)(k & 1);
            const uint32_t d12 = tmem + kColD12 + 64u * par;
            TC_LAP(w_e3);      // (conv2(k-1) has left the t1 tiles: waited for before E2(k-1))
            bar_wait_warp(&s_c1done[k & 1], (uint32_t)((k >> 1) & 1));    // conv1(k) done
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_LAP(w_p0);
            epilogue_t(d12, s_b1, true);
            TC_LAP(w_e1);
            bar_wait_warp(&s_phase[1], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_LAP(w_p1);
            epilogue_t(d12, s_b2, false);
            if (lane == 0) bar_arrive(&s_d12free[k & 1]);        // (after epilogue_t's tcgen05 fence and __syncwarp)
            TC_LAP(w_e2);
        }
#ifdef RR_HEAD_TC_TRACE
        if (tid == 0) { TC_TRACE_VAL(16, w_p0); TC_TRACE_VAL(17, w_e1); TC_TRACE_VAL(18, w_x); TC_TRACE_VAL(19, w_p1); TC_TRACE_VAL(20, w_e2); TC_TRACE_VAL(21, w_e3); }
#endif
    } else {
        // E3: y = D3 + b3 + x, relu (resnet.py:49-50), regressor per row, mean over the 9 rows.  One warp owns its 32
        // rows (two RoIs) over all 256 channels, 16 at a time; the residual of the next 16 is in flight meanwhile.
        for (int k = 0; k < n_my; ++k) {
            const int roi0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTcRois;
            const int nroi = min(kTcRois, live - roi0);
            const uint32_t par = (uint32_t)(k & 1);
            const bool ekeep = pixel_row && (em >> 4) < nroi;
            bar_wait_warp(&s_xdone[k % kXRing], (uint32_t)((k / kXRing) & 1));     // the loaders' copy of x and s_sb are complete
            const float* res = src.roi_feat;                    // residual source of this thread's row
            bool res_rows = false;                              // true: [9][256] copy, false: [256][9] feature
            if (ekeep) {
                const int rl = em >> 4, rr = em & 15, p = 3 * ((rr - 4) >> 2) + (rr & 3), n = roi0 + rl;
                res_rows = s_sb[k % kXRing][rl] >= 0;
                res = res_rows ? src.scratch + (size_t)n * 2304 + p * 256 : src.roi_feat + (size_t)n * 2304 + p;
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float xa[16], xb[16];
            auto load_res = [&](int c0, float (&x)[16]) {
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = 0.f;
                if (ekeep) {
                    if (res_rows) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 t = ld_global_f4(res + c0 + 4 * e);
                            x[4 * e] = t.x; x[4 * e + 1] = t.y; x[4 * e + 2] = t.z; x[4 * e + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) x[e] = __ldg(res + (c0 + e) * 9);
                    }
                }
            };
            auto block = [&](int c0, const float (&x)[16]) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(32 * eq) << 16) + kColD3 + (uint32_t)c0, v);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float y = fmaxf(v[e] + s_b3[c0 + e] + x[e], 0.f);
                    const float4 w = s_wr[c0 + e];
                    acc.x = fmaf(y, w.x, acc.x); acc.y = fmaf(y, w.y, acc.y); acc.z = fmaf(y, w.z, acc.z); acc.w = fmaf(y, w.w, acc.w);
                }
            };
            load_res(0, xa);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                bar_wait_warp(&s_phase[2 + half], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int c0 = 128 * half; c0 < 128 * half + 128; c0 += 32) {
                    load_res(c0 + 16, xb);
                    block(c0, xa);
                    if (c0 + 32 < 256) load_res(c0 + 32, xa);
                    block(c0 + 16, xb);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(&s_d3free);
            if (!ekeep) acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {                   // the 16 rows of a RoI sit in one half warp
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            }
            const int rl = 2 * eq + (lane >> 4);
            if ((lane & 15) == 0 && rl < nroi)
                reinterpret_cast<float4*>(reg)[roi0 + rl] =
                    make_float4(acc.x / 9.0f + __ldg(f + kOffBr), acc.y / 9.0f + __ldg(f + kOffBr + 1),
                                acc.z / 9.0f + __ldg(f + kOffBr + 2), acc.w / 9.0f + __ldg(f + kOffBr + 3));
        }
    }
    TC_TRACE(12);
    TC_TRACE_VAL(13, n_my);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");      // both epilogue groups are out of TMEM
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

int head_tc_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded, float* reg, cudaStream_t st) {
    int rc = 0;
    static OncePerDevice attr_once; int attr_dev;
    if (attr_once.need(&attr_dev)) {
        RR_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem), rc);
        if (rc == 0) attr_once.mark(attr_dev);
    }
    if (((uintptr_t)folded & 15) != 0) return RR_E_BADARG;     // the weight stream is copied in 16-byte units
    if (src.partial && !src.scratch) return RR_E_BADARG;
    const int n_tiles = (n_cap + kTcRois - 1) / kTcRois;
    const int sms = sms_for_persistent();
    launch_pdl(head_tc_kernel, dim3(n_tiles < sms ? n_tiles : sms), dim3(kTcBlock), kTcSmem, st, src, n_rois_dev, n_cap, folded, reg);
    RR_LAUNCHED_K(rc, "head_tc_kernel", st);
    return rc;
}

}  // namespace rr

#ifdef RR_HEAD_TC_TRACE
RR_API int rr_debug_head_trace(unsigned long long* host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, rr::g_tc_trace, sizeof(unsigned long long) * (size_t)n_words);
}
#endif
