// Re-regression head (eval mode) for sm_100a: Bottleneck(256,64) -> avg-pool -> 1x1 conv 256->4
// on the [N,256,3,3] RoI features, one launch, fp32 FFMA.
//
// Replaces RRNet.forward_stage2 -> FasterRCNNDetector.forward -> Bottleneck.forward
// (models/rrnet.py:155-157, detectors/fasterrcnn_detector.py:13-18, backbones/resnet.py:33-53),
// i.e. 3 cuDNN convolutions + 3 batch-norms + 3 ReLUs + residual add + pooling + 1 more
// convolution (11+ launches on tiny tensors) in the reference.
//
// rr_head_fold folds the eval-mode BatchNorms into the convolutions once:
//     bn(z) = (z - mean) / sqrt(var + 1e-5) * gamma + beta  ->  W' = W * s, b' = beta - mean * s.
// Folded block layout (floats; rr_head.cuh), followed by the tensor-core weight image of rr_head_tc.cu:
//     [W1 : 256 x 64  (k-major)] [b1 : 64]
//     [W2 : 64 cin x 9 taps x 64 cout] [b2 : 64]
//     [W3 : 64 x 256  (k-major)] [b3 : 256]
//     [Wr : 4 x 256] [br : 4]
// Kernel: see head_forward_kernel.  Weight traffic is what bounds a warp-per-RoI formulation (284 KB of
// folded weights per RoI from L2 = 3.2 GB per config-2 step); staging each weight chunk once per CTA in
// shared memory for 16 RoIs cuts it 16x.  The 3x3 convolution on the zero-padded 3x3 map only issues
// the 49 valid (tap, pixel) products.
#include "rr_head.cuh"

namespace rr {

constexpr int kHeadWarps = 16;                   // RoIs per CTA (one warp each); weights are shared by all of them
constexpr int kHeadThreads = kHeadWarps * 32;
constexpr int kPix = 9, kPixPad = 12;
constexpr int kWChunk = 2304;                     // floats per staged weight chunk (9216 B)
constexpr int kC1 = 32, kC2 = 4, kC3 = 8;         // input channels per chunk: conv1 32x64, conv2 4x9x64, conv3 8x256
constexpr int kChunks1 = 256 / kC1, kChunks2 = 64 / kC2, kChunks3 = 64 / kC3;
constexpr int kWarpFloats = 2 * kC1 * kPixPad + 2 * 64 * kPixPad;       // x chunk double buffer | t1 | t2
constexpr int kHeadSmem = (2 * kWChunk + kHeadWarps * kWarpFloats) * (int)sizeof(float);

__global__ void head_fold_kernel(const float* __restrict__ w1, const float* __restrict__ bn1,
                                 const float* __restrict__ w2, const float* __restrict__ bn2,
                                 const float* __restrict__ w3, const float* __restrict__ bn3,
                                 const float* __restrict__ wr, const float* __restrict__ br,
                                 float* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto scale = [](const float* bn, int ch, int o) { return bn[o] / sqrtf(bn[3 * ch + o] + 1e-5f); };
    if (i < 256 * 64) {                       // W1 [o=64][c=256] -> [c][o]
        int c = i / 64, o = i % 64;
        f[kOffW1 + i] = w1[o * 256 + c] * scale(bn1, 64, o);
    }
    if (i < 9 * 64 * 64) {                    // W2 [o][c][ky][kx] -> [c][tap][o]  (a chunk of input channels is contiguous)
        int c = i / 576, tap = (i / 64) % 9, o = i % 64;
        f[kOffW2 + i] = w2[(o * 64 + c) * 9 + tap] * scale(bn2, 64, o);
    }
    if (i < 64 * 256) {                       // W3 [o=256][c=64] -> [c][o]
        int c = i / 256, o = i % 256;
        f[kOffW3 + i] = w3[o * 64 + c] * scale(bn3, 256, o);
    }
    if (i < 64) {
        f[kOffB1 + i] = bn1[64 + i] - bn1[128 + i] * scale(bn1, 64, i);
        f[kOffB2 + i] = bn2[64 + i] - bn2[128 + i] * scale(bn2, 64, i);
    }
    if (i < 256) f[kOffB3 + i] = bn3[256 + i] - bn3[512 + i] * scale(bn3, 256, i);
    if (i < 1024) f[kOffWr + i] = wr[i];
    if (i < 4) f[kOffBr + i] = br[i];
}

// x[c][0..8] of RoI n for channel c = 32*j + lane  (9 values per lane)
__device__ __forceinline__ void head_load_x(const HeadSrc& src, int n, int sb, int pieces, float cnt, int j, int lane,
                                            float (&v)[kPix]) {
    const int c = 32 * j + lane;
    if (sb < 0) {
        const float* xn = src.roi_feat + (size_t)n * 256 * kPix + c * kPix;
#pragma unroll
        for (int p = 0; p < kPix; ++p) v[p] = __ldg(xn + p);
    } else {
#pragma unroll
        for (int p = 0; p < kPix; ++p) v[p] = 0.f;
        for (int k = 0; k < pieces; ++k) {
            const float* pp = src.partial + (size_t)(sb + k) * kPix * 256 + c;
#pragma unroll
            for (int p = 0; p < kPix; ++p) v[p] += __ldg(pp + p * 256);
        }
        const float inv = 1.0f / cnt;               // same rounding as roi_combine_kernel: sum * (1/count)
#pragma unroll
        for (int p = 0; p < kPix; ++p) v[p] = __fmul_rn(v[p], inv);   // never contracted into the residual add
    }
}

__device__ __forceinline__ void load9(const float* s, float (&x)[kPix]) {
    const float4* q = reinterpret_cast<const float4*>(s);
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w; x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w; x[8] = q2.x;
}

__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ src, int floats, int tid) {
    for (int i = tid * 4; i < floats; i += kHeadThreads * 4) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One warp per RoI, 16 RoIs per CTA marching in lock step through 32 weight chunks that are staged once
// per CTA in shared memory (cp.async, double buffered): 8 chunks of conv1 (32 input channels x 64),
// 16 of conv2 (4 x 9 taps x 64), 8 of conv3 (8 x 256).  lane = output channel (2 per lane for conv1/2,
// 8 per lane for conv3: 72 accumulators).  x is streamed through a per-warp double buffer during conv1
// and re-read for the residual; t1/t2 stay in shared memory with pixel stride 12 (three 128-bit
// broadcast loads per channel).
__global__ void __launch_bounds__(kHeadThreads, 1)
head_forward_kernel(HeadSrc src, const int* __restrict__ n_rois_dev, int n_cap,
                    const float* __restrict__ f, float* __restrict__ reg) {
    RR_PDL_PROLOGUE();
    extern __shared__ float4 s_raw[];
    float* s_w = reinterpret_cast<float*>(s_raw);                       // [2][kWChunk]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* s_x = s_w + 2 * kWChunk + warp * kWarpFloats;                // [2][32][12]
    float* s_t1 = s_x + 2 * kC1 * kPixPad;                              // [64][12]
    float* s_t2 = s_t1 + 64 * kPixPad;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    if (blockIdx.x * kHeadWarps >= live) return;                        // whole CTA idle
    const int n = blockIdx.x * kHeadWarps + warp;
    const bool on = n < live;
    int sb = -1, pieces = 0;
    float cnt = 1.f;
    if (on && src.partial) {
        sb = src.slot[n]; pieces = src.pieces[n]; cnt = src.count[n];
        if (src.combined && sb >= 0) { sb = n; pieces = pieces > 0 ? 1 : 0; if (src.combined == 1) cnt = 1.f; }       // one row per RoI (2: still to be scaled)
    }

    stage_weights(s_w, f + kOffW1, kC1 * 64, tid);                      // chunk 0
    float xv[kPix];
    if (on) {
        head_load_x(src, n, sb, pieces, cnt, 0, lane, xv);
#pragma unroll
        for (int p = 0; p < kPix; ++p) s_x[lane * kPixPad + p] = xv[p];
    }

    // ------------------------------ conv1 1x1 256->64 + bn1 + relu ------------------------------
    float a0[kPix], a1[kPix];
    {
        const float b0 = __ldg(f + kOffB1 + lane), b1 = __ldg(f + kOffB1 + lane + 32);
#pragma unroll
        for (int p = 0; p < kPix; ++p) { a0[p] = b0; a1[p] = b1; }
    }
    for (int ch = 0; ch < kChunks1; ++ch) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                                // chunk ch landed; everyone left chunk ch-1
        if (ch + 1 < kChunks1) stage_weights(s_w + ((ch + 1) & 1) * kWChunk, f + kOffW1 + (ch + 1) * kC1 * 64, kC1 * 64, tid);
        else stage_weights(s_w + ((ch + 1) & 1) * kWChunk, f + kOffW2, kC2 * 576, tid);
        if (on) {
            if (ch + 1 < kChunks1) head_load_x(src, n, sb, pieces, cnt, ch + 1, lane, xv);      // prefetch next x chunk
            const float* w = s_w + (ch & 1) * kWChunk + lane;
            const float* xs = s_x + (ch & 1) * kC1 * kPixPad;
#pragma unroll 4
            for (int c = 0; c < kC1; ++c) {
                const float w0 = w[c * 64], w1v = w[c * 64 + 32];
                float x[kPix];
                load9(xs + c * kPixPad, x);
#pragma unroll
                for (int p = 0; p < kPix; ++p) { a0[p] = fmaf(w0, x[p], a0[p]); a1[p] = fmaf(w1v, x[p], a1[p]); }
            }
            if (ch + 1 < kChunks1) {
                float* xd = s_x + ((ch + 1) & 1) * kC1 * kPixPad + lane * kPixPad;
#pragma unroll
                for (int p = 0; p < kPix; ++p) xd[p] = xv[p];
            }
            __syncwarp();
        }
    }
    if (on) {
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
            s_t1[lane * kPixPad + p] = fmaxf(a0[p], 0.f);
            s_t1[(lane + 32) * kPixPad + p] = fmaxf(a1[p], 0.f);
        }
        __syncwarp();
    }

    // ------------------------------ conv2 3x3 pad 1 on the 3x3 map, 64->64 + bn2 + relu ------------------------------
    {
        const float b0 = __ldg(f + kOffB2 + lane), b1 = __ldg(f + kOffB2 + lane + 32);
#pragma unroll
        for (int p = 0; p < kPix; ++p) { a0[p] = b0; a1[p] = b1; }
    }
    for (int ch = 0; ch < kChunks2; ++ch) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int nxt = kChunks1 + ch + 1;
        if (ch + 1 < kChunks2) stage_weights(s_w + (nxt & 1) * kWChunk, f + kOffW2 + (ch + 1) * kC2 * 576, kC2 * 576, tid);
        else stage_weights(s_w + (nxt & 1) * kWChunk, f + kOffW3, kC3 * 256, tid);
        if (on) {
            const float* wb = s_w + ((kChunks1 + ch) & 1) * kWChunk + lane;
#pragma unroll 2
            for (int cc = 0; cc < kC2; ++cc) {
                float tv[kPix];
                load9(s_t1 + (ch * kC2 + cc) * kPixPad, tv);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* w = wb + (cc * 9 + ky * 3 + kx) * 64;
                        const float w0 = w[0], w1v = w[32];
#pragma unroll
                        for (int py = 0; py < 3; ++py) {
                            const int yy = py + ky - 1;
                            if (yy < 0 || yy > 2) continue;
#pragma unroll
                            for (int px = 0; px < 3; ++px) {
                                const int xx = px + kx - 1;
                                if (xx < 0 || xx > 2) continue;
                                a0[py * 3 + px] = fmaf(w0, tv[yy * 3 + xx], a0[py * 3 + px]);
                                a1[py * 3 + px] = fmaf(w1v, tv[yy * 3 + xx], a1[py * 3 + px]);
                            }
                        }
                    }
            }
        }
    }
    if (on) {
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
            s_t2[lane * kPixPad + p] = fmaxf(a0[p], 0.f);
            s_t2[(lane + 32) * kPixPad + p] = fmaxf(a1[p], 0.f);
        }
        __syncwarp();
    }

    // ------------------------------ conv3 1x1 64->256 + bn3 (8 output channels per lane) ------------------------------
    float a[8][kPix];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float bj = __ldg(f + kOffB3 + lane + 32 * j);
#pragma unroll
        for (int p = 0; p < kPix; ++p) a[j][p] = bj;
    }
    for (int ch = 0; ch < kChunks3; ++ch) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int cur = kChunks1 + kChunks2 + ch;
        if (ch + 1 < kChunks3) stage_weights(s_w + ((cur + 1) & 1) * kWChunk, f + kOffW3 + (ch + 1) * kC3 * 256, kC3 * 256, tid);
        if (on) {
            const float* wb = s_w + (cur & 1) * kWChunk + lane;
#pragma unroll 2
            for (int cc = 0; cc < kC3; ++cc) {
                float tv[kPix];
                load9(s_t2 + (ch * kC3 + cc) * kPixPad, tv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float wj = wb[cc * 256 + 32 * j];
#pragma unroll
                    for (int p = 0; p < kPix; ++p) a[j][p] = fmaf(wj, tv[p], a[j][p]);
                }
            }
        }
    }
    if (!on) return;

    // ------------------------------ + residual, relu, avg-pool, regressor 256->4 ------------------------------
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        head_load_x(src, n, sb, pieces, cnt, j, lane, xv);
        const int o = lane + 32 * j;
        float pooled = 0.f;
#pragma unroll
        for (int p = 0; p < kPix; ++p) pooled += fmaxf(a[j][p] + xv[p], 0.f);                       // resnet.py:49-50
        pooled = pooled / 9.0f;                                                                    // avg-pool
        r0 = fmaf(__ldg(f + kOffWr + o), pooled, r0);
        r1 = fmaf(__ldg(f + kOffWr + 256 + o), pooled, r1);
        r2 = fmaf(__ldg(f + kOffWr + 512 + o), pooled, r2);
        r3 = fmaf(__ldg(f + kOffWr + 768 + o), pooled, r3);
    }
    r0 = warp_sum(r0); r1 = warp_sum(r1); r2 = warp_sum(r2); r3 = warp_sum(r3);
    if (lane == 0) {
        float4 o4 = make_float4(r0 + __ldg(f + kOffBr), r1 + __ldg(f + kOffBr + 1),
                                r2 + __ldg(f + kOffBr + 2), r3 + __ldg(f + kOffBr + 3));
        reinterpret_cast<float4*>(reg)[n] = o4;
    }
}

int head_ffma_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded,
                         float* reg, cudaStream_t st) {
    int rc = 0;
    static OncePerDevice attr_once; int attr_dev;
    if (attr_once.need(&attr_dev)) {
        RR_CUDA(cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem), rc);
        if (rc == 0) attr_once.mark(attr_dev);
    }
    const int grid = (n_cap + kHeadWarps - 1) / kHeadWarps;
    launch_pdl(head_forward_kernel, dim3(grid), dim3(kHeadThreads), kHeadSmem, st, src, n_rois_dev, n_cap, folded, reg);
    RR_LAUNCHED_K(rc, "head_forward_kernel", st);
    return rc;
}

// algo: 0 = tcgen05 tensor cores (3xTF32), 1 = fp32 FFMA
int head_forward_launch(const float* roi_feat, const int32_t* n_rois_dev, int n_cap, const float* folded,
                        float* reg, int algo, cudaStream_t st) {
    HeadSrc src = {roi_feat, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    return algo == 1 ? head_ffma_launch_src(src, n_rois_dev, n_cap, folded, reg, st)
                     : head_tc_launch_src(src, n_rois_dev, n_cap, folded, reg, st);
}

// fused eval path: inputs straight from the tile-centric RoIAlign's partial slots
// roi_feat is the eval workspace's own feature buffer here: the rows of tile-path RoIs are unused by RoIAlign
// (combine == 0), so the tensor-core head keeps its residual copy of x there
int head_forward_launch_partial(float* roi_feat, const float* partial, const int* slot, const int* pieces,
                                const float* count, const int32_t* n_rois_dev, int n_cap, const float* folded,
                                float* reg, int algo, int combined, cudaStream_t st) {
    HeadSrc src = {roi_feat, combined ? roi_feat : partial, slot, pieces, count, roi_feat, combined};
    return algo == 1 ? head_ffma_launch_src(src, n_rois_dev, n_cap, folded, reg, st)
                     : head_tc_launch_src(src, n_rois_dev, n_cap, folded, reg, st);
}

}  // namespace rr

using namespace rr;

RR_API size_t rr_head_folded_floats(void) { return (size_t)kFoldedFloats; }

RR_API int rr_head_fold(const float* w1, const float* bn1, const float* w2, const float* bn2,
                        const float* w3, const float* bn3, const float* wr, const float* br,
                        float* folded, void* stream) {
    if (!w1 || !bn1 || !w2 || !bn2 || !w3 || !bn3 || !wr || !br || !folded) return RR_E_BADARG;
    int rc = 0;
    head_fold_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, bn1, w2, bn2, w3, bn3, wr, br, folded);
    RR_LAUNCHED_K(rc, "head_fold_kernel", (cudaStream_t)stream);
    const int r2 = head_fold_tc_launch(folded, (cudaStream_t)stream);     // (hi, lo) tf32 weight tiles for the tensor-core kernel
    return rc ? rc : r2;
}

RR_API int rr_head_forward(const float* roi_feat, const int32_t* n_rois_dev, int n_cap,
                           const float* folded, int algo, float* reg, void* stream) {
    if (n_cap == 0) return 0;
    if (!roi_feat || !folded || !reg || n_cap < 0 || algo < 0 || algo > 1) return RR_E_BADARG;
    if (((uintptr_t)reg & 15) || ((uintptr_t)folded & 15)) return RR_E_ALIGN;
    return head_forward_launch(roi_feat, n_rois_dev, n_cap, folded, reg, algo, (cudaStream_t)stream);
}
