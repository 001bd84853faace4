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
// Folded block layout (floats):
//     [W1 : 256 x 64  (k-major)] [b1 : 64]
//     [W2 : 9 taps x 64 cin x 64 cout] [b2 : 64]
//     [W3 : 64 x 256  (k-major)] [b3 : 256]
//     [Wr : 4 x 256] [br : 4]
// Kernel: one warp per RoI.  The RoI's 256x9 feature block is staged in shared memory (pixel
// stride 12 so a channel's 9 pixels are three 128-bit broadcast loads); lane = output channel
// modulo 32, so weight loads are coalesced 128-byte rows served by L1/L2 (the 284 KB of folded
// weights are shared by every warp on the chip).  The 3x3 convolution on the zero-padded 3x3
// map only issues the 49 valid (tap, pixel) products.
#include "rr_common.cuh"

namespace rr {

constexpr int kOffW1 = 0;
constexpr int kOffB1 = kOffW1 + 256 * 64;
constexpr int kOffW2 = kOffB1 + 64;
constexpr int kOffB2 = kOffW2 + 9 * 64 * 64;
constexpr int kOffW3 = kOffB2 + 64;
constexpr int kOffB3 = kOffW3 + 64 * 256;
constexpr int kOffWr = kOffB3 + 256;
constexpr int kOffBr = kOffWr + 4 * 256;
constexpr int kFoldedFloats = kOffBr + 4;

constexpr int kHeadWarps = 4;
constexpr int kPix = 9, kPixPad = 12;

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
    if (i < 9 * 64 * 64) {                    // W2 [o][c][ky][kx] -> [tap][c][o]
        int tap = i / 4096, c = (i / 64) % 64, o = i % 64;
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

__global__ void __launch_bounds__(kHeadWarps * 32)
head_forward_kernel(const float* __restrict__ x, const int* __restrict__ n_rois_dev, int n_cap,
                    const float* __restrict__ f, float* __restrict__ reg) {
    extern __shared__ float4 s_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * kHeadWarps + warp;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    if (n >= live) return;                    // warps are independent (only __syncwarp below)
    float* s_x = reinterpret_cast<float*>(s_raw) + warp * (256 + 64 + 64) * kPixPad;
    float* s_t1 = s_x + 256 * kPixPad;
    float* s_t2 = s_t1 + 64 * kPixPad;
    const float* xn = x + (size_t)n * 256 * kPix;
    for (int i = lane; i < 256 * kPix; i += 32) {
        int c = i / kPix, p = i - c * kPix;
        s_x[c * kPixPad + p] = __ldg(xn + i);
    }
    __syncwarp();

    // ---- conv1 1x1 256->64 + bn1 + relu: lane owns o = lane, lane+32 ----
    {
        float a0[kPix], a1[kPix];
        const float b0 = __ldg(f + kOffB1 + lane), b1 = __ldg(f + kOffB1 + lane + 32);
#pragma unroll
        for (int p = 0; p < kPix; ++p) { a0[p] = b0; a1[p] = b1; }
        const float* w = f + kOffW1 + lane;
#pragma unroll 4
        for (int c = 0; c < 256; ++c) {
            const float w0 = __ldg(w + c * 64), w1v = __ldg(w + c * 64 + 32);
            const float4* xp = reinterpret_cast<const float4*>(s_x + c * kPixPad);
            const float4 q0 = xp[0], q1 = xp[1], q2 = xp[2];
            const float xv[kPix] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
#pragma unroll
            for (int p = 0; p < kPix; ++p) { a0[p] = fmaf(w0, xv[p], a0[p]); a1[p] = fmaf(w1v, xv[p], a1[p]); }
        }
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
            s_t1[lane * kPixPad + p] = fmaxf(a0[p], 0.f);
            s_t1[(lane + 32) * kPixPad + p] = fmaxf(a1[p], 0.f);
        }
    }
    __syncwarp();

    // ---- conv2 3x3 pad 1 on the 3x3 map, 64->64 + bn2 + relu ----
    {
        float a0[kPix], a1[kPix];
        const float b0 = __ldg(f + kOffB2 + lane), b1 = __ldg(f + kOffB2 + lane + 32);
#pragma unroll
        for (int p = 0; p < kPix; ++p) { a0[p] = b0; a1[p] = b1; }
#pragma unroll 2
        for (int c = 0; c < 64; ++c) {
            const float4* tp = reinterpret_cast<const float4*>(s_t1 + c * kPixPad);
            const float4 q0 = tp[0], q1 = tp[1], q2 = tp[2];
            const float tv[kPix] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* w = f + kOffW2 + ((ky * 3 + kx) * 64 + c) * 64 + lane;
                    const float w0 = __ldg(w), w1v = __ldg(w + 32);
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
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
            s_t2[lane * kPixPad + p] = fmaxf(a0[p], 0.f);
            s_t2[(lane + 32) * kPixPad + p] = fmaxf(a1[p], 0.f);
        }
    }
    __syncwarp();

    // ---- conv3 1x1 64->256 + bn3 + residual + relu + avg-pool, then regressor 256->4 ----
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {             // 4 output channels per lane per half
        float a[4][kPix];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float bj = __ldg(f + kOffB3 + lane + 32 * (half * 4 + j));
#pragma unroll
            for (int p = 0; p < kPix; ++p) a[j][p] = bj;
        }
#pragma unroll 2
        for (int c = 0; c < 64; ++c) {
            const float4* tp = reinterpret_cast<const float4*>(s_t2 + c * kPixPad);
            const float4 q0 = tp[0], q1 = tp[1], q2 = tp[2];
            const float tv[kPix] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
            const float* w = f + kOffW3 + c * 256 + lane + 128 * half;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wj = __ldg(w + 32 * j);
#pragma unroll
                for (int p = 0; p < kPix; ++p) a[j][p] = fmaf(wj, tv[p], a[j][p]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = lane + 32 * (half * 4 + j);
            float pooled = 0.f;
#pragma unroll
            for (int p = 0; p < kPix; ++p) pooled += fmaxf(a[j][p] + s_x[o * kPixPad + p], 0.f);   // resnet.py:49-50
            pooled = pooled / 9.0f;                                                                // avg-pool
            r0 = fmaf(__ldg(f + kOffWr + o), pooled, r0);
            r1 = fmaf(__ldg(f + kOffWr + 256 + o), pooled, r1);
            r2 = fmaf(__ldg(f + kOffWr + 512 + o), pooled, r2);
            r3 = fmaf(__ldg(f + kOffWr + 768 + o), pooled, r3);
        }
    }
    r0 = warp_sum(r0); r1 = warp_sum(r1); r2 = warp_sum(r2); r3 = warp_sum(r3);
    if (lane == 0) {
        float4 o4 = make_float4(r0 + __ldg(f + kOffBr), r1 + __ldg(f + kOffBr + 1),
                                r2 + __ldg(f + kOffBr + 2), r3 + __ldg(f + kOffBr + 3));
        reinterpret_cast<float4*>(reg)[n] = o4;
    }
}

int head_forward_launch(const float* roi_feat, const int32_t* n_rois_dev, int n_cap, const float* folded,
                        float* reg, cudaStream_t st) {
    int rc = 0;
    const size_t smem = (size_t)kHeadWarps * (256 + 64 + 64) * kPixPad * sizeof(float);   // 72 KB
    RR_CUDA(cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), rc);
    const int grid = (n_cap + kHeadWarps - 1) / kHeadWarps;
    head_forward_kernel<<<grid, kHeadWarps * 32, smem, st>>>(roi_feat, n_rois_dev, n_cap, folded, reg);
    RR_LAUNCHED(rc);
    return rc;
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
    RR_LAUNCHED(rc);
    return rc;
}

RR_API int rr_head_forward(const float* roi_feat, const int32_t* n_rois_dev, int n_cap,
                           const float* folded, float* reg, void* stream) {
    if (n_cap == 0) return 0;
    if (!roi_feat || !folded || !reg || n_cap < 0) return RR_E_BADARG;
    if (((uintptr_t)reg & 15) || ((uintptr_t)folded & 15)) return RR_E_ALIGN;
    return head_forward_launch(roi_feat, n_rois_dev, n_cap, folded, reg, (cudaStream_t)stream);
}
