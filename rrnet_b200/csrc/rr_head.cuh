// Shared declarations of the re-regression head kernels (rr_head.cu: fp32 FFMA, rr_head_tc.cu: tcgen05).
// Internal header.
#pragma once
#include "rr_common.cuh"

namespace rr {

// Folded parameter block, fp32 part (floats):
//     [W1 : 256 x 64  (k-major)] [b1 : 64]
//     [W2 : 64 cin x 9 taps x 64 cout] [b2 : 64]
//     [W3 : 64 x 256  (k-major)] [b3 : 256]
//     [Wr : 4 x 256] [br : 4]
constexpr int kOffW1 = 0;
constexpr int kOffB1 = kOffW1 + 256 * 64;
constexpr int kOffW2 = kOffB1 + 64;
constexpr int kOffB2 = kOffW2 + 9 * 64 * 64;
constexpr int kOffW3 = kOffB2 + 64;
constexpr int kOffB3 = kOffW3 + 64 * 256;
constexpr int kOffWr = kOffB3 + 256;
constexpr int kOffBr = kOffWr + 4 * 256;
constexpr int kFoldedF32 = kOffBr + 4;
static_assert(kFoldedF32 % 4 == 0, "the tensor-core image that follows must stay 16-byte aligned");

// ... followed by the tensor-core image: weight tiles of 64 rows x 128 bytes (K-major, 128-byte swizzle, exactly the
// shared-memory image a tcgen05.mma B operand wants), each as a (hi, lo) pair for the three-product split.
// RR_HEAD_F16 (default): fp16 elements, 64 K values per row, 17 steps
//     steps  0.. 3  conv1, K chunk kc            B[o][k] = W1[64kc+k][o]
//     steps  4..12  conv2, tap                   B[o][k] = W2[k][tap][o]
//     steps 13..16  conv3, N quarter q           B[n][k] = W3[k][64q+n]
// else: tf32 elements, 32 K values per row, 34 steps
//     steps  0.. 7  conv1, K chunk kc            B[o][k] = W1[32kc+k][o]
//     steps  8..25  conv2, (tap, K chunk)        B[o][k] = W2[32kc+k][tap][o]
//     steps 26..33  conv3, (N quarter q, chunk)  B[n][k] = W3[32kc+k][64q+n]
#ifndef RR_HEAD_F16
#define RR_HEAD_F16 1
#endif
constexpr int kTcSteps = RR_HEAD_F16 ? 17 : 34;
constexpr int kTcTileFloats = 64 * 32;                 // one 8 KB tile (64 rows x 128 bytes)
constexpr int kTcStepFloats = 2 * kTcTileFloats;       // hi | lo
constexpr int kOffTc = kFoldedF32;
constexpr int kFoldedFloats = kFoldedF32 + kTcSteps * kTcStepFloats;

// Where a RoI's 256x9 input comes from: the materialised RoIAlign output, or (fused eval path) the
// partial slots of the tile-centric RoIAlign, summed in slot order and scaled by 1/count.
struct HeadSrc {
    const float* roi_feat;     // [n_cap,256,3,3]; read when partial == nullptr or slot[n] < 0 (direct-path RoI)
    const float* partial;      // [slot][9][256] or nullptr
    const int* slot;           // [n_cap] first slot, < 0: direct path
    const int* pieces;         // [n_cap] number of slots (0: all-zero output)
    const float* count;        // [n_cap] divisor
    float* scratch;            // [n_cap,9,256] workspace for the tensor-core head's residual copy of x (tile-path RoIs), with partial
    int combined;              // 1: `partial` holds one finished, already scaled [9][256] row per tile-path RoI AT ROW n (the
                               // tile kernel combined the slots itself): one piece per RoI, no scaling, no residual copy;
                               // 2: the same row holds the UNSCALED sum (atomic accumulation): scaled by 1/count, copy written in place
};

int head_ffma_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded, float* reg, cudaStream_t st);
int head_tc_launch_src(HeadSrc src, const int32_t* n_rois_dev, int n_cap, const float* folded, float* reg, cudaStream_t st);
int head_fold_tc_launch(float* folded, cudaStream_t st);

}  // namespace rr
