// Internals shared by the decode kernels (rr_decode.cu) and the fused stage-1 tail (rr_tail.cu).
#pragma once
#include "rr_common.cuh"

namespace rr {

constexpr int kCap = 16384;          // candidate capacity per image (== RR_MAX_TOPK)
constexpr int kSampleThreads = 1024;
constexpr int kSamplesPerThread = 32;
constexpr int kCollectThreads = 256;
constexpr int kStage = 1024;         // per-CTA staging entries in decode_collect_kernel
constexpr int kSelectThreads = 1024;

struct DecodeWs {
    unsigned int* thr_key;            // [B]
    unsigned int* count;              // [B]
    unsigned long long* cand;         // [B][kCap]   (key << 32) | ~flat_index
    unsigned int* maxima;             // [B][1024]   folded sample maxima of the fused stage-1 tail (rr_tail.cu)
    size_t bytes;
};
static DecodeWs carve_decode(void* ws, int B) {
    Carver cv(ws);
    DecodeWs w;
    w.thr_key = cv.take<unsigned int>(B);
    w.count = cv.take<unsigned int>(B);
    w.cand = cv.take<unsigned long long>((size_t)B * kCap);
    w.maxima = cv.take<unsigned int>((size_t)B * 1024);
    w.bytes = cv.off;
    return w;
}

// append one candidate: CTA staging first, straight to the image's global list once the staging is full
__device__ __forceinline__ void push_candidate(unsigned key, unsigned flat, unsigned long long* s_stage,
                                               int* s_n, unsigned int* g_count,
                                               unsigned long long* g_cand) {
    unsigned long long e = ((unsigned long long)key << 32) | (unsigned long long)(~flat);
    int slot = atomicAdd(s_n, 1);
    if (slot < kStage) {
        s_stage[slot] = e;
    } else {                            // staging full (dense hits): go straight to global
        unsigned pos = atomicAdd(g_count, 1u);
        if (pos < (unsigned)kCap) g_cand[pos] = e;
    }
}

}  // namespace rr
