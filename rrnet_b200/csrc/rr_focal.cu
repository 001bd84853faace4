// Penalty-reduced focal loss on heat-map logits for sm_100a, forward and backward.
//
// Replaces focal_loss_for_hm (modules/loss/functional.py:25-51) together with the caller's
// clamp(sigmoid(logits), 1e-4, 1-1e-4) (operators/rrnet_operator.py:55): about twelve
// element-wise kernels, three reductions and ten 21 MB temporaries in the reference (and twice
// that again in autograd's backward) become
//   forward  : one streaming pass over logits+gt (2 x 4 B/element), deterministic two-level sum
//   backward : one pass reading logits+gt and writing the gradient (3 x 4 B/element)
//   fwd_bwd  : both in ONE cooperative launch; the second phase re-reads logits/gt from L2
//              (config 3: 42 MB < 126 MB L2), so HBM sees 3 x 4 B/element in total.
// Sums: fp32 per-thread partials over a few elements, then double precision across the block
// and across blocks in a fixed order (last-block-done), so the loss is bit-reproducible.
#include "rr_gauss.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace rr {

constexpr int kFocalThreads = 256;
constexpr float kLo = 1e-4f, kHi = 1.0f - 1e-4f;

struct FocalWs {
    double* partial;        // [grid][3]
    unsigned int* ticket;   // [1]
    int* num_pos;           // [1]  centre count of the fused render + focal forward+backward
    size_t bytes;
};
static int focal_grid(long long n) {
    long long want = (n / 4 + kFocalThreads * 4 - 1) / (kFocalThreads * 4);   // >= 4 float4 per thread
    long long cap = (long long)kSMs * 8;
    return (int)max(1LL, min(want, cap));
}
static FocalWs carve_focal(void* ws, size_t max_grid = (size_t)kSMs * 8) {
    Carver cv(ws);
    FocalWs w;
    w.partial = cv.take<double>(max_grid * 3);
    w.ticket = cv.take<unsigned int>(1);
    w.num_pos = cv.take<int>(1);
    w.bytes = cv.off;
    return w;
}

// Transcendentals: one ex2, one rcp and one lg2 per element per pass (MUFU approximations, each good to ~2^-22
// relative / absolute in the ranges that matter here; the loss and the gradients stay well inside the 1e-5 bar against the
// fp32 reference - tests/test_gpu_parity.py::test_focal_*).  expf / logf / IEEE reciprocal cost ~40 more instructions per
// element and made the kernels transcendental-bound at a fifth of the HBM rate.
__device__ __forceinline__ float focal_sigmoid(float z) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + __expf(-z)));      // z << 0: 1 + inf -> 0, clamped to kLo below
    return r;
}

__device__ __forceinline__ void focal_term(float z, float g, float& pos, float& neg, int& npos) {
    const float s = focal_sigmoid(z);
    const float p = fminf(fmaxf(s, kLo), kHi);                    // rrnet_operator.py:55
    if (g == 1.0f) {                                              // functional.py:33,40
        const float q = 1.0f - p;
        pos += __logf(p) * (q * q);
        npos += 1;
    } else if (g < 1.0f) {                                        // :34,36,41
        const float q = 1.0f - g;
        const float w = (q * q) * (q * q);
        neg += __logf(1.0f - p) * (p * p) * w;
    }
}

// d loss / d z for one element, already multiplied by `scale` (= upstream * -1/num_pos)
__device__ __forceinline__ float focal_grad(float z, float g, float scale) {
    const float s = focal_sigmoid(z);
    if (s < kLo || s > kHi) return 0.0f;                          // clamp blocks the gradient
    const float p = s, q = 1.0f - s;
    float d = 0.0f;
    if (g == 1.0f) {
        d = q * q * q - 2.0f * p * (q * q) * __logf(p);
    } else if (g < 1.0f) {
        const float r = 1.0f - g;
        const float w = (r * r) * (r * r);
        d = w * (2.0f * (p * p) * q * __logf(q) - p * p * p);
    }
    return d * scale;
}

__device__ __forceinline__ void focal_accumulate(const float* __restrict__ logits, const float* __restrict__ gt,
                                                 long long n, float& pos, float& neg, int& npos) {
    const long long n4 = n >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(logits);
    const float4* g4 = reinterpret_cast<const float4*>(gt);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {                    // two independent float4 pairs in flight
        const float4 za = ld_stream_f4(z4 + i), ga = ld_stream_f4(g4 + i);
        const float4 zb = ld_stream_f4(z4 + i + stride), gb = ld_stream_f4(g4 + i + stride);
        focal_term(za.x, ga.x, pos, neg, npos); focal_term(za.y, ga.y, pos, neg, npos);
        focal_term(za.z, ga.z, pos, neg, npos); focal_term(za.w, ga.w, pos, neg, npos);
        focal_term(zb.x, gb.x, pos, neg, npos); focal_term(zb.y, gb.y, pos, neg, npos);
        focal_term(zb.z, gb.z, pos, neg, npos); focal_term(zb.w, gb.w, pos, neg, npos);
    }
    for (; i < n4; i += stride) {
        const float4 za = ld_stream_f4(z4 + i), ga = ld_stream_f4(g4 + i);
        focal_term(za.x, ga.x, pos, neg, npos); focal_term(za.y, ga.y, pos, neg, npos);
        focal_term(za.z, ga.z, pos, neg, npos); focal_term(za.w, ga.w, pos, neg, npos);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {               // tail (n not a multiple of 4)
        const long long j = (n4 << 2) + threadIdx.x;
        focal_term(logits[j], gt[j], pos, neg, npos);
    }
}

// block-level double reduction of (pos, neg, npos) -> partial[blockIdx]; returns true in the
// LAST block to finish (which then owns the final fixed-order sum).
__device__ __forceinline__ bool focal_block_reduce(float pos, float neg, int npos, double* partial,
                                                   unsigned int* ticket) {
    __shared__ double s_p[kFocalThreads / 32], s_n[kFocalThreads / 32], s_c[kFocalThreads / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dp = warp_sum((double)pos), dn = warp_sum((double)neg), dc = warp_sum((double)npos);
    if (lane == 0) { s_p[warp] = dp; s_n[warp] = dn; s_c[warp] = dc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < kFocalThreads / 32; ++w) { a += s_p[w]; b += s_n[w]; c += s_c[w]; }
        partial[blockIdx.x * 3 + 0] = a;
        partial[blockIdx.x * 3 + 1] = b;
        partial[blockIdx.x * 3 + 2] = c;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    return s_last;
}

// final sum by the last block, in a fixed order (thread t adds partials t, t+256, ...; then a fixed
// shuffle tree and a fixed sweep over the warps) -> bit-reproducible, and parallel
__device__ __forceinline__ void focal_finish(const double* partial, float* stats) {
    __shared__ double s_f[3][kFocalThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = 0, b = 0, c = 0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += kFocalThreads) {
        a += __ldcg(partial + k * 3 + 0);
        b += __ldcg(partial + k * 3 + 1);
        c += __ldcg(partial + k * 3 + 2);
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { s_f[0][warp] = a; s_f[1][warp] = b; s_f[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = 0;
        for (int w = 0; w < kFocalThreads / 32; ++w) { a += s_f[0][w]; b += s_f[1][w]; c += s_f[2][w]; }
        const double loss = (c == 0.0) ? -b : -(a + b) / c;            // functional.py:47-50
        stats[0] = (float)loss; stats[1] = (float)a; stats[2] = (float)b; stats[3] = (float)c;
    }
}

__global__ void __launch_bounds__(kFocalThreads)
focal_forward_kernel(const float* __restrict__ logits, const float* __restrict__ gt, long long n,
                     double* __restrict__ partial, unsigned int* __restrict__ ticket,
                     float* __restrict__ stats) {
    float pos = 0.f, neg = 0.f;
    int npos = 0;
    focal_accumulate(logits, gt, n, pos, neg, npos);
    if (focal_block_reduce(pos, neg, npos, partial, ticket)) {     // block-uniform: the last block to finish
        __threadfence();
        focal_finish(partial, stats);
        if (threadIdx.x == 0) *ticket = 0u;                       // ready for the next launch
    }
}

__device__ __forceinline__ void focal_write_grads(const float* __restrict__ logits, const float* __restrict__ gt,
                                                  long long n, float scale, float* __restrict__ grad) {
    const long long n4 = n >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(logits);
    const float4* g4 = reinterpret_cast<const float4*>(gt);
    float4* o4 = reinterpret_cast<float4*>(grad);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 z = __ldg(z4 + i), g = __ldg(g4 + i);
        float4 o;
        o.x = focal_grad(z.x, g.x, scale); o.y = focal_grad(z.y, g.y, scale);
        o.z = focal_grad(z.z, g.z, scale); o.w = focal_grad(z.w, g.w, scale);
        __stcs(o4 + i, o);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long j = (n4 << 2) + threadIdx.x;
        grad[j] = focal_grad(logits[j], gt[j], scale);
    }
}

__global__ void __launch_bounds__(kFocalThreads)
focal_backward_kernel(const float* __restrict__ logits, const float* __restrict__ gt, long long n,
                      const float* __restrict__ stats, float upstream, float* __restrict__ grad) {
    const float npos = stats[3];
    const float scale = upstream * ((npos == 0.f) ? -1.0f : -1.0f / npos);
    focal_write_grads(logits, gt, n, scale, grad);
}

__global__ void __launch_bounds__(kFocalThreads)
focal_fwd_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ gt, long long n,
                     float upstream, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                     float* __restrict__ stats, float* __restrict__ grad) {
    float pos = 0.f, neg = 0.f;
    int npos = 0;
    focal_accumulate(logits, gt, n, pos, neg, npos);
    if (focal_block_reduce(pos, neg, npos, partial, ticket)) {
        __threadfence();
        focal_finish(partial, stats);
        if (threadIdx.x == 0) { *ticket = 0u; __threadfence(); }
    }
    cg::this_grid().sync();
    const float np = __ldcg(stats + 3);
    const float scale = upstream * ((np == 0.f) ? -1.0f : -1.0f / np);
    focal_write_grads(logits, gt, n, scale, grad);
}

// --------------------------------------------------------------------------------------------
// Fused target render + focal loss: the ground-truth heat-map never reaches HBM.
// A CTA owns kFusedPix consecutive pixels of one (image, class) plane as a TILE IN SHARED MEMORY.  It filters the
// image's annotation rows for its class into shared memory (geometry as rr_render.cu), its warps splat the windows
// of those objects into the tile (one warp per object, hm = max(hm, G) by atomicMax on the uint bit pattern: the same
// values and the same order independence as render_kernel) - only window pixels are evaluated, like the stand-alone
// render - and then every thread takes its 16 pixels through the focal term and / or its derivative.
//   forward            : logits once                      (B*C*h*w*4 bytes)
//   backward           : logits once + gradient write
//   forward + backward : ONE pass, logits once + gradient write (42 MB at config 3, against 84 MB for render -> focal
//                        fwd+bwd): the gradient scale -1/num_pos is known up front because num_pos = the number of
//                        distinct (class, centre cell) pairs among the drawn objects (a Gaussian is 1 at its centre
//                        and nowhere else), counted by focal_count_pos_kernel from the annotations alone.
// --------------------------------------------------------------------------------------------
constexpr int kFusedPix = 4096;                    // pixels per CTA (16 per thread, four float4)
constexpr int kFusedObj = 256;                     // objects staged per round
constexpr int kCountMax = 2048;                    // objects per image the centre count stages in shared memory

struct FusedObj { float cxi, cyi, denom; int xa, xb, ya, yb; };

enum { kFusedFwd = 0, kFusedBwd = 1, kFusedBoth = 2 };

template <int kMode>
__device__ __forceinline__ void fused_body(const float* __restrict__ logits, const float* __restrict__ annos,
                                           const int* __restrict__ n_obj, int max_n, int img_w, int Hh, int Wh,
                                           float sf, int cls_num, int tiles, float scale, float* __restrict__ grad,
                                           float* __restrict__ gt_out, float& pos, float& neg, int& npos) {
    __shared__ unsigned int s_gt[kFusedPix];
    __shared__ FusedObj s_obj[kFusedObj];
    __shared__ int s_nobj;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane_id = blockIdx.x / tiles, tile = blockIdx.x - plane_id * tiles;     // plane = b * cls_num + c
    const int b = plane_id / cls_num, c = plane_id - b * cls_num;
    const int HW = Hh * Wh;
    const int base_px = tile * kFusedPix;
    const int row_lo = base_px / Wh, row_hi = min(base_px + kFusedPix - 1, HW - 1) / Wh;
    for (int i = tid; i < kFusedPix; i += kFocalThreads) s_gt[i] = 0u;
    const int nb = min(n_obj[b], max_n);
    for (int base = 0; base < nb; base += kFusedObj) {
        __syncthreads();                                                              // tile zeroed / previous round drawn
        if (tid == 0) s_nobj = 0;
        __syncthreads();
        const int k = base + tid;
        if (k < nb) {
            const ObjGauss o = obj_gauss(annos + ((size_t)b * max_n + k) * 8, img_w, Hh, Wh, sf, cls_num);
            // keep the objects of this class whose window meets this CTA's rows
            if (o.cls == c && o.xb > o.xa && o.yb > o.ya && o.yb > row_lo && o.ya <= row_hi) {
                const int slot = atomicAdd(&s_nobj, 1);
                s_obj[slot] = {o.cxi, o.cyi, o.denom, o.xa, o.xb, max(o.ya, row_lo), min(o.yb, row_hi + 1)};
            }
        }
        __syncthreads();
        const int n = s_nobj;
        for (int j = warp; j < n; j += kFocalThreads / 32) {                          // one warp per object window
            const FusedObj o = s_obj[j];
            const int ww = o.xb - o.xa, total = ww * (o.yb - o.ya);
            for (int t = lane; t < total; t += 32) {
                const int y = o.ya + t / ww, x = o.xa + t - (t / ww) * ww;
                const int idx = y * Wh + x - base_px;
                if ((unsigned)idx < (unsigned)kFusedPix) {                            // first / last row of a tile may be partial
                    const float dx = (float)x - o.cxi, dy = (float)y - o.cyi;
                    const float v = expf(-__fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), o.denom));   // == obj_value
                    atomicMax(&s_gt[idx], __float_as_uint(v));
                }
            }
        }
    }
    __syncthreads();
    const float* zp = logits + (size_t)plane_id * HW;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int lp = tid * 4 + q * (kFusedPix / 4), p = base_px + lp;               // float4 index stride keeps loads coalesced
        if (p >= HW) continue;                                                        // HW % 4 == 0 (checked by the host)
        const float4 z = __ldg(reinterpret_cast<const float4*>(zp + p));
        const uint4 gu = *reinterpret_cast<const uint4*>(&s_gt[lp]);
        const float g0 = __uint_as_float(gu.x), g1 = __uint_as_float(gu.y), g2 = __uint_as_float(gu.z), g3 = __uint_as_float(gu.w);
        if (gt_out) *reinterpret_cast<float4*>(gt_out + (size_t)plane_id * HW + p) = make_float4(g0, g1, g2, g3);
        if (kMode != kFusedFwd) {
            float4 o;
            o.x = focal_grad(z.x, g0, scale); o.y = focal_grad(z.y, g1, scale);
            o.z = focal_grad(z.z, g2, scale); o.w = focal_grad(z.w, g3, scale);
            __stcs(reinterpret_cast<float4*>(grad + (size_t)plane_id * HW + p), o);
        }
        if (kMode != kFusedBwd) {
            focal_term(z.x, g0, pos, neg, npos); focal_term(z.y, g1, pos, neg, npos);
            focal_term(z.z, g2, pos, neg, npos); focal_term(z.w, g3, pos, neg, npos);
        }
    }
}

__global__ void __launch_bounds__(kFocalThreads)
focal_render_forward_kernel(const float* __restrict__ logits, const float* __restrict__ annos,
                            const int* __restrict__ n_obj, int max_n, int img_w, int Hh, int Wh, float sf,
                            int cls_num, int tiles, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                            float* __restrict__ stats, float* __restrict__ gt_out) {
    float pos = 0.f, neg = 0.f;
    int npos = 0;
    fused_body<kFusedFwd>(logits, annos, n_obj, max_n, img_w, Hh, Wh, sf, cls_num, tiles, 0.f, nullptr, gt_out, pos, neg, npos);
    if (focal_block_reduce(pos, neg, npos, partial, ticket)) {
        __threadfence();
        focal_finish(partial, stats);
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

__global__ void __launch_bounds__(kFocalThreads)
focal_render_backward_kernel(const float* __restrict__ logits, const float* __restrict__ annos,
                             const int* __restrict__ n_obj, int max_n, int img_w, int Hh, int Wh, float sf,
                             int cls_num, int tiles, const float* __restrict__ stats, float upstream,
                             float* __restrict__ grad) {
    const float np = stats[3];
    const float scale = upstream * ((np == 0.f) ? -1.0f : -1.0f / np);
    float pos = 0.f, neg = 0.f;
    int npos = 0;
    fused_body<kFusedBwd>(logits, annos, n_obj, max_n, img_w, Hh, Wh, sf, cls_num, tiles, scale, grad, nullptr, pos, neg, npos);
}

// num_pos from the annotations: one CTA per image counts the drawn objects whose (class, centre cell) no earlier drawn
// object of the image shares.  An object is drawn iff its class is valid and its clipped window is not empty; the
// window then contains the centre (rr_gauss.cuh: ya <= cyi < yb, xa <= cxi < xb), where the Gaussian is exactly 1.
__global__ void __launch_bounds__(kFocalThreads)
focal_count_pos_kernel(const float* __restrict__ annos, const int* __restrict__ n_obj, int max_n, int img_w,
                       int Hh, int Wh, float sf, int cls_num, int* __restrict__ num_pos) {
    __shared__ int s_key[kCountMax];                 // (class * Hh + cy) * Wh + cx of a drawn object, -1 otherwise
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nb = min(min(n_obj[b], max_n), kCountMax);
    for (int k = tid; k < nb; k += kFocalThreads) {
        const ObjGauss o = obj_gauss(annos + ((size_t)b * max_n + k) * 8, img_w, Hh, Wh, sf, cls_num);
        const int cx = (int)o.cxi, cy = (int)o.cyi;
        const bool drawn = o.cls >= 0 && o.xb > o.xa && o.yb > o.ya && cy >= o.ya && cy < o.yb && cx >= o.xa && cx < o.xb;
        s_key[k] = drawn ? (o.cls * Hh + cy) * Wh + cx : -1;
    }
    __syncthreads();
    int mine = 0;
    for (int k = tid; k < nb; k += kFocalThreads) {
        const int key = s_key[k];
        if (key < 0) continue;
        bool first = true;
        for (int j = 0; j < k; ++j) first &= (s_key[j] != key);
        mine += first;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((tid & 31) == 0 && mine) atomicAdd(num_pos, mine);
}

__global__ void __launch_bounds__(kFocalThreads)
focal_render_fwd_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ annos,
                            const int* __restrict__ n_obj, int max_n, int img_w, int Hh, int Wh, float sf,
                            int cls_num, int tiles, const int* __restrict__ num_pos, float upstream,
                            double* __restrict__ partial, unsigned int* __restrict__ ticket,
                            float* __restrict__ stats, float* __restrict__ grad) {
    const int np = *num_pos;
    const float scale = upstream * ((np == 0) ? -1.0f : -1.0f / (float)np);
    float pos = 0.f, neg = 0.f;
    int npos = 0;
    fused_body<kFusedBoth>(logits, annos, n_obj, max_n, img_w, Hh, Wh, sf, cls_num, tiles, scale, grad, nullptr, pos, neg, npos);
    if (focal_block_reduce(pos, neg, npos, partial, ticket)) {
        __threadfence();
        focal_finish(partial, stats);                             // stats[3] = the gt == 1 elements seen: equals *num_pos
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace rr

using namespace rr;

RR_API size_t rr_focal_workspace_bytes(int64_t n) {
    (void)n;
    return carve_focal(nullptr).bytes;
}

static int focal_check(const float* logits, const float* gt, int64_t n, const void* ws, size_t ws_bytes,
                       bool need_ws) {
    if (!logits || !gt || n <= 0) return RR_E_BADARG;
    if (((uintptr_t)logits & 15) || ((uintptr_t)gt & 15)) return RR_E_ALIGN;
    if (need_ws) {
        if (!ws) return RR_E_BADARG;
        if (ws_bytes < carve_focal(nullptr).bytes || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    }
    return 0;
}

RR_API int rr_focal_forward(const float* logits, const float* gt, int64_t n, float* stats,
                            void* ws, size_t ws_bytes, void* stream) {
    int rc = focal_check(logits, gt, n, ws, ws_bytes, true);
    if (rc) return rc;
    if (!stats) return RR_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    FocalWs w = carve_focal(ws);
    RR_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned int), st), rc);
    focal_forward_kernel<<<focal_grid(n), kFocalThreads, 0, st>>>(logits, gt, n, w.partial, w.ticket, stats);
    RR_LAUNCHED_K(rc, "focal_forward_kernel", st);
    return rc;
}

RR_API int rr_focal_backward(const float* logits, const float* gt, int64_t n, const float* stats,
                             float upstream, float* grad, void* stream) {
    int rc = focal_check(logits, gt, n, nullptr, 0, false);
    if (rc) return rc;
    if (!stats || !grad) return RR_E_BADARG;
    if ((uintptr_t)grad & 15) return RR_E_ALIGN;
    focal_backward_kernel<<<focal_grid(n), kFocalThreads, 0, (cudaStream_t)stream>>>(logits, gt, n, stats, upstream, grad);
    RR_LAUNCHED_K(rc, "focal_backward_kernel", (cudaStream_t)stream);
    return rc;
}

RR_API int rr_focal_fwd_bwd(const float* logits, const float* gt, int64_t n, float upstream,
                            float* stats, float* grad, void* ws, size_t ws_bytes, void* stream) {
    int rc = focal_check(logits, gt, n, ws, ws_bytes, true);
    if (rc) return rc;
    if (!stats || !grad) return RR_E_BADARG;
    if ((uintptr_t)grad & 15) return RR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    FocalWs w = carve_focal(ws);
    RR_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned int), st), rc);
    // cooperative launch: every CTA must be co-resident
    int dev = 0, sms = kSMs, per_sm = 0;
    RR_CUDA(cudaGetDevice(&dev), rc);
    RR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), rc);
    RR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, focal_fwd_bwd_kernel, kFocalThreads, 0), rc);
    if (rc) return rc;
    int grid = min(focal_grid(n), max(1, per_sm * sms));
    long long n_ll = n;
    void* args[] = {(void*)&logits, (void*)&gt, (void*)&n_ll, (void*)&upstream, (void*)&w.partial,
                    (void*)&w.ticket, (void*)&stats, (void*)&grad};
    RR_CUDA(cudaLaunchCooperativeKernel((void*)focal_fwd_bwd_kernel, dim3(grid), dim3(kFocalThreads), args, 0, st), rc);
    RR_LAUNCHED_K(rc, "focal_fwd_bwd_kernel", st);
    return rc;
}

// ------------------------------------------------------------------ fused render + focal
static int fused_dims(int B, int cls_num, int img_h, int img_w, int sf, int* Hh, int* Wh, int* tiles, long long* grid) {
    if (B <= 0 || cls_num <= 0 || img_h <= 0 || img_w <= 0 || sf <= 0) return RR_E_BADARG;
    *Hh = img_h / sf; *Wh = img_w / sf;
    const long long HW = (long long)*Hh * *Wh;
    if (HW <= 0 || (*Wh & 3)) return RR_E_RANGE;                     // rows are read as float4 groups
    *tiles = (int)((HW + kFusedPix - 1) / kFusedPix);
    *grid = (long long)B * cls_num * *tiles;
    if (*grid > 0x7fffffffLL) return RR_E_RANGE;
    return 0;
}

RR_API size_t rr_focal_render_workspace_bytes(int B, int cls_num, int img_h, int img_w, int scale_factor) {
    int Hh, Wh, tiles; long long grid;
    if (fused_dims(B, cls_num, img_h, img_w, scale_factor, &Hh, &Wh, &tiles, &grid)) return 0;
    return carve_focal(nullptr, (size_t)grid).bytes;
}

RR_API int rr_focal_render_forward(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                                   int img_h, int img_w, int scale_factor, int cls_num,
                                   float* stats, float* gt_out, void* ws, size_t ws_bytes, void* stream) {
    int Hh, Wh, tiles; long long grid;
    int rc = fused_dims(B, cls_num, img_h, img_w, scale_factor, &Hh, &Wh, &tiles, &grid);
    if (rc) return rc;
    if (!logits || !stats || !ws || max_n < 0 || (max_n > 0 && (!annos || !n_obj))) return RR_E_BADARG;
    if (((uintptr_t)logits & 15) || ((uintptr_t)gt_out & 15)) return RR_E_ALIGN;
    if (ws_bytes < carve_focal(nullptr, (size_t)grid).bytes || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    if (max_n == 0 && !n_obj) return RR_E_BADARG;                    // n_obj is always read (may hold zeros)
    cudaStream_t st = (cudaStream_t)stream;
    FocalWs w = carve_focal(ws, (size_t)grid);
    RR_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned int), st), rc);
    focal_render_forward_kernel<<<(unsigned)grid, kFocalThreads, 0, st>>>(logits, annos, n_obj, max_n, img_w, Hh, Wh,
                                                                         (float)scale_factor, cls_num, tiles,
                                                                         w.partial, w.ticket, stats, gt_out);
    RR_LAUNCHED_K(rc, "focal_render_forward_kernel", st);
    return rc;
}

RR_API int rr_focal_render_backward(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                                    int img_h, int img_w, int scale_factor, int cls_num,
                                    const float* stats, float upstream, float* grad, void* stream) {
    int Hh, Wh, tiles; long long grid;
    int rc = fused_dims(B, cls_num, img_h, img_w, scale_factor, &Hh, &Wh, &tiles, &grid);
    if (rc) return rc;
    if (!logits || !stats || !grad || !n_obj || max_n < 0 || (max_n > 0 && !annos)) return RR_E_BADARG;
    if (((uintptr_t)logits & 15) || ((uintptr_t)grad & 15)) return RR_E_ALIGN;
    focal_render_backward_kernel<<<(unsigned)grid, kFocalThreads, 0, (cudaStream_t)stream>>>(
        logits, annos, n_obj, max_n, img_w, Hh, Wh, (float)scale_factor, cls_num, tiles, stats, upstream, grad);
    RR_LAUNCHED_K(rc, "focal_render_backward_kernel", (cudaStream_t)stream);
    return rc;
}

RR_API int rr_focal_render_fwd_bwd(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                                   int img_h, int img_w, int scale_factor, int cls_num, float upstream,
                                   float* stats, float* grad, void* ws, size_t ws_bytes, void* stream) {
    int Hh, Wh, tiles; long long grid;
    int rc = fused_dims(B, cls_num, img_h, img_w, scale_factor, &Hh, &Wh, &tiles, &grid);
    if (rc) return rc;
    if (!logits || !stats || !grad || !ws || !n_obj || max_n < 0 || (max_n > 0 && !annos)) return RR_E_BADARG;
    if (((uintptr_t)logits & 15) || ((uintptr_t)grad & 15)) return RR_E_ALIGN;
    if (ws_bytes < carve_focal(nullptr, (size_t)grid).bytes || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    FocalWs w = carve_focal(ws, (size_t)grid);
    RR_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned int), st), rc);
    if (max_n > kCountMax) {           // more objects per image than the centre count stages: two passes (forward, then backward)
        focal_render_forward_kernel<<<(unsigned)grid, kFocalThreads, 0, st>>>(logits, annos, n_obj, max_n, img_w, Hh, Wh,
                                                                             (float)scale_factor, cls_num, tiles,
                                                                             w.partial, w.ticket, stats, nullptr);
        RR_LAUNCHED_K(rc, "focal_render_forward_kernel", st);
        focal_render_backward_kernel<<<(unsigned)grid, kFocalThreads, 0, st>>>(
            logits, annos, n_obj, max_n, img_w, Hh, Wh, (float)scale_factor, cls_num, tiles, stats, upstream, grad);
        RR_LAUNCHED_K(rc, "focal_render_backward_kernel", st);
        return rc;
    }
    RR_CUDA(cudaMemsetAsync(w.num_pos, 0, sizeof(int), st), rc);
    if (max_n > 0) {
        focal_count_pos_kernel<<<B, kFocalThreads, 0, st>>>(annos, n_obj, max_n, img_w, Hh, Wh, (float)scale_factor, cls_num,
                                                            w.num_pos);
        RR_LAUNCHED_K(rc, "focal_count_pos_kernel", st);
    }
    focal_render_fwd_bwd_kernel<<<(unsigned)grid, kFocalThreads, 0, st>>>(logits, annos, n_obj, max_n, img_w, Hh, Wh,
                                                                         (float)scale_factor, cls_num, tiles, w.num_pos,
                                                                         upstream, w.partial, w.ticket, stats, grad);
    RR_LAUNCHED_K(rc, "focal_render_fwd_bwd_kernel", st);
    return rc;
}
