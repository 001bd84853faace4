// RoIAlign (3x3 bins, adaptive sampling, aligned=False) with the ReLU of the feature map fused
// in, for sm_100a.
//
// Replaces torchvision.ops.roi_align(torch.relu(pre_feat[-1]), bxyxys, (3,3))
// (models/rrnet.py:51): the reference first writes a full ReLU copy of the feature map
// (2 x 1.07 GB of HBM traffic at config 2) and then gathers 4 taps per sample per channel.
//
// Formulation.  For one RoI the bilinear sample positions and weights are the same for all
// channels and the bilinear sum is separable:
//     out[c,ph,pw] = (1/count) * sum_Y sum_X Ay[ph][Y] * Ax[pw][X] * relu(F[c,Y,X])
// where Ay[ph][Y] (resp. Ax[pw][X]) accumulates, over the samples of bin ph (pw), the weight
// that sample puts on pixel row Y (column X): hy on y_low, ly on y_high, 0 if the sample is
// outside [-1, H] (torchvision zeroes those samples).  A sample is valid iff both its row and
// its column are valid, so validity is separable too.  Every pixel therefore contributes to a
// RoI independently of its neighbours, which allows the sum to be cut along tile borders.
//
// Two kernels implement it:
//
//  * TILE path (default).  The RoI windows of a detector's top-K boxes overlap heavily (4.5x at
//    config 2: 5.3 GB of window bytes over a 1.07 GB map), so gathering per RoI is bound by L2,
//    not HBM.  Instead the map is cut into 32x24-pixel tiles; a work item is (tile, 32 channels,
//    <= 32 RoI pieces).  A persistent CTA brings the tile once from HBM into shared memory - by
//    ONE TMA tensor copy per tile into a swizzled [y][c][x] layout (roi_tile_tma_kernel, the
//    default), or with coalesced 128-byte row loads, ReLU on the way in and an odd channel stride
//    (roi_tile_kernel: W % 4 != 0, or algo = 2) - and every warp then evaluates (piece, bin
//    column) units with lane = channel: no cross-lane reduction, all 32 lanes busy whatever the
//    piece shape, weights warp-uniform, bank-conflict free in both layouts.  Pieces of one RoI land in private partial slots
//    ([slot][9][C], coalesced stores) and a combine kernel adds them in a fixed order, so the
//    result is deterministic (no floating-point atomics).  Each feature element is read from
//    HBM once per step (tiles without RoIs are never read).
//    Supporting kernels: roi_prep (geometry, separable weights, partial-slot and direct-list
//    allocation, per-tile piece counts; its last CTA lays out the tile lists and work items),
//    roi_fill (tile lists), roi_combine.
//
//  * DIRECT path.  One CTA per RoI walks the RoI window straight from global memory with
//    lanes on consecutive columns.  Used for RoIs whose window exceeds 64 pixels on an axis,
//    when the partial-slot budget is exhausted, when C is not a multiple of 32, or on request
//    (algo = 1).
//
// The separable order of summation differs from torchvision's sample-by-sample sum at the
// 1e-7 level (all terms are >= 0 after ReLU, so there is no cancellation); parity is held to
// 1e-5 relative against the CPU oracle.
#include <cuda.h>

#include "rr_common.cuh"

namespace rr {

constexpr int kRoiThreads = 256;
constexpr int kMaxWin = 96;          // direct path: window rows/cols held in shared memory

constexpr int kTW = 32;              // tile width  (one 128-byte line per (channel,row))
constexpr int kTH = 24;              // tile height
constexpr int kTC = 32;              // channels per work item (lane = channel)
constexpr int kChStride = kTW * kTH + 1;   // odd: lane = channel reads hit 32 different banks
constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kChunk = 32;           // RoI pieces per work item
constexpr int kTileTileFloats = (kTC * kChStride + 3) / 4 * 4;      // tile, padded to 16 bytes
constexpr int kTileSmem = (kTileTileFloats + kChunk * 2 * 4 + kChunk * kTH * 4) * (int)sizeof(float);   // + piece descriptors + row weights
constexpr int kMaxWinT = 64;         // tile path: window extent limit per axis
constexpr int kMaxPieces = 12;       // ceil-spans of a 64-window: 3 tile columns x 4 tile rows
constexpr int kSlotsPerRoi = 4;      // partial-slot budget: kSlotsPerRoi * n_cap + #tiles

enum { kFlagTile = 0, kFlagDirect = 1, kFlagZero = 2 };
enum { kCtlItems = 0, kCtlTicket = 1, kCtlDirect = 2, kCtlSlots = 3, kCtlPrepDone = 4, kCtlWords = 8 };

// Per-axis sample geometry of one RoI (torchvision roi_align_kernel: pre_calc_for_bilinear_interpolate).
struct AxisGeom {
    float start, bin;   // roi start, bin size
    int grid;           // samples per bin
    int lo, n;          // first pixel touched, number of pixels touched (window extent)
};

struct __align__(16) RoiPrep {
    int img, flags;
    int x_lo, nx, y_lo, ny;           // pixel window touched by valid samples
    int tx0, ty0, ntx, nty;           // tile range of the window
    int slot_base;                    // first partial slot (ntx*nty consecutive slots)
    float count;                      // samples per bin (divisor)
    signed char cx_lo[4], cx_hi[4];   // per bin column: window-relative pixel range (lo > hi: empty)
    signed char cy_lo[4], cy_hi[4];
};
static_assert(sizeof(RoiPrep) == 64, "RoiPrep is 64 bytes");

// position/weights of sample (p, i) on one axis; returns false when the sample is out of range
__device__ __forceinline__ bool axis_sample(float start, float bin, int grid, int p, int i, int size,
                                            int& low, int& high, float& l, float& h) {
    float v = start + p * bin + (float)(i + .5f) * bin / (float)grid;
    if (v < -1.0f || v > (float)size) return false;
    if (v <= 0.f) v = 0.f;
    low = (int)v;
    if (low >= size - 1) { high = low = size - 1; v = (float)low; } else high = low + 1;
    l = v - (float)low;
    h = 1.f - l;
    return true;
}

__device__ __forceinline__ AxisGeom axis_geom(float a, float b, int size) {
    AxisGeom g;
    float len = b - a;
    len = fmaxf(len, 1.0f);                       // aligned=False: roi size at least 1
    g.start = a;
    g.bin = len / (float)RR_POOL;
    g.grid = (int)ceilf(len / (float)RR_POOL);
    // pixel extent touched by valid samples
    int lo = size, hi = -1;
    // first and last sample positions bound the extent (positions are monotone in (p,i))
    for (int e = 0; e < 2; ++e) {
        // scan from the low end (e=0) / high end (e=1) for the first valid sample
        int total = RR_POOL * g.grid;
        for (int t = 0; t < total; ++t) {
            int s = e == 0 ? t : total - 1 - t;
            int p = s / g.grid, i = s - p * g.grid;
            int l0, h0; float wl, wh;
            if (axis_sample(g.start, g.bin, g.grid, p, i, size, l0, h0, wl, wh)) {
                lo = min(lo, l0);
                hi = max(hi, h0);
                break;
            }
        }
    }
    g.lo = lo;
    g.n = (hi >= lo) ? (hi - lo + 1) : 0;
    return g;
}

// weight that bin p puts on pixel (lo + k): sum over the bin's samples, in sample order
__device__ __forceinline__ float axis_weight(const AxisGeom& g, int p, int k, int size) {
    float acc = 0.f;
    const int pix = g.lo + k;
    for (int i = 0; i < g.grid; ++i) {
        int l0, h0; float wl, wh;
        if (!axis_sample(g.start, g.bin, g.grid, p, i, size, l0, h0, wl, wh)) continue;
        if (l0 == pix) acc += wh;
        if (h0 == pix) acc += wl;                 // l0 == h0 at the far edge: wl == 0 there
    }
    return acc;
}

// --------------------------------------------------------------------------------------------
// DIRECT path: one RoI per CTA iteration, straight from global memory.
// --------------------------------------------------------------------------------------------
__device__ void roi_direct_one(const float* __restrict__ feat, const float* __restrict__ r,
                               int B, int C, int H, int W, int relu, float* __restrict__ o,
                               float (*s_ax)[kMaxWin], float (*s_ay)[kMaxWin], AxisGeom* s_g) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = (int)r[0];
    if (bi < 0 || bi >= B) {                       // invalid image index: defined output, no fault
        for (int i = tid; i < C * RR_POOL * RR_POOL; i += blockDim.x) o[i] = 0.f;
        return;
    }
    __syncthreads();                               // previous iteration done with the shared tables
    if (tid == 0) s_g[0] = axis_geom(r[1], r[3], W);
    if (tid == 32) s_g[1] = axis_geom(r[2], r[4], H);
    __syncthreads();
    const AxisGeom gx = s_g[0], gy = s_g[1];
    const float count = (float)max(gx.grid * gy.grid, 1);
    if (gx.n == 0 || gy.n == 0) {                  // RoI entirely outside the map -> zeros
        for (int i = tid; i < C * RR_POOL * RR_POOL; i += blockDim.x) o[i] = 0.f;
        return;
    }
    const bool fits = gx.n <= kMaxWin && gy.n <= kMaxWin;
    if (fits) {
        for (int t = tid; t < RR_POOL * gx.n; t += blockDim.x) {
            int p = t / gx.n, k = t - p * gx.n;
            s_ax[p][k] = axis_weight(gx, p, k, W);
        }
        for (int t = tid; t < RR_POOL * gy.n; t += blockDim.x) {
            int p = t / gy.n, k = t - p * gy.n;
            s_ay[p][k] = axis_weight(gy, p, k, H);
        }
    }
    __syncthreads();

    const float* fb = feat + (size_t)bi * C * H * W;
    const int nwarp = blockDim.x >> 5;
    for (int c = warp; c < C; c += nwarp) {
        const float* plane = fb + (size_t)c * H * W + (size_t)gy.lo * W + gx.lo;
        float acc[RR_POOL][RR_POOL];
#pragma unroll
        for (int a = 0; a < RR_POOL; ++a)
#pragma unroll
            for (int q = 0; q < RR_POOL; ++q) acc[a][q] = 0.f;
        for (int x0 = 0; x0 < gx.n; x0 += 32) {
            const int x = x0 + lane;
            const bool on = x < gx.n;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f;
            if (fits) {
                int y = 0;
                for (; y + 3 < gy.n; y += 4) {       // 4 independent row loads in flight
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = on ? __ldg(plane + (size_t)(y + u) * W + x) : 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float f = relu ? fmaxf(v[u], 0.f) : v[u];
                        t0 = fmaf(s_ay[0][y + u], f, t0);
                        t1 = fmaf(s_ay[1][y + u], f, t1);
                        t2 = fmaf(s_ay[2][y + u], f, t2);
                    }
                }
                for (; y < gy.n; ++y) {
                    float f = on ? __ldg(plane + (size_t)y * W + x) : 0.f;
                    if (relu) f = fmaxf(f, 0.f);
                    t0 = fmaf(s_ay[0][y], f, t0);
                    t1 = fmaf(s_ay[1][y], f, t1);
                    t2 = fmaf(s_ay[2][y], f, t2);
                }
            } else {                               // huge RoI: weights recomputed on the fly
                for (int y = 0; y < gy.n; ++y) {
                    float f = on ? __ldg(plane + (size_t)y * W + x) : 0.f;
                    if (relu) f = fmaxf(f, 0.f);
                    t0 = fmaf(axis_weight(gy, 0, y, H), f, t0);
                    t1 = fmaf(axis_weight(gy, 1, y, H), f, t1);
                    t2 = fmaf(axis_weight(gy, 2, y, H), f, t2);
                }
            }
            float ax0 = 0.f, ax1 = 0.f, ax2 = 0.f;
            if (on) {
                if (fits) { ax0 = s_ax[0][x]; ax1 = s_ax[1][x]; ax2 = s_ax[2][x]; }
                else { ax0 = axis_weight(gx, 0, x, W); ax1 = axis_weight(gx, 1, x, W); ax2 = axis_weight(gx, 2, x, W); }
            }
            acc[0][0] = fmaf(t0, ax0, acc[0][0]); acc[0][1] = fmaf(t0, ax1, acc[0][1]); acc[0][2] = fmaf(t0, ax2, acc[0][2]);
            acc[1][0] = fmaf(t1, ax0, acc[1][0]); acc[1][1] = fmaf(t1, ax1, acc[1][1]); acc[1][2] = fmaf(t1, ax2, acc[1][2]);
            acc[2][0] = fmaf(t2, ax0, acc[2][0]); acc[2][1] = fmaf(t2, ax1, acc[2][1]); acc[2][2] = fmaf(t2, ax2, acc[2][2]);
        }
        // 9-value warp reduction; lane q (< 9) ends up writing output q
        float mine = 0.f;
#pragma unroll
        for (int a = 0; a < RR_POOL; ++a)
#pragma unroll
            for (int q = 0; q < RR_POOL; ++q) {
                float s = warp_sum(acc[a][q]);
                if (lane == a * RR_POOL + q) mine = s;
            }
        if (lane < RR_POOL * RR_POOL) o[(size_t)c * (RR_POOL * RR_POOL) + lane] = mine / count;
    }
}

// Grid-stride over the list of RoIs routed to the direct path (built by roi_prep).
__global__ void __launch_bounds__(kRoiThreads)
roi_direct_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                  const int* __restrict__ direct_list, const int* __restrict__ ctl,
                  int B, int C, int H, int W, int relu, float* __restrict__ out) {
    RR_PDL_PROLOGUE();
    __shared__ float s_ax[RR_POOL][kMaxWin];
    __shared__ float s_ay[RR_POOL][kMaxWin];
    __shared__ AxisGeom s_g[2];
    const int n_direct = ctl[kCtlDirect];
    for (int i = blockIdx.x; i < n_direct; i += gridDim.x) {
        const int n = direct_list[i];
        roi_direct_one(feat, rois + (size_t)n * 5, B, C, H, W, relu,
                       out + (size_t)n * C * (RR_POOL * RR_POOL), s_ax, s_ay, s_g);
    }
}

// --------------------------------------------------------------------------------------------
// TILE path, step 1: per-RoI geometry, separable weights, tile range, per-tile piece counts.
// One warp per RoI.  meta[n] = number of tile pieces (0: output is all zeros), or -1: direct path.
// --------------------------------------------------------------------------------------------
struct TileDims { int ntx, nty, tiles_per_img; };

constexpr int kMaxSamples = 72;      // 3 * grid, grid <= 22 for windows of at most kMaxWinT pixels

// Per-warp table of one axis' bilinear samples (in sample order) + per-bin pixel ranges.
struct AxisTable { int low[kMaxSamples]; int high[kMaxSamples]; float wl[kMaxSamples]; };

__device__ __forceinline__ void axis_table(const AxisGeom& g, int size, int lane, AxisTable& tb,
                                           signed char* lo3, signed char* hi3) {
    int lo[RR_POOL], hi[RR_POOL];
#pragma unroll
    for (int p = 0; p < RR_POOL; ++p) { lo[p] = 127; hi[p] = -1; }
    const int total = RR_POOL * g.grid;
    for (int t = lane; t < total; t += 32) {
        const int p = t / g.grid, i = t - p * g.grid;
        int l0 = -1, h0 = -1; float wl = 0.f, wh;
        if (axis_sample(g.start, g.bin, g.grid, p, i, size, l0, h0, wl, wh)) {
#pragma unroll
            for (int q = 0; q < RR_POOL; ++q)
                if (q == p) { lo[q] = min(lo[q], l0 - g.lo); hi[q] = max(hi[q], h0 - g.lo); }
        } else {                                   // out of range: below the map -1, above it INT_MAX, so that `low`
            const float v = g.start + p * g.bin + (float)(i + .5f) * g.bin / (float)g.grid;   // stays non-decreasing
            l0 = h0 = (v < -1.0f) ? -1 : 0x7fffffff;
        }
        tb.low[t] = l0; tb.high[t] = h0; tb.wl[t] = wl;
    }
#pragma unroll
    for (int p = 0; p < RR_POOL; ++p) {
        lo3[p] = (signed char)__reduce_min_sync(0xffffffffu, lo[p]);
        hi3[p] = (signed char)__reduce_max_sync(0xffffffffu, hi[p]);
    }
    lo3[3] = 127; hi3[3] = -1;
    __syncwarp();
}

// weight bin p puts on pixel pix: the table's samples of that bin, in sample order (== axis_weight).  Sample positions
// are monotone, so the samples that touch pix (low == pix, or low == pix - 1 through `high`) are a contiguous run: found
// by bisection instead of walking all `grid` samples (the walk made roi_prep issue bound: 36 us at config 2).
__device__ __forceinline__ float table_weight(const AxisTable& tb, int grid, int p, int pix) {
    float acc = 0.f;
    const int t0 = p * grid;
    int a = 0, b = grid;                           // first sample with low >= pix - 1
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (tb.low[t0 + mid] < pix - 1) a = mid + 1; else b = mid;
    }
    for (int i = a; i < grid; ++i) {
        const int l0 = tb.low[t0 + i], h0 = tb.high[t0 + i];
        if (l0 > pix) break;
        const float wl = tb.wl[t0 + i];
        if (l0 == pix) acc += 1.f - wl;
        if (h0 == pix && l0 >= 0) acc += wl;
    }
    return acc;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = lane_id(), warp = warp_id(), nwarp = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                               // s_warp reuse across calls
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? s_warp[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;                   // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + inc - v;
}

#ifndef RR_PREP_MINB
#define RR_PREP_MINB 5           // 48 registers: five CTAs per SM instead of three (measured 42.0 -> 37.9 us)
#endif
__global__ void __launch_bounds__(256, RR_PREP_MINB)
roi_prep_kernel(const float* __restrict__ rois, const int* __restrict__ n_rois_dev, int n_cap,
                int B, int C, int H, int W, int force_direct, TileDims td, int n_tiles, int slot_cap,
                RoiPrep* __restrict__ prep, int* __restrict__ meta, int* __restrict__ slot, float* __restrict__ cnt_arr,
                float* __restrict__ wx, float4* __restrict__ wy4, int* __restrict__ tile_count,
                int* __restrict__ tile_off, int4* __restrict__ items, int* __restrict__ direct_list, int* __restrict__ ctl) {
    RR_PDL_PROLOGUE();
    __shared__ AxisTable s_tab[8][2];
    __shared__ int s_warp[33];
    __shared__ int s_last, s_m[8], s_off[8], s_base;
    const int lane = lane_id();
    const int n = blockIdx.x * (blockDim.x >> 5) + warp_id();
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    AxisTable& tx = s_tab[warp_id()][0];
    AxisTable& ty = s_tab[warp_id()][1];
    RoiPrep rp;
    AxisGeom gx, gy;
    gx.n = gy.n = 0; gx.grid = gy.grid = 1; gx.lo = gy.lo = 0; gx.start = gy.start = gx.bin = gy.bin = 0.f;
    int m_warp = 0;
    if (n < live) {
        const float* r = rois + (size_t)n * 5;
        rp.img = (int)r[0];
        rp.flags = kFlagZero;
        rp.x_lo = rp.nx = rp.y_lo = rp.ny = 0;
        rp.tx0 = rp.ty0 = rp.ntx = rp.nty = 0;
        rp.slot_base = 0;
        rp.count = 1.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) { rp.cx_lo[p] = rp.cy_lo[p] = 127; rp.cx_hi[p] = rp.cy_hi[p] = -1; }
        int m = 0;
        if (rp.img >= 0 && rp.img < B) {
            gx = axis_geom(r[1], r[3], W); gy = axis_geom(r[2], r[4], H);
            if (gx.n > 0 && gy.n > 0) {
                rp.x_lo = gx.lo; rp.nx = gx.n; rp.y_lo = gy.lo; rp.ny = gy.n;
                rp.count = (float)max(gx.grid * gy.grid, 1);
                if (force_direct || gx.n > kMaxWinT || gy.n > kMaxWinT || (C % kTC) != 0 ||
                    RR_POOL * gx.grid > kMaxSamples || RR_POOL * gy.grid > kMaxSamples) {
                    rp.flags = kFlagDirect;
                    m = -1;
                } else {
                    rp.flags = kFlagTile;
                    rp.tx0 = gx.lo / kTW; rp.ntx = (gx.lo + gx.n - 1) / kTW - rp.tx0 + 1;
                    rp.ty0 = gy.lo / kTH; rp.nty = (gy.lo + gy.n - 1) / kTH - rp.ty0 + 1;
                    m = rp.ntx * rp.nty;
                }
            }
        }
        m_warp = m;
    }
    // partial slots (one per tile piece) and the direct list: handed out with ONE atomic per CTA (8 RoIs).  The slot
    // NUMBERS therefore change from run to run; the results do not (a RoI's pieces are summed in the order of its own slots).
    if (lane == 0) s_m[warp_id()] = m_warp;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int q = 0; q < 8; ++q) { const int v = s_m[q]; s_off[q] = tot; tot += v > 0 ? v : 0; }
        s_base = tot ? atomicAdd(ctl + kCtlSlots, tot) : 0;
    }
    __syncthreads();
    if (n < live) {
        const int m = m_warp;
        int sb = 0;
        if (lane == 0) {
            if (m > 0) {
                sb = s_base + s_off[warp_id()];
                if (sb + m > slot_cap) { sb = -1; }                    // slot budget exhausted: direct path
            } else if (m < 0) {
                sb = -1;
            }
            if (sb < 0) direct_list[atomicAdd(ctl + kCtlDirect, 1)] = n;
        }
        sb = __shfl_sync(0xffffffffu, sb, 0);
        if (m > 0 && sb >= 0) {
            axis_table(gx, W, lane, tx, rp.cx_lo, rp.cx_hi);
            axis_table(gy, H, lane, ty, rp.cy_lo, rp.cy_hi);
            float* wxn = wx + (size_t)n * (RR_POOL * kMaxWinT);
            float4* wyn = wy4 + (size_t)n * kMaxWinT;
            // One (axis, bin, pixel) weight per lane and round, only for the pixels inside the bin's range (a bin covers
            // about a third of the window): with lanes on pixels and a loop over the three bins of both axes every
            // round ran six divergent table look-ups, most of them for lanes outside the bin.  Column weights outside a
            // bin's range are never read (roi_fill masks by the range); row weights outside it must be zero.
            for (int k = lane; k < gy.n; k += 32) wyn[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            int cnt[2 * RR_POOL], total = 0;
#pragma unroll
            for (int p = 0; p < RR_POOL; ++p) {
                cnt[p] = max(rp.cx_hi[p] - rp.cx_lo[p] + 1, 0);
                cnt[RR_POOL + p] = max(rp.cy_hi[p] - rp.cy_lo[p] + 1, 0);
            }
#pragma unroll
            for (int q = 0; q < 2 * RR_POOL; ++q) total += cnt[q];
            for (int t = lane; t < total; t += 32) {
                int q = 0, r = t, lo = rp.cx_lo[0];
#pragma unroll
                for (int j = 0; j < 2 * RR_POOL - 1; ++j)
                    if (q == j && r >= cnt[j]) {
                        r -= cnt[j]; q = j + 1;
                        lo = j + 1 < RR_POOL ? rp.cx_lo[(j + 1) % RR_POOL] : rp.cy_lo[(j + 1) % RR_POOL];
                    }
                const bool is_y = q >= RR_POOL;
                const int p = is_y ? q - RR_POOL : q;
                const int k = lo + r;
                const float wgt = table_weight(is_y ? ty : tx, is_y ? gy.grid : gx.grid, p, (is_y ? gy.lo : gx.lo) + k);
                if (is_y) reinterpret_cast<float*>(wyn + k)[p] = wgt;
                else wxn[p * kMaxWinT + k] = wgt;
            }
            if (lane < m) {
                const int tyi = rp.ty0 + lane / rp.ntx, txi = rp.tx0 + lane % rp.ntx;
                atomicAdd(tile_count + rp.img * td.tiles_per_img + tyi * td.ntx + txi, 1);
            }
        }
        if (lane == 0) { prep[n] = rp; meta[n] = m; slot[n] = sb; cnt_arr[n] = rp.count; }
    }
    // ---- the last CTA to finish lays out the tile lists: list offset per tile and the work items (tile, list start,
    //      piece count), in tile order (this used to be a separate one-CTA kernel) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ctl + kCtlPrepDone, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const int tid = threadIdx.x, nt = blockDim.x;
        const int per = (n_tiles + nt - 1) / nt;
        const int t0 = min(tid * per, n_tiles), t1 = min(t0 + per, n_tiles);
        int sum = 0, chunks = 0;
        for (int t = t0; t < t1; ++t) { const int c = __ldcg(tile_count + t); sum += c; chunks += (c + kChunk - 1) / kChunk; }
        int total, n_items;
        int off = block_exclusive_scan(sum, s_warp, total);
        int ioff = block_exclusive_scan(chunks, s_warp, n_items);
        for (int t = t0; t < t1; ++t) {
            const int c = __ldcg(tile_count + t);
            tile_off[t] = off;
            for (int k = 0; k < c; k += kChunk) items[ioff++] = make_int4(t, off + k, min(kChunk, c - k), 0);
            off += c;
        }
        if (tid == 0) { tile_off[n_tiles] = total; ctl[kCtlItems] = n_items; ctl[kCtlTicket] = 0; }
    }
}

// --------------------------------------------------------------------------------------------
// TILE path, step 3: tile lists of piece descriptors (order inside a tile is irrelevant: every
// piece owns its slot).  A descriptor is 8 ints:
//   roi, slot, rows = r0 | nrows<<8 | wy_off<<16, cols[3] = c0 | ncols<<8, pieces of the RoI, its first slot
// (r0/c0 tile-local start, wy_off offset of the first row inside the RoI window), followed in two side
// arrays by the piece's slices of the separable weights: list_wx[pos][3][32], list_wy[pos][kTH] (valid entries only).
// --------------------------------------------------------------------------------------------
#ifndef RR_FILL_MINB
#define RR_FILL_MINB 6           // CTAs per SM the register budget is cut for: the kernel is a chain of latencies (measured 20.5 -> 18.4 us)
#endif
__global__ void __launch_bounds__(256, RR_FILL_MINB)
roi_fill_kernel(const RoiPrep* __restrict__ prep, const int* __restrict__ slot,
                const float* __restrict__ wx, const float4* __restrict__ wy4,
                const int* __restrict__ n_rois_dev, int n_cap, TileDims td,
                const int* __restrict__ tile_off, int* __restrict__ tile_fill,
                int4* __restrict__ list, float* __restrict__ list_wx, float4* __restrict__ list_wy) {
    RR_PDL_PROLOGUE();
    const int lane = lane_id();
    const int n = blockIdx.x * (blockDim.x >> 5) + warp_id();      // one warp per RoI
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    if (n >= live) return;
    const RoiPrep rp = prep[n];
    const int sb = slot[n];
    if (rp.flags != kFlagTile || sb < 0) return;
    const float* wxn = wx + (size_t)n * (RR_POOL * kMaxWinT);
    const float4* wyn = wy4 + (size_t)n * kMaxWinT;
    // list positions of all pieces of the RoI at once: lane p takes piece p (one atomic latency instead of one per piece)
    const int m = rp.ntx * rp.nty;                                 // <= kMaxPieces
    int my_pos = 0;
    if (lane < m) {
        const int t = rp.img * td.tiles_per_img + (rp.ty0 + lane / rp.ntx) * td.ntx + rp.tx0 + lane % rp.ntx;
        my_pos = tile_off[t] + atomicAdd(tile_fill + t, 1);
    }
    // Only the valid rows / columns of a slice are written: every consumer stops at nrows / ncols (or masks by them), so the
    // entries past them are never used - and the kernel is a chain of dependent latencies, not of bytes: two pieces per
    // iteration keep the loads of one under the stores of the other.
#pragma unroll 2
    for (int pc = 0; pc < m; ++pc) {
        const int j = pc / rp.ntx, i = pc - j * rp.ntx;
        const int py0 = (rp.ty0 + j) * kTH;
        const int r0 = max(rp.y_lo, py0), r1 = min(rp.y_lo + rp.ny - 1, py0 + kTH - 1);
        const int nrows = r1 - r0 + 1, wyo = r0 - rp.y_lo;
        const int rows = (r0 - py0) | (nrows << 8) | (wyo << 16);
        const int px0 = (rp.tx0 + i) * kTW;
        int cols[RR_POOL], ncol[RR_POOL], wxo[RR_POOL];
#pragma unroll
        for (int p = 0; p < RR_POOL; ++p) {
            const int c0 = max(rp.x_lo + rp.cx_lo[p], px0), c1 = min(rp.x_lo + rp.cx_hi[p], px0 + kTW - 1);
            ncol[p] = max(c1 - c0 + 1, 0);
            wxo[p] = c0 - rp.x_lo;
            cols[p] = ncol[p] > 0 ? ((c0 - px0) | (ncol[p] << 8)) : 0;
        }
        const int pos = __shfl_sync(0xffffffffu, my_pos, pc);
        if (lane == 0) {
            list[2 * pos] = make_int4(n, sb + pc, rows, cols[0]);
            list[2 * pos + 1] = make_int4(cols[1], cols[2], m, sb);       // .z pieces of the RoI, .w its first slot
        }
        // the piece's slices of the separable weights, in list order
        if (lane < nrows) list_wy[(size_t)pos * kTH + lane] = wyn[wyo + lane];
#pragma unroll
        for (int p = 0; p < RR_POOL; ++p)
            if (lane < ncol[p]) list_wx[((size_t)pos * RR_POOL + p) * kTW + lane] = wxn[p * kMaxWinT + wxo[p] + lane];
    }
}

// --------------------------------------------------------------------------------------------
// TILE path, step 4: persistent CTAs pull (work item, channel group) tickets.
// --------------------------------------------------------------------------------------------
template <int NJ>
__device__ __forceinline__ void unit_rows(const float* __restrict__ fr, const float4* __restrict__ s_wy, int nrows,
                                          const float (&w)[8], float& a0, float& a1, float& a2) {
#pragma unroll 2
    for (int y = 0; y < nrows; ++y, fr += kTW) {
        float s = w[0] * fr[0];
#pragma unroll
        for (int j = 1; j < NJ; ++j) s = fmaf(w[j], fr[j], s);
        const float4 wy = s_wy[y];                 // warp-uniform address: one broadcast wavefront
        a0 = fmaf(wy.x, s, a0);
        a1 = fmaf(wy.y, s, a1);
        a2 = fmaf(wy.z, s, a2);
    }
}

// one channel of the tile: kTH coalesced 128-byte rows -> shared memory, ReLU on the way in.
// p points at (row py0, clamped column) of the channel plane; rows past the map are clamped too
// (those tile cells are never read: piece ranges are clipped to the map), so no load is predicated.
__device__ __forceinline__ float ldg_row(const float* p, unsigned row_bytes, unsigned y) {
    unsigned long long a;                          // one IMAD.WIDE.U32 per row address
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(row_bytes), "r"(y), "l"(p));
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(a));
    return v;
}

template <bool kFull>
__device__ __forceinline__ void stage_channels(const float* __restrict__ p, size_t plane, float* __restrict__ d,
                                               int W, int rows_valid, int relu) {
    constexpr int kPer = kTC / kTileWarps;         // channels per warp: all kPer*kTH row loads are in flight together
    const unsigned row_bytes = (unsigned)W * 4u;
    float v[kPer][kTH];
#pragma unroll
    for (int cc = 0; cc < kPer; ++cc)
#pragma unroll
        for (int y = 0; y < kTH; ++y)
            v[cc][y] = ldg_row(p + cc * plane, row_bytes, kFull ? (unsigned)y : (unsigned)min(y, rows_valid - 1));
#pragma unroll
    for (int cc = 0; cc < kPer; ++cc)
#pragma unroll
        for (int y = 0; y < kTH; ++y) d[cc * kChStride + y * kTW] = relu ? fmaxf(v[cc][y], 0.f) : v[cc][y];
}

#ifdef RR_TILE_TRACE         // tools/tile_trace.py: per-CTA phase times of roi_tile_kernel (never defined in the shipped build)
__device__ unsigned long long g_tile_trace[512 * 16];
__device__ __forceinline__ long long tile_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TILE_LAP(acc) do { const long long n_ = tile_now(); (acc) += n_ - lap_; lap_ = n_; } while (0)
#else
#define TILE_LAP(acc) do { } while (0)
#endif

// Persistent CTAs, two per SM (while one stages its next tile the other one computes).  Per ticket:
// stage tile + piece tables -> barrier -> every warp evaluates its (piece, bin column) units -> barrier.
// The next ticket is fetched by one thread during the compute phase.
__global__ void __launch_bounds__(kTileThreads, 2)
roi_tile_kernel(const float* __restrict__ feat, const int4* __restrict__ list,
                const float* __restrict__ list_wx, const float4* __restrict__ list_wy,
                const int4* __restrict__ items, const int* __restrict__ tile_off,
                const int* __restrict__ tile_fill, int* __restrict__ ctl,
                int C, int H, int W, int relu, TileDims td, float* __restrict__ partial) {
    RR_PDL_PROLOGUE();
    extern __shared__ float s_tile[];              // [kTC][kChStride] | int4 s_desc[kChunk][2] | float4 s_wy[kChunk][kTH]
    __shared__ int s_work[2][8];                   // ticket decode: g, list start, pieces (-1: done), px0, py0, img
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int4* s_desc = reinterpret_cast<int4*>(s_tile + kTileTileFloats);
    float4* s_wy = reinterpret_cast<float4*>(s_desc + 2 * kChunk);

    auto fetch = [&](int* wk) {                    // one thread: next ticket -> decoded work item
        const int ngroups = C / kTC;
        const int work = atomicAdd(ctl + kCtlTicket, 1);
        int n_pieces = -1;
        const int n_items = ctl[kCtlItems];
        if (work < n_items * ngroups) {
            const int4 it = items[work / ngroups];       // (group-major ticket order, x-adjacent tiles back to back: 3 % slower)
            const int t = it.x;
            const int img = t / td.tiles_per_img, trem = t - img * td.tiles_per_img;
            const int ty = trem / td.ntx, tx = trem - ty * td.ntx;
            n_pieces = max(min(it.z, tile_off[t] + tile_fill[t] - it.y), 0);
            wk[0] = work % ngroups; wk[1] = it.y;
            wk[3] = tx * kTW; wk[4] = ty * kTH; wk[5] = img;
        }
        wk[2] = n_pieces;
    };
    constexpr int kFetchTid = kTileThreads - 32;   // lane 0 of the last warp (it gets the fewest units)
    if (tid == kFetchTid) fetch(s_work[0]);
    __syncthreads();
#ifdef RR_TILE_TRACE
    long long lap_ = tile_now(), t_stage = 0, t_bar1 = 0, t_units = 0, t_bar2 = 0, n_tickets = 0, n_units_mine = 0;
    const long long t_begin = lap_;
#endif
    for (int buf = 0;; buf ^= 1) {
        const int n_pieces = s_work[buf][2];
        if (n_pieces < 0) break;
        const int g = s_work[buf][0], list0 = s_work[buf][1];
        const int px0 = s_work[buf][3], py0 = s_work[buf][4], img = s_work[buf][5];

        // ---- piece descriptors and row weights of this item -> shared (contiguous in list order) ----
        if (tid < 2 * n_pieces) s_desc[tid] = __ldg(list + 2 * (size_t)list0 + tid);
        for (int i = tid; i < n_pieces * kTH; i += kTileThreads) s_wy[i] = __ldg(list_wy + (size_t)list0 * kTH + i);
        // ---- stage the tile: each warp brings in kTC/kTileWarps channels, ReLU on the way in ----
        {
            const int gx = min(px0 + lane, W - 1);
            const int rows_valid = min(kTH, H - py0);
            const int c = warp * (kTC / kTileWarps);
            const size_t plane = (size_t)H * W;
            const float* p = feat + (((size_t)img * C + (size_t)g * kTC + c) * H + py0) * W + gx;
            float* d = s_tile + c * kChStride + lane;
            if (rows_valid == kTH) stage_channels<true>(p, plane, d, W, rows_valid, relu);
            else stage_channels<false>(p, plane, d, W, rows_valid, relu);
        }
        TILE_LAP(t_stage);
        __syncthreads();
        TILE_LAP(t_bar1);
        if (tid == kFetchTid) fetch(s_work[buf ^ 1]);   // next ticket, overlapped with the compute below

        // ---- units: (piece, bin column pw), lane = channel ----
        const int n_units = n_pieces * RR_POOL;
        const float* wxp = list_wx + (size_t)list0 * RR_POOL * kTW + lane;     // unit u: wxp[u * kTW]
        float wxv = warp < n_units ? __ldg(wxp + warp * kTW) : 0.f;
        for (int u = warp; u < n_units; u += kTileWarps) {
            const int piece = u / RR_POOL, pw = u - piece * RR_POOL;
            const int4 d0 = s_desc[2 * piece], d1 = s_desc[2 * piece + 1];
            const int cols = pw == 0 ? d0.w : (pw == 1 ? d1.x : d1.y);
            const int r0 = d0.z & 0xff, nrows = (d0.z >> 8) & 0xff;
            const int c0 = cols & 0xff, ncols = (cols >> 8) & 0xff;
            const float wcur = wxv;
            if (u + kTileWarps < n_units) wxv = __ldg(wxp + (u + kTileWarps) * kTW);   // prefetch the next unit's
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            const float* fb = s_tile + lane * kChStride + r0 * kTW + c0;
            const float4* wyp = s_wy + piece * kTH;
            for (int cc = 0; cc < ncols; cc += 8) {
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = __shfl_sync(0xffffffffu, wcur, (cc + j) & 31);
                const float* fr = fb + cc;
                switch (min(ncols - cc, 8)) {
                    case 1: unit_rows<1>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 2: unit_rows<2>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 3: unit_rows<3>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 4: unit_rows<4>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 5: unit_rows<5>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 6: unit_rows<6>(fr, wyp, nrows, w, a0, a1, a2); break;
                    case 7: unit_rows<7>(fr, wyp, nrows, w, a0, a1, a2); break;
                    default: unit_rows<8>(fr, wyp, nrows, w, a0, a1, a2); break;
                }
            }
            float* po = partial + ((size_t)d0.y * (RR_POOL * RR_POOL) + pw) * C + g * kTC + lane;
            po[0] = a0;
            po[(size_t)RR_POOL * C] = a1;
            po[(size_t)2 * RR_POOL * C] = a2;
#ifdef RR_TILE_TRACE
            ++n_units_mine;
#endif
        }
        TILE_LAP(t_units);
        __syncthreads();                           // shared tables free again; s_work[buf ^ 1] is visible
        TILE_LAP(t_bar2);
#ifdef RR_TILE_TRACE
        ++n_tickets;
#endif
    }
#ifdef RR_TILE_TRACE
    if (lane == 0 && (warp == 0 || warp == kTileWarps - 1) && blockIdx.x < 512) {
        unsigned long long* o = g_tile_trace + blockIdx.x * 16 + (warp ? 8 : 0);
        o[0] = t_stage; o[1] = t_bar1; o[2] = t_units; o[3] = t_bar2; o[4] = n_tickets; o[5] = n_units_mine; o[6] = tile_now() - t_begin;
    }
#endif
}

// --------------------------------------------------------------------------------------------
// TILE path, step 4 (default when W % 4 == 0): the same tickets, but the tile arrives by TMA.
//
// One CTA per SM, 23 warps.  The last warp is the producer: it pulls tickets, issues ONE
// cp.async.bulk.tensor (box 32 x | 32 c | 24 y of a 4-D tensor map whose dimensions are ordered
// x, c, y, image) per ticket into one of two 96 KB tile buffers and copies the ticket's piece
// tables next to it; `full` / `empty` mbarriers hand the buffers over, so the load of ticket i+1
// runs under the evaluation of ticket i and no thread ever issues a tile load or store.
// Layout in shared memory (SWIZZLE_128B): row (y, c) is 128 bytes = the 32 pixels of one tile row of
// channel c, rows in (y, c) order, so the 8-row swizzle period runs over c: 16-byte chunk q of row
// (y, c) sits at chunk q ^ (c & 7).  lane = channel: a quarter warp reading the SAME logical chunk
// of 8 consecutive channels touches 8 different physical chunks = all 32 banks once: LDS.128 without
// conflicts, 4 pixels per load.  The price is chunk granularity: a unit reads the aligned chunks
// covering its columns (weights outside the unit are zero) and the ReLU moves into the evaluation.
// The 22 consumer warps take (piece, bin column) units from a shared counter and move on to the
// next ticket on their own (no CTA-wide barrier).
// --------------------------------------------------------------------------------------------
#ifndef RR_T2_MERGED
#define RR_T2_MERGED 0       // 1: a unit is a whole piece (three bin columns in one pass).  Measured on the B200 (config 2):
#endif                       // 0.581 ms against 0.458 ms for (piece, bin column) units - see unit_rows_m
// RR_T2_TMEM 1 (NOT the default - measured slower): the first kT2TmRows rows of every tile are ALSO copied into tensor
// memory (64 x tcgen05.cp.32x128b.warpx4 straight from the swizzled tile: 32 channels -> 32 lanes, replicated into the four
// lane quarters so that any warp can read them; 512 columns = 2 tile buffers x 8 rows x 32 pixels, no MMA is ever issued) and
// the units take those rows with tcgen05.ld.32x32b.x4 - the same four pixels of the lane's channel an LDS.128 returns -
// instead of through the load/store unit, whose shared-memory wavefronts bound this kernel.  tools/tmem_tile_probe.cu on
// the B200: the layout works as designed (0 mismatches) and the two read paths do add up (LDS.128 alone 128 B/clk/SM,
// tcgen05.ld alone up to 222, both at once 120 + 134).  But (1) a tcgen05.cp costs ~75 cycles whatever its shape (32x128b
// 512 bytes, 128x256b 4 KB: ~130), i.e. 5 200 cycles per ticket for 8 rows, and while copies run LDS.128 drops to ~50
// B/clk/SM - the copies take more from the LSU than the tensor-memory reads give back; (2) LDTM wants its address in a
// uniform register: every base address costs WARPSYNC + R2UR, rows must be compile-time offsets (pairs of rows per block).
// Kernel time at config 2, bit-identical results: 0.463 ms without, 0.525 with; copies but no tensor-memory reads 0.497,
// reads but no copies 0.489, neither (just the extra code and barriers) 0.477.  Kept for the record.
#ifndef RR_T2_TMEM
#define RR_T2_TMEM 0
#endif
#ifndef RR_T2_THREADS
#define RR_T2_THREADS (RR_T2_MERGED ? 640 : (RR_T2_TMEM ? 768 : 736))    // consumer warps + the producer (+ the tensor-memory copier) (merged units: 19 warps, 102 registers).  Measured (RoIAlign stage, ms): 416: 0.557, 480: 0.538,
#endif                       // 544: 0.524, 608: 0.518, 672: 0.513, 736: 0.512, 800 (spills): 0.534, 1024 (64 registers): 0.533
// RR_T2_MERGED 1: a unit is a whole piece (three bin columns in one pass); 0: (piece, bin column) units
#ifndef RR_T2_UNROLL
#define RR_T2_UNROLL 2       // rows per iteration of the unit loop (3 and 4 were not faster at 544 - 672 threads)
#endif
constexpr int kT2Threads = RR_T2_THREADS;
constexpr int kT2Unroll = RR_T2_UNROLL;
constexpr int kT2Consumers = kT2Threads / 32 - 1 - (RR_T2_TMEM ? 1 : 0);
#ifndef RR_T2_TM_ROWS
#define RR_T2_TM_ROWS 8
#endif
#ifndef RR_T2_TM_MIN_PIECES
#define RR_T2_TM_MIN_PIECES 0            // tickets with fewer pieces are evaluated from shared memory only (no copy)
#endif
constexpr int kT2TmRows = RR_T2_TM_ROWS; // tile rows mirrored in tensor memory (8 x 32 pixels = 256 columns per tile buffer)
constexpr int kT2TmCols = 512;
static_assert(2 * kT2TmRows * kTW <= kT2TmCols && kT2TmRows <= kTH, "two tile buffers' rows fit the 512 columns");
constexpr int kT2Passes = 4;             // unit size classes, handed out largest first
constexpr int kT2TileBytes = kTC * kTH * kTW * (int)sizeof(float);            // 98304 = one TMA box
constexpr int kT2TableBytes = kChunk * 2 * 16 + kChunk * kTH * 16;            // descriptors + row weights
constexpr int kT2Smem = 2 * kT2TileBytes + 2 * kT2TableBytes + 1024;          // + slack to align the tiles to 1024 bytes

__device__ __forceinline__ uint32_t t2_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t2_bar_wait_warp(unsigned long long* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        uint32_t ok;
        for (;;) {
            // suspend-time hint: the thread sleeps in hardware until the phase completes (or ~1 us passes) instead of
            // polling - a spinning warp takes issue slots from the evaluating ones
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(t2_saddr(bar)), "r"(parity), "r"(1000u) : "memory");
            if (ok) break;
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void t2_bar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(t2_saddr(bar)) : "memory");
}

// packed fp32 FMA (SASS FFMA2): the FMA rate of two FFMAs for one issue slot
__device__ __forceinline__ unsigned long long t2_pack(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ unsigned long long t2_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// rows of one unit over NQ aligned 16-byte chunks: s(y) = sum_x w[x] relu(f[y][x]), a_ph += wy[ph][y] s(y)
template <int NQ, bool kRelu>
__device__ __forceinline__ void unit_rows_q(const float* __restrict__ rowp, const int (&off)[3],
                                            const float4* __restrict__ s_wy, int nrows, const ulonglong2 (&w)[3],
                                            float& a0, float& a1, float& a2) {
    const float* p[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) p[j] = rowp + off[j];
#pragma unroll kT2Unroll
    for (int y = 0; y < nrows; ++y) {
        unsigned long long s01 = 0ull, s23 = 0ull;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            float4 v = *reinterpret_cast<const float4*>(p[j]);
            p[j] += kTC * kTW;
            if (kRelu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            s01 = t2_fma2(w[j].x, t2_pack(v.x, v.y), s01);
            s23 = t2_fma2(w[j].y, t2_pack(v.z, v.w), s23);
        }
        float sx, sy, sz, sw;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(s01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(sz), "=f"(sw) : "l"(s23));
        const float s = (sx + sz) + (sy + sw);
        const float4 wy = s_wy[y];                 // warp-uniform address: broadcast
        a0 = fmaf(wy.x, s, a0);
        a1 = fmaf(wy.y, s, a1);
        a2 = fmaf(wy.z, s, a2);
    }
}

// ---- tensor-memory side of a unit (RR_T2_TMEM) ----
__device__ __forceinline__ void t2_ldtm4(uint32_t taddr, uint32_t (&r)[4]) {      // lane <- its TMEM lane, 4 consecutive columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void t2_ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the registers of a tcgen05.ld must not be read before the wait: tie their uses to a point after it
__device__ __forceinline__ void t2_pin(uint32_t (&r)[4]) { asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])); }
// shared-memory matrix descriptor of 32 rows x 16 bytes inside a K-major SWIZZLE_128B tile (8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t t2_sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int NQ, bool kRelu>
__device__ __forceinline__ void unit_row_t(const uint32_t (&r)[3][4], const float4 wy, const ulonglong2 (&w)[3],
                                           float& a0, float& a1, float& a2) {
    unsigned long long s01 = 0ull, s23 = 0ull;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        float4 v = make_float4(__uint_as_float(r[j][0]), __uint_as_float(r[j][1]), __uint_as_float(r[j][2]), __uint_as_float(r[j][3]));
        if (kRelu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        s01 = t2_fma2(w[j].x, t2_pack(v.x, v.y), s01);
        s23 = t2_fma2(w[j].y, t2_pack(v.z, v.w), s23);
    }
    float sx, sy, sz, sw;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(s01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(sz), "=f"(sw) : "l"(s23));
    const float s = (sx + sz) + (sy + sw);
    a0 = fmaf(wy.x, s, a0);
    a1 = fmaf(wy.y, s, a1);
    a2 = fmaf(wy.z, s, a2);
}

// rows of one unit that live in tensor memory: the arithmetic of unit_rows_q, the pixels from tcgen05.ld.  ta = the
// warp's lane quarter | column of (the unit's first row, its first chunk of this group).  LDTM takes its address from a
// UNIFORM register plus an immediate: with a dynamic row / chunk offset every load costs a WARPSYNC + R2UR of its own
// (measured: 0.70 ms for the kernel instead of 0.46), so the (at most kT2TmRows) rows and the chunks of the group are
// compile-time offsets from ONE base.  tcgen05.wait::ld is free (the loads sit on the ordinary scoreboard).
template <int NQ, bool kRelu>
__device__ __forceinline__ void unit_rows_t(const uint32_t ta, const float4* __restrict__ s_wy, const int npairs,
                                            const ulonglong2 (&w)[3], float& a0, float& a1, float& a2) {
#pragma unroll
    for (int p = 0; p < kT2TmRows / 2; ++p) {              // rows come in PAIRS (an odd last row is left to the shared-memory loop):
        if (p < npairs) {                                  // one straight-line block and one base-address move per pair
            uint32_t ra[3][4], rb[3][4];
#pragma unroll
            for (int j = 0; j < NQ; ++j) t2_ldtm4(ta + (uint32_t)(2 * p * kTW + 4 * j), ra[j]);
#pragma unroll
            for (int j = 0; j < NQ; ++j) t2_ldtm4(ta + (uint32_t)((2 * p + 1) * kTW + 4 * j), rb[j]);
            const float4 wy0 = s_wy[2 * p], wy1 = s_wy[2 * p + 1];
            t2_ldtm_wait();
#pragma unroll
            for (int j = 0; j < NQ; ++j) { t2_pin(ra[j]); t2_pin(rb[j]); }
            unit_row_t<NQ, kRelu>(ra, wy0, w, a0, a1, a2);
            unit_row_t<NQ, kRelu>(rb, wy1, w, a0, a1, a2);
        }
    }
}

// MERGED units (-DRR_T2_MERGED=1, NOT the default): a unit is a whole piece - all three bin columns in one pass over the piece's
// aligned chunks.  Every 16-byte chunk of the tile is then read ONCE per piece (the bin columns of a narrow piece share
// chunks; per piece row 3.7 chunk reads instead of 5.0) and the row weights once instead of three times: -35 % shared-
// memory wavefronts per piece row, the kernel's limiter.  Price: every chunk is multiplied into all three bin columns
// (weights outside a bin's range are zero) and the weights of a pass take 6 registers per chunk.
// Result: slower (0.581 ms against 0.458 ms).  A third as many units per ticket (~25 pieces for 19 warps) wrecks the
// balance inside a ticket, 96 registers cost three warps, and the kernel is not bound by wavefronts alone: taking the
// ReLU out of it (-27 % instructions, feat_is_relu) only gains 3 %.  It is a latency chain per unit (LDS -> FMNMX ->
// FFMA2 -> row reduction) on 22 warps that no single pipe saturates (LSU 71 %, issue 62 %).  Kept for the record.
template <int NQ, bool kRelu>
__device__ __forceinline__ void unit_rows_m(const float* __restrict__ rowp, const int (&off)[3],
                                            const float4* __restrict__ s_wy, int nrows, const ulonglong2 (&w)[3][3],
                                            float (&a)[3][3]) {
    const float* p[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) p[j] = rowp + off[j];
#pragma unroll kT2Unroll
    for (int y = 0; y < nrows; ++y) {
        unsigned long long s01[3] = {0ull, 0ull, 0ull}, s23[3] = {0ull, 0ull, 0ull};
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            float4 v = *reinterpret_cast<const float4*>(p[j]);
            p[j] += kTC * kTW;
            if (kRelu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            const unsigned long long v01 = t2_pack(v.x, v.y), v23 = t2_pack(v.z, v.w);
#pragma unroll
            for (int pw = 0; pw < 3; ++pw) {
                s01[pw] = t2_fma2(w[pw][j].x, v01, s01[pw]);
                s23[pw] = t2_fma2(w[pw][j].y, v23, s23[pw]);
            }
        }
        const float4 wy = s_wy[y];                 // warp-uniform address: broadcast, once for the three bin columns
#pragma unroll
        for (int pw = 0; pw < 3; ++pw) {
            float sx, sy, sz, sw;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(s01[pw]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sz), "=f"(sw) : "l"(s23[pw]));
            const float sv = (sx + sz) + (sy + sw);
            a[0][pw] = fmaf(wy.x, sv, a[0][pw]);
            a[1][pw] = fmaf(wy.y, sv, a[1][pw]);
            a[2][pw] = fmaf(wy.z, sv, a[2][pw]);
        }
    }
}

template <bool kRelu>
__global__ void __launch_bounds__(kT2Threads, 1)
roi_tile_tma_kernel(const __grid_constant__ CUtensorMap tmap, const int4* __restrict__ list,
                    const float* __restrict__ list_wx, const float4* __restrict__ list_wy,
                    const int4* __restrict__ items, const int* __restrict__ tile_off,
                    const int* __restrict__ tile_fill, int* __restrict__ ctl,
                    int C, TileDims td, float* __restrict__ partial,
                    int* __restrict__ arrived, const float* __restrict__ cnt_arr, float* __restrict__ combined,
                    float* __restrict__ atomic_rows) {
    RR_PDL_PROLOGUE();
    extern __shared__ unsigned char s_raw[];
    __shared__ int s_work[2][4];                   // g, list start, pieces (-1: no more tickets)
    __shared__ int s_next[2];                      // unit counter of the ticket in buffer b
    __shared__ unsigned char s_order[2][kChunk * RR_POOL];     // the ticket's units, largest size class first
    __shared__ unsigned long long s_full[2], s_empty[2];
    __shared__ __align__(16) float s_wal[kT2Consumers][kTW];   // per consumer warp: the current unit's column weights
    // In-kernel combine (combined != nullptr; rr_set_option(RR_OPT_COMBINE_IN_TILE_KERNEL)).  A RoI's pieces are finished
    // by different CTAs at different times.  When a ticket is over (its `empty` barrier), the producer makes the ticket's
    // slot stores visible (one fence), counts the ticket's pieces into arrived[roi][group] and, for every RoI whose LAST
    // piece this was, queues a combine job (first slot, pieces, roi, group) with the ticket that goes into the buffer
    // next; the consumers run the jobs after the ticket's units: slots summed in slot order, scaled by 1 / count,
    // written as the RoI's [9][C] row - what roi_combine_kernel computes, bit for bit.  Nobody waits for another CTA and
    // the order of summation is fixed: deterministic.  Measured at config 2: the head drops from 0.167 to 0.135 ms (one
    // row per RoI instead of 3.1 slots), but this kernel goes from 0.448 to 0.558 ms - the jobs re-read the 282 MB of
    // slots in short latency-bound bursts - so the default stays: slots to the head.
    __shared__ int4 s_jobs[2][kChunk];
    __shared__ int s_njobs[2], s_nextjob[2];
    __shared__ unsigned long long s_jobsready[2];            // the producer has written the jobs that ride with buffer b's ticket
#if RR_T2_TMEM
    __shared__ unsigned long long s_tfull[2];                // rows 0..kT2TmRows-1 of buffer b's tile are in tensor memory (tcgen05.commit)
    __shared__ unsigned long long s_done;                    // every consumer warp is past its last tensor-memory read
    __shared__ uint32_t s_tmem;
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* base = s_raw + ((1024u - (t2_saddr(s_raw) & 1023u)) & 1023u);
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_saddr(&s_full[b])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_saddr(&s_empty[b])), "r"(kT2Consumers) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_saddr(&s_jobsready[b])), "r"(1) : "memory");
#if RR_T2_TMEM
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_saddr(&s_tfull[b])), "r"(1) : "memory");
#endif
        }
#if RR_T2_TMEM
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_saddr(&s_done)), "r"(kT2Consumers) : "memory");
#endif
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#if RR_T2_TMEM
    if (warp == kT2Consumers + 1) {                          // the copier warp owns the allocation (all 512 columns of the SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t2_saddr(&s_tmem)), "n"(kT2TmCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
    __syncthreads();
#if RR_T2_TMEM
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp == kT2Consumers + 1) {
        // ------------------------------ tensor-memory copier ------------------------------
        // As soon as a tile has landed (`full`): 64 x tcgen05.cp.32x128b.warpx4 - row y, 16-byte chunk q of the 32
        // channels -> columns 32 y + 4 q .. + 3 of the buffer's 256-column region, all four lane quarters - and one
        // commit onto `tfull`.  The region's previous reader (ticket i - 2) is done: the producer only refilled the
        // buffer after its `empty` phase, which every consumer reaches after its last tcgen05.wait::ld.
        if (lane == 0) {
            for (int i = 0;; ++i) {
                const int b = i & 1;
                uint32_t ok;
                for (;;) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(t2_saddr(&s_full[b])), "r"((uint32_t)((i >> 1) & 1)), "r"(1000u) : "memory");
                    if (ok) break;
                }
                const int n_pieces = *reinterpret_cast<volatile int*>(&s_work[b][2]);
                if (n_pieces < 0) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (n_pieces > 0 && n_pieces >= RR_T2_TM_MIN_PIECES) {
                    const uint32_t src = t2_saddr(base + b * kT2TileBytes), dst = tmem + (uint32_t)(b * kT2TmRows * kTW);
#pragma unroll 1
                    for (int y = 0; y < kT2TmRows; ++y) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;"
                                         ::"r"(dst + (uint32_t)(y * kTW + 4 * q)), "l"(t2_sw128_desc(src + (uint32_t)(y * kTC * kTW * 4 + 16 * q))) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t2_saddr(&s_tfull[b])) : "memory");
            }
            uint32_t ok;
            for (;;) {                                       // all consumers are through: the columns can go back
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(t2_saddr(&s_done)), "r"(0u), "r"(1000u) : "memory");
                if (ok) break;
            }
        }
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kT2TmCols) : "memory");
        return;
    }
#endif

    if (warp == kT2Consumers) {
        // ------------------------------ producer warp ------------------------------
        // Tickets are pulled in blocks of `blk` consecutive ones = the same work item for blk channel groups: one
        // chain of dependent loads (ticket counter -> item -> list fill) per block instead of per ticket, and the
        // item's piece tables, already in a buffer from the block's earlier ticket, are not copied again.
        const int ngroups = C / kTC;
        const int blk = (ngroups & 3) == 0 ? 4 : ((ngroups & 1) == 0 ? 2 : 1);
        const int n_tickets = ctl[kCtlItems] * ngroups;
        int work0 = 0, g0 = 0, list0 = 0, n_pieces = -1, px0 = 0, py0 = 0, img = 0;
        auto fetch = [&]() {                       // lane 0
            work0 = atomicAdd(ctl + kCtlTicket, blk);
            n_pieces = -1; g0 = list0 = px0 = py0 = img = 0;
            if (work0 < n_tickets) {
                const int4 it = items[work0 / ngroups];
                const int t = it.x;
                img = t / td.tiles_per_img;
                const int trem = t - img * td.tiles_per_img;
                const int ty = trem / td.ntx;
                n_pieces = max(min(it.z, tile_off[t] + tile_fill[t] - it.y), 0);
                g0 = work0 % ngroups; list0 = it.y;
                px0 = (trem - ty * td.ntx) * kTW; py0 = ty * kTH;
            }
        };
        int pos = blk, tables_of[2] = {-1, -1};
        int prev_list0[2] = {0, 0}, prev_np[2] = {0, 0}, prev_g[2] = {0, 0};     // the ticket that used buffer b before
        int flush = 0;                             // combine: job-only pseudo tickets after the last real one (1, 2), then the end (3)
        for (int i = 0;; ++i) {
            const int b = i & 1;
            if (flush == 0 && pos == blk) {
                if (lane == 0) fetch();
                n_pieces = __shfl_sync(0xffffffffu, n_pieces, 0);
                list0 = __shfl_sync(0xffffffffu, list0, 0);
                work0 = __shfl_sync(0xffffffffu, work0, 0);
                g0 = __shfl_sync(0xffffffffu, g0, 0);
                pos = 0;
                if (n_pieces < 0 && combined) flush = 1;          // out of tickets: two job-only rounds carry the last tickets' jobs
            }
            const bool real = flush == 0;          // a ticket with a tile (n_pieces >= 0) or, without combine, the terminator (-1)
            if (flush > 0) { n_pieces = flush <= 2 ? 0 : -1; ++flush; }
            const int g = g0 + pos;
            if (real) ++pos;
            const bool new_tables = real && n_pieces > 0 && tables_of[b] != work0;
            if (real) tables_of[b] = work0;
            if (i >= 2) t2_bar_wait_warp(&s_empty[b], (uint32_t)(((i >> 1) - 1) & 1));
            unsigned char* tile = base + b * kT2TileBytes;
            int4* desc = reinterpret_cast<int4*>(base + 2 * kT2TileBytes + b * kT2TableBytes);
            float4* wy = reinterpret_cast<float4*>(desc + 2 * kChunk);
            if (lane == 0) {
                s_work[b][0] = g; s_work[b][1] = list0; s_work[b][2] = n_pieces; s_next[b] = 0;
                if (real && n_pieces >= 0) {
                    const uint32_t bar = t2_saddr(&s_full[b]);
                    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(kT2TileBytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                                 "[%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(t2_saddr(tile)), "l"(&tmap), "r"(px0), "r"(g * kTC), "r"(py0), "r"(img), "r"(bar) : "memory");
                }
            }
            if (new_tables) {                      // piece tables: 16-byte asynchronous copies, all in flight at once
                const int4* src_d = list + 2 * (size_t)list0;
                for (int j = lane; j < 2 * n_pieces; j += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(t2_saddr(desc + j)), "l"(src_d + j) : "memory");
                const float4* src_w = list_wy + (size_t)list0 * kTH;
                for (int j = lane; j < n_pieces * kTH; j += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(t2_saddr(wy + j)), "l"(src_w + j) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (new_tables) {                      // order the units by size class (rows x chunks), largest first
                const int n_units = RR_T2_MERGED ? n_pieces : n_pieces * RR_POOL;
                int cls[RR_POOL], cnt[kT2Passes];
#pragma unroll
                for (int k = 0; k < kT2Passes; ++k) cnt[k] = 0;
#pragma unroll
                for (int r = 0; r < RR_POOL; ++r) {
                    const int u = r * 32 + lane;
                    cls[r] = kT2Passes;
                    if (u < n_units) {
#if RR_T2_MERGED
                        const int4 d0 = desc[2 * u], d1 = desc[2 * u + 1];
                        int qlo = 8, qhi = -1;
                        const int colsv[RR_POOL] = {d0.w, d1.x, d1.y};
#pragma unroll
                        for (int pw = 0; pw < RR_POOL; ++pw) {
                            const int c0 = colsv[pw] & 0xff, nc = (colsv[pw] >> 8) & 0xff;
                            if (nc > 0) { qlo = min(qlo, c0 >> 2); qhi = max(qhi, (c0 + nc - 1) >> 2); }
                        }
                        const int size = ((d0.z >> 8) & 0xff) * max(qhi - qlo + 1, 0);
                        cls[r] = size >= 96 ? 0 : (size >= 48 ? 1 : (size >= 20 ? 2 : 3));
#else
                        const int piece = u / RR_POOL, pw = u - piece * RR_POOL;
                        const int4 d0 = desc[2 * piece], d1 = desc[2 * piece + 1];
                        const int cols = pw == 0 ? d0.w : (pw == 1 ? d1.x : d1.y);
                        const int c0 = cols & 0xff, ncols = (cols >> 8) & 0xff;
                        const int size = ((d0.z >> 8) & 0xff) * (ncols > 0 ? ((c0 + ncols - 1) >> 2) - (c0 >> 2) + 1 : 0);
                        cls[r] = size >= 48 ? 0 : (size >= 24 ? 1 : (size >= 10 ? 2 : 3));
#endif
                    }
#pragma unroll
                    for (int k = 0; k < kT2Passes; ++k) cnt[k] += __popc(__ballot_sync(0xffffffffu, cls[r] == k));
                }
                int off[kT2Passes];
                off[0] = 0;
#pragma unroll
                for (int k = 1; k < kT2Passes; ++k) off[k] = off[k - 1] + cnt[k - 1];
#pragma unroll
                for (int r = 0; r < RR_POOL; ++r)
#pragma unroll
                    for (int k = 0; k < kT2Passes; ++k) {
                        const unsigned m = __ballot_sync(0xffffffffu, cls[r] == k);
                        if (cls[r] == k) s_order[b][off[k] + __popc(m & ((1u << lane) - 1u))] = (unsigned char)(r * 32 + lane);
                        off[k] += __popc(m);
                    }
                __syncwarp();
            }
            if (lane == 0) t2_bar_arrive(&s_full[b]);              // the tile and its tables: the consumers may start on the units
            if (combined) {
                // The ticket that was in this buffer before (i - 2) is complete (its `empty` phase): make its slot stores
                // visible device-wide (ONE fence by this warp: release cumulativity over the consumers' stores, which the
                // mbarrier ordered before this point - the pattern of a grid-wide barrier), count its pieces in; a RoI whose
                // last piece this was becomes a combine job.  Off the consumers' path: they only need the jobs after the units.
                int njobs = 0;
                if (prev_np[b] > 0) {
                    __threadfence();
                    bool last = false;
                    int4 job = make_int4(0, 0, 0, 0);
                    if (lane < prev_np[b]) {
                        const int4 d0 = __ldg(list + 2 * (size_t)(prev_list0[b] + lane));
                        const int4 d1 = __ldg(list + 2 * (size_t)(prev_list0[b] + lane) + 1);
                        const int old = atomicAdd(arrived + (size_t)d0.x * ngroups + prev_g[b], 1);
                        last = old == d1.z - 1;
                        job = make_int4(d1.w, d1.z, d0.x, prev_g[b]);          // first slot, pieces, roi, channel group
                    }
                    const unsigned lm = __ballot_sync(0xffffffffu, last);
                    if (last) s_jobs[b][__popc(lm & ((1u << lane) - 1u))] = job;
                    njobs = __popc(lm);
                    __threadfence();                   // the other CTAs' slot stores (released by their counts) before our loads
                }
                if (lane == 0) { s_njobs[b] = njobs; s_nextjob[b] = 0; }
                prev_list0[b] = list0; prev_np[b] = (real && n_pieces > 0) ? n_pieces : 0; prev_g[b] = g;
                __syncwarp();
                if (lane == 0) t2_bar_arrive(&s_jobsready[b]);
            }
            if (n_pieces < 0) break;
        }
        return;
    }

    // ------------------------------ consumer warps ------------------------------
    const int k7 = lane & 7;
#if RR_T2_TMEM
    const uint32_t tq = tmem + ((uint32_t)(32 * (warp & 3)) << 16);       // a warp reaches the lane quarter warp % 4 (all four hold the same rows)
#endif
    for (int i = 0;; ++i) {
        const int b = i & 1;
        t2_bar_wait_warp(&s_full[b], (uint32_t)((i >> 1) & 1));
        const int n_pieces = s_work[b][2];
        if (n_pieces < 0) break;
#if RR_T2_TMEM
        bool tm_ready = false;                             // this warp has seen `tfull` of this ticket
        auto tm_wait = [&]() {
            t2_bar_wait_warp(&s_tfull[b], (uint32_t)((i >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tm_ready = true;
        };
#endif
        const int g = s_work[b][0], list0 = s_work[b][1];
        const float* tile = reinterpret_cast<const float*>(base + b * kT2TileBytes);
        const int4* s_desc = reinterpret_cast<const int4*>(base + 2 * kT2TileBytes + b * kT2TableBytes);
        const float4* s_wy = reinterpret_cast<const float4*>(s_desc + 2 * kChunk);
#if RR_T2_MERGED
        const int n_units = n_pieces;
        const float* wxp = list_wx + (size_t)list0 * RR_POOL * kTW + lane;     // piece u, bin column pw: wxp[(u * 3 + pw) * kTW]
        auto next_unit = [&]() {                           // pieces are handed out largest first (ordered by the producer)
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_next[b], 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            return u < n_units ? (int)s_order[b][u] : n_units;
        };
        int u = next_unit();
        float wxv[RR_POOL] = {0.f, 0.f, 0.f};
        if (u < n_units) {
#pragma unroll
            for (int pw = 0; pw < RR_POOL; ++pw) wxv[pw] = __ldg(wxp + (u * RR_POOL + pw) * kTW);
        }
        while (u < n_units) {
            const int piece = u;
            float wal[RR_POOL];
            const int4 d0 = s_desc[2 * piece], d1 = s_desc[2 * piece + 1];
            const int colsv[RR_POOL] = {d0.w, d1.x, d1.y};
            const int r0 = d0.z & 0xff, nrows = (d0.z >> 8) & 0xff;
            int qlo = 8, qhi = -1;
            // the three bin columns' weights re-indexed by tile column (0 outside the bin's range)
#pragma unroll
            for (int pw = 0; pw < RR_POOL; ++pw) {
                const int c0 = colsv[pw] & 0xff, nc = (colsv[pw] >> 8) & 0xff;
                wal[pw] = __shfl_sync(0xffffffffu, wxv[pw], (lane - c0) & 31);
                if (lane < c0 || lane >= c0 + nc) wal[pw] = 0.f;
                if (nc > 0) { qlo = min(qlo, c0 >> 2); qhi = max(qhi, (c0 + nc - 1) >> 2); }
            }
            u = next_unit();                               // next piece and its column weights, under this piece's rows
            if (u < n_units) {
#pragma unroll
                for (int pw = 0; pw < RR_POOL; ++pw) wxv[pw] = __ldg(wxp + (u * RR_POOL + pw) * kTW);
            }
            float a[3][3];
#pragma unroll
            for (int ph = 0; ph < 3; ++ph)
#pragma unroll
                for (int pw = 0; pw < 3; ++pw) a[ph][pw] = 0.f;
            const float* rowp = tile + (r0 * kTC + lane) * kTW;
            const float4* wyp = s_wy + piece * kTH;
            for (int q = qlo; q <= qhi; q += 3) {
                ulonglong2 w[3][3];
                int off[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int qq = (q + j) & 7;
                    off[j] = (qq ^ k7) << 2;
#pragma unroll
                    for (int pw = 0; pw < RR_POOL; ++pw) {         // a chunk's four weights of a bin column: warp-uniform after the shuffles
                        const float w0 = __shfl_sync(0xffffffffu, wal[pw], 4 * qq), w1 = __shfl_sync(0xffffffffu, wal[pw], 4 * qq + 1);
                        const float w2 = __shfl_sync(0xffffffffu, wal[pw], 4 * qq + 2), w3 = __shfl_sync(0xffffffffu, wal[pw], 4 * qq + 3);
                        w[pw][j].x = t2_pack(w0, w1);
                        w[pw][j].y = t2_pack(w2, w3);
                    }
                }
                switch (min(qhi - q + 1, 3)) {
                    case 1: unit_rows_m<1, kRelu>(rowp, off, wyp, nrows, w, a); break;
                    case 2: unit_rows_m<2, kRelu>(rowp, off, wyp, nrows, w, a); break;
                    default: unit_rows_m<3, kRelu>(rowp, off, wyp, nrows, w, a); break;
                }
            }
            float* po = partial + (size_t)d0.y * (RR_POOL * RR_POOL) * C + g * kTC + lane;
#pragma unroll
            for (int ph = 0; ph < 3; ++ph)
#pragma unroll
                for (int pw = 0; pw < 3; ++pw) po[(size_t)(ph * RR_POOL + pw) * C] = a[ph][pw];
        }
#else
        const int n_units = n_pieces * RR_POOL;
        const float* wxp = list_wx + (size_t)list0 * RR_POOL * kTW + lane;     // unit u: wxp[u * kTW]
        // Units are handed out largest first (the producer warp has ordered them by size class): the buffer can
        // only be refilled when its last unit is done, so the stragglers should be the small ones.
        auto next_unit = [&]() {
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_next[b], 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            return u < n_units ? (int)s_order[b][u] : n_units;
        };
        int u = next_unit();
        float wxv = u < n_units ? __ldg(wxp + u * kTW) : 0.f;
        while (u < n_units) {
            const int ucur = u;
            const float wcur = wxv;
            u = next_unit();                               // next unit and its column weights, under this unit's rows
            if (u < n_units) wxv = __ldg(wxp + u * kTW);
            const int piece = ucur / RR_POOL, pw = ucur - piece * RR_POOL;
            const int4 d0 = s_desc[2 * piece], d1 = s_desc[2 * piece + 1];
            const int cols = pw == 0 ? d0.w : (pw == 1 ? d1.x : d1.y);
            const int r0 = d0.z & 0xff, nrows = (d0.z >> 8) & 0xff;
            const int c0 = cols & 0xff, ncols = (cols >> 8) & 0xff;
            // weights re-indexed by tile column (0 outside the unit) -> the warp's scratch row: a chunk's four
            // weights come back as one broadcast 128-bit read
            float wal = __shfl_sync(0xffffffffu, wcur, (lane - c0) & 31);
            if (lane < c0 || lane >= c0 + ncols) wal = 0.f;
            __syncwarp();
            s_wal[warp][lane] = wal;
            __syncwarp();
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if (ncols > 0) {
                const float4* wyp = s_wy + piece * kTH;
                const ulonglong2* wq = reinterpret_cast<const ulonglong2*>(s_wal[warp]);
                const int q1 = (c0 + ncols - 1) >> 2;
#if RR_T2_TMEM
                // rows below kT2TmRows come from tensor memory, the rest from the shared-memory tile (same arithmetic,
                // same order: rows ascending inside a chunk group)
                const int nt = n_pieces >= RR_T2_TM_MIN_PIECES ? (max(min(nrows, kT2TmRows - r0), 0) & ~1) : 0;     // an even number of rows
                if (nt > 0 && !tm_ready) tm_wait();
                const float* rowp = tile + ((r0 + nt) * kTC + lane) * kTW;
                const uint32_t ta = tq + (uint32_t)(b * kT2TmRows * kTW + r0 * kTW);
#else
                const float* rowp = tile + (r0 * kTC + lane) * kTW;
#endif
                for (int q = c0 >> 2; q <= q1; q += 3) {
                    ulonglong2 w[3];
                    int off[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        off[j] = (((q + j) & 7) ^ k7) << 2;
                        w[j] = wq[(q + j) & 7];
                    }
#if RR_T2_TMEM
                    const uint32_t tg = ta + (uint32_t)(4 * q);      // the group's chunks q .. q + 2 never wrap (q + j <= q1 <= 7)
                    switch (min(q1 - q + 1, 3)) {
                        case 1:
                            if (nt > 0) unit_rows_t<1, kRelu>(tg, wyp, nt >> 1, w, a0, a1, a2);
                            unit_rows_q<1, kRelu>(rowp, off, wyp + nt, nrows - nt, w, a0, a1, a2);
                            break;
                        case 2:
                            if (nt > 0) unit_rows_t<2, kRelu>(tg, wyp, nt >> 1, w, a0, a1, a2);
                            unit_rows_q<2, kRelu>(rowp, off, wyp + nt, nrows - nt, w, a0, a1, a2);
                            break;
                        default:
                            if (nt > 0) unit_rows_t<3, kRelu>(tg, wyp, nt >> 1, w, a0, a1, a2);
                            unit_rows_q<3, kRelu>(rowp, off, wyp + nt, nrows - nt, w, a0, a1, a2);
                            break;
                    }
#else
                    switch (min(q1 - q + 1, 3)) {
                        case 1: unit_rows_q<1, kRelu>(rowp, off, wyp, nrows, w, a0, a1, a2); break;
                        case 2: unit_rows_q<2, kRelu>(rowp, off, wyp, nrows, w, a0, a1, a2); break;
                        default: unit_rows_q<3, kRelu>(rowp, off, wyp, nrows, w, a0, a1, a2); break;
                    }
#endif
                }
            }
            if (atomic_rows) {
                // RR_OPT_COMBINE_IN_TILE_KERNEL = 2: no slots - the unit ADDS its three bins into the RoI's own (zeroed) [9][C]
                // row with fire-and-forget float reductions.  The order of a RoI's pieces is whatever the hardware makes it:
                // bit-reproducible for RoIs of one or two pieces only (a + b is commutative), last-bit differences otherwise.
                float* po = atomic_rows + ((size_t)d0.x * (RR_POOL * RR_POOL) + pw) * C + g * kTC + lane;
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(po), "f"(a0) : "memory");
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(po + (size_t)RR_POOL * C), "f"(a1) : "memory");
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(po + (size_t)2 * RR_POOL * C), "f"(a2) : "memory");
            } else {
                float* po = partial + ((size_t)d0.y * (RR_POOL * RR_POOL) + pw) * C + g * kTC + lane;
                po[0] = a0;
                po[(size_t)RR_POOL * C] = a1;
                po[(size_t)2 * RR_POOL * C] = a2;
            }
        }
#endif
        if (combined) {
            t2_bar_wait_warp(&s_jobsready[b], (uint32_t)((i >> 1) & 1));
            const int njobs = s_njobs[b];
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&s_nextjob[b], 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= njobs) break;
                const int4 job = s_jobs[b][j];         // first slot, pieces, roi, channel group
                const float inv = 1.0f / __ldg(cnt_arr + job.z);
                constexpr int kBins = RR_POOL * RR_POOL;
                float acc[kBins];
#pragma unroll
                for (int q = 0; q < kBins; ++q) acc[q] = 0.f;
                const float* src = partial + (size_t)job.x * kBins * C + job.w * kTC + lane;
                for (int p = 0; p < job.y; ++p, src += (size_t)kBins * C) {      // slot order, like roi_combine_kernel
                    float v[kBins];
#pragma unroll
                    for (int q = 0; q < kBins; ++q) v[q] = __ldcg(src + (size_t)q * C);
#pragma unroll
                    for (int q = 0; q < kBins; ++q) acc[q] += v[q];
                }
                float* dst = combined + (size_t)job.z * kBins * C + job.w * kTC + lane;
#pragma unroll
                for (int q = 0; q < kBins; ++q) dst[(size_t)q * C] = acc[q] * inv;
            }
        }
#if RR_T2_TMEM
        // Every warp sees `tfull` before it releases the buffer, whether it read tensor memory or not: the copier is then
        // never more than one phase behind (parity waits stay unambiguous) and is done with the tile before the next TMA
        // may overwrite it.
        if (!tm_ready) tm_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
        __syncwarp();
        if (lane == 0) t2_bar_arrive(&s_empty[b]);
    }
#if RR_T2_TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) t2_bar_arrive(&s_done);
#endif
}

// The 4-D tensor map of the feature maps, dimensions ordered (x, c, y, image) so that a box lands in shared
// memory as [y][c][x].  cuTensorMapEncodeTiled is a host-only driver entry point, resolved through the runtime.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_tile_tmap(const float* feat, int B, int C, int H, int W, CUtensorMap* out) {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    if (!fn || (W & 3) || (reinterpret_cast<uintptr_t>(feat) & 15)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {kTW, kTC, kTH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(feat), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TILE path, step 5: out[n,c,bin] = (sum over the RoI's pieces, fixed order) / count.
__global__ void __launch_bounds__(256)
roi_combine_kernel(const RoiPrep* __restrict__ prep, const int* __restrict__ slot,
                   const int* __restrict__ n_rois_dev, int n_cap, int C,
                   const float* __restrict__ partial, float* __restrict__ out) {
    RR_PDL_PROLOGUE();
    extern __shared__ float s_out[];               // [C][9]
    const int n = blockIdx.x;
    const int live = n_rois_dev ? min(*n_rois_dev, n_cap) : n_cap;
    if (n >= live) return;
    const int sb = slot[n];
    if (sb < 0) return;                            // written by roi_direct_kernel
    const RoiPrep rp = prep[n];
    const int pieces = rp.flags == kFlagTile ? rp.ntx * rp.nty : 0;
    constexpr int kBins = RR_POOL * RR_POOL;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc[kBins];
#pragma unroll
        for (int q = 0; q < kBins; ++q) acc[q] = 0.f;
        for (int p = 0; p < pieces; ++p) {
            const float* src = partial + (size_t)(sb + p) * kBins * C + c;
#pragma unroll
            for (int q = 0; q < kBins; ++q) acc[q] += __ldg(src + (size_t)q * C);
        }
        const float inv = 1.0f / rp.count;          // the fused head applies the same sum * (1/count)
#pragma unroll
        for (int q = 0; q < kBins; ++q) s_out[c * kBins + q] = acc[q] * inv;
    }
    __syncthreads();
    float* o = out + (size_t)n * C * kBins;
    for (int i = threadIdx.x; i < C * kBins; i += blockDim.x) o[i] = s_out[i];
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
struct RoiWs {
    RoiPrep* prep; int* meta; int* slot; float* cnt; float* wx; float4* wy4;
    int* zeroed; size_t zeroed_bytes;              // tile_count | tile_fill | ctl | arrived  (one memset)
    int* tile_count; int* tile_fill; int* ctl;
    int* arrived;                                  // [n_cap][C / 32] pieces of a RoI finished per channel group (in-kernel combine)
    int* tile_off; int4* items; int* direct_list; int4* list; float* list_wx; float4* list_wy; float* partial;
    int n_tiles, slot_cap; TileDims td;
    size_t bytes;
};

static RoiWs carve_roi(void* ws, int n_cap, int B, int C, int H, int W) {
    RoiWs w;
    w.td.ntx = (W + kTW - 1) / kTW;
    w.td.nty = (H + kTH - 1) / kTH;
    w.td.tiles_per_img = w.td.ntx * w.td.nty;
    w.n_tiles = B * w.td.tiles_per_img;
    w.slot_cap = kSlotsPerRoi * n_cap + w.n_tiles;
    Carver cv(ws);
    w.prep = cv.take<RoiPrep>(n_cap);
    w.meta = cv.take<int>((size_t)n_cap + 4);
    w.slot = cv.take<int>((size_t)n_cap + 4);
    w.cnt = cv.take<float>((size_t)n_cap);
    w.wx = cv.take<float>((size_t)n_cap * RR_POOL * kMaxWinT);
    w.wy4 = cv.take<float4>((size_t)n_cap * kMaxWinT);
    const size_t n_arrived = (size_t)n_cap * (size_t)((C + kTC - 1) / kTC);
    w.zeroed = cv.take<int>((size_t)2 * w.n_tiles + kCtlWords + n_arrived);
    w.zeroed_bytes = ((size_t)2 * w.n_tiles + kCtlWords + n_arrived) * sizeof(int);
    w.tile_count = w.zeroed; w.tile_fill = w.zeroed + w.n_tiles; w.ctl = w.zeroed + 2 * w.n_tiles;
    w.arrived = w.ctl + kCtlWords;
    w.tile_off = cv.take<int>((size_t)w.n_tiles + 1);
    w.items = cv.take<int4>((size_t)w.n_tiles + (size_t)n_cap * kMaxPieces / kChunk + 1);
    w.direct_list = cv.take<int>(n_cap);
    w.list = cv.take<int4>((size_t)2 * n_cap * kMaxPieces);
    w.list_wx = cv.take<float>((size_t)n_cap * kMaxPieces * RR_POOL * kTW);
    w.list_wy = cv.take<float4>((size_t)n_cap * kMaxPieces * kTH);
    w.partial = cv.take<float>((size_t)w.slot_cap * RR_POOL * RR_POOL * C);
    w.bytes = cv.off;
    return w;
}

size_t roi_align_ws_bytes(int n_cap, int B, int C, int H, int W) {
    return carve_roi(nullptr, n_cap, B, C, H, W).bytes;
}

// views into the workspace for a consumer that sums the partial slots itself (the fused head)
void roi_align_ws_views(void* ws, int n_cap, int B, int C, int H, int W, const float** partial, const int** slot,
                        const int** pieces, const float** count) {
    RoiWs w = carve_roi(ws, n_cap, B, C, H, W);
    *partial = w.partial; *slot = w.slot; *pieces = w.meta; *count = w.cnt;
}

// combine == 0: leave the tile-path RoIs as partial slots (out only receives the direct-path RoIs)
// combine == 1: roi_combine_kernel materialises out [n,C,3,3]
// combine == 3: the TMA tile kernel's units add their bins straight into the RoI's zeroed [9][C] ROW in out with float
//               reductions (no slots; unscaled sums; *rows_mode = 2, or 0 when the call fell back to plain slots)
// combine == 2: the TMA tile kernel combines a RoI's slots itself as soon as its last piece is done and writes the RoI's
//               [9][C] ROW into out (same buffer, other layout; direct-path RoIs still get [C][3][3]); *rows_mode tells the
//               caller whether that happened (1) or the call fell back to plain slots (0: no TMA, load-staged tiles)
int roi_align_launch(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                     int B, int C, int H, int W, int relu, int algo, int combine, float* out, void* ws,
                     cudaStream_t st, int* rows_mode) {
    int rc = 0;
    if (rows_mode) *rows_mode = 0;
    RoiWs w = carve_roi(ws, n_cap, B, C, H, W);
    RR_CUDA(cudaMemsetAsync(w.zeroed, 0, w.zeroed_bytes, st), rc);
    if (rc) return rc;
    const int force_direct = algo == 1;
    launch_pdl(roi_prep_kernel, dim3((n_cap + 7) / 8), dim3(256), 0, st, rois, n_rois_dev, n_cap, B, C, H, W, force_direct, w.td,
               w.n_tiles, w.slot_cap, w.prep, w.meta, w.slot, w.cnt, w.wx, w.wy4, w.tile_count, w.tile_off, w.items,
               w.direct_list, w.ctl);
    RR_LAUNCHED_K(rc, "roi_prep_kernel", st);
    if (!force_direct && C % kTC == 0) {
        launch_pdl(roi_fill_kernel, dim3((n_cap + 7) / 8), dim3(256), 0, st, w.prep, w.slot, w.wx, w.wy4, n_rois_dev, n_cap, w.td,
                                                        w.tile_off, w.tile_fill, w.list, w.list_wx, w.list_wy);
        RR_LAUNCHED_K(rc, "roi_fill_kernel", st);
        static OncePerDevice attr_once; int attr_dev;
        if (attr_once.need(&attr_dev)) {
            RR_CUDA(cudaFuncSetAttribute(roi_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem), rc);
            RR_CUDA(cudaFuncSetAttribute(roi_tile_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2Smem), rc);
            RR_CUDA(cudaFuncSetAttribute(roi_tile_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2Smem), rc);
            if (rc == 0) attr_once.mark(attr_dev);
        }
        CUtensorMap tm;
        if (algo != 2 && make_tile_tmap(feat, B, C, H, W, &tm)) {      // TMA-staged tiles (needs 16-byte rows)
            if (combine == 3)                                           // the rows the units add into start from zero
                RR_CUDA(cudaMemsetAsync(out, 0, (size_t)n_cap * C * RR_POOL * RR_POOL * sizeof(float), st), rc);
            if (relu)
                launch_pdl(roi_tile_tma_kernel<true>, dim3(sms_for_persistent()), dim3(kT2Threads), kT2Smem, st, 
                    tm, w.list, w.list_wx, w.list_wy, w.items, w.tile_off, w.tile_fill, w.ctl, C, w.td, w.partial,
                    w.arrived, w.cnt, combine == 2 ? out : (float*)nullptr, combine == 3 ? out : (float*)nullptr);
            else
                launch_pdl(roi_tile_tma_kernel<false>, dim3(sms_for_persistent()), dim3(kT2Threads), kT2Smem, st, 
                    tm, w.list, w.list_wx, w.list_wy, w.items, w.tile_off, w.tile_fill, w.ctl, C, w.td, w.partial,
                    w.arrived, w.cnt, combine == 2 ? out : (float*)nullptr, combine == 3 ? out : (float*)nullptr);
            RR_LAUNCHED_K(rc, "roi_tile_tma_kernel", st);
            if (combine == 2 && rows_mode) *rows_mode = 1;
            if (combine == 3 && rows_mode) *rows_mode = 2;
        } else {                                                        // tiles staged through the load/store path
            launch_pdl(roi_tile_kernel, dim3(2 * sms_for_persistent()), dim3(kTileThreads), kTileSmem, st, feat, w.list, w.list_wx, w.list_wy, w.items,
                                                                      w.tile_off, w.tile_fill, w.ctl, C, H, W, relu,
                                                                      w.td, w.partial);
            RR_LAUNCHED_K(rc, "roi_tile_kernel", st);
        }
    }
    if (combine == 1) {
        launch_pdl(roi_combine_kernel, dim3(n_cap), dim3(256), (size_t)C * RR_POOL * RR_POOL * sizeof(float), st, 
            w.prep, w.slot, n_rois_dev, n_cap, C, w.partial, out);
        RR_LAUNCHED_K(rc, "roi_combine_kernel", st);
    }
    launch_pdl(roi_direct_kernel, dim3(force_direct ? 8 * kSMs : 2 * sms_for_persistent()), dim3(kRoiThreads), 0, st, 
        feat, rois, w.direct_list, w.ctl, B, C, H, W, relu, out);
    RR_LAUNCHED_K(rc, "roi_direct_kernel", st);
    return rc;
}

// --------------------------------------------------------------------------------------------
// BACKWARD (training): d loss / d feat of roi_align(relu(feat)) from d loss / d out [n,C,3,3].
// The transpose of the tile path, as a GATHER: a CTA owns one (tile, 32-channel group), warp y owns row y of the
// tile and keeps that row's 32 pixels of its lane's channel in registers; it walks over the tile's pieces in list
// order and adds  sum_pw wx[pw][x] * (sum_ph wy[ph][y] * G[c,ph,pw] / count)  for the pieces that cover its row.
// Every output element is written exactly once with a plain store, no atomics on the data (torchvision's backward
// scatters 4 atomicAdds per sample); the pieces of a tile are summed in ascending RoI order: bit-reproducible.  The ReLU mask is applied in the coalesced write-out.  RoIs of the direct path
// (windows over 64 pixels) are added afterwards by a per-RoI kernel with atomicAdd.
// --------------------------------------------------------------------------------------------
#ifndef RR_BWD_VEC
#define RR_BWD_VEC 1
#endif
// One CTA of 24 warps (= the 24 tile rows) per (tile, channel group).  Measured alternatives (config 2, kernel time): two CTAs
// of 12 rows each so that two fit an SM (the write-out scratch then overlays the piece tables, which needs a block barrier
// before the write-out) 1.71 ms, three of 8 rows 2.72 ms, one CTA with that barrier 1.61 ms - against 1.49 ms: the warps of a
// CTA finish their piece loops at different times and the early ones' write-out already runs under the others' loops.
constexpr int kBwdThreads = 32 * kTH;                 // 24 warps = 24 tile rows
constexpr int kBwdWxFloats = kChunk * RR_POOL * kTW;  // weights placed at tile columns
constexpr int kBwdGFloats = kChunk * kTC * RR_POOL * RR_POOL;
constexpr int kBwdMaxSort = 2048;                      // pieces of one tile that are brought into RoI order (more: list order)
constexpr int kBwdSmemFloats = kChunk * 2 * 4 + kChunk * kTH * 4 + kBwdWxFloats + kBwdGFloats + 2 * kChunk + kTH * kTC * 33 + 2 * kBwdMaxSort;
constexpr int kBwdSmem = kBwdSmemFloats * (int)sizeof(float);

__global__ void __launch_bounds__(kBwdThreads, 1)
roi_tile_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ grad_out, const int4* __restrict__ list,
                    const float* __restrict__ list_wx, const float4* __restrict__ list_wy, const float* __restrict__ cnt_arr,
                    const int* __restrict__ tile_off, const int* __restrict__ tile_fill,
                    int C, int H, int W, int relu, TileDims td, float* __restrict__ grad_feat) {
    extern __shared__ float s_bwd[];
    int4* s_desc = reinterpret_cast<int4*>(s_bwd);
    float4* s_wy = reinterpret_cast<float4*>(s_desc + 2 * kChunk);
    float* s_wx = reinterpret_cast<float*>(s_wy + kChunk * kTH);      // [piece][pw][tile column]
    float* s_g = s_wx + kBwdWxFloats;                                // [piece][channel][9]
    float* s_inv = s_g + kBwdGFloats;
    int* s_rng = reinterpret_cast<int*>(s_inv + kChunk);             // per piece: first | last tile column with a weight
    float* s_t = s_inv + 2 * kChunk;                                 // [row][channel][33]
    int* s_ids = reinterpret_cast<int*>(s_t + kTH * kTC * 33);       // RoI index of every piece of the tile
    int* s_perm = s_ids + kBwdMaxSort;                               // pieces in ascending RoI order
    const int tid = threadIdx.x, lane = tid & 31, y = tid >> 5;
    const int ngroups = C / kTC;
    const int t = blockIdx.x / ngroups, g = blockIdx.x - t * ngroups;
    const int n_p = tile_fill[t];
    const int img = t / td.tiles_per_img, trem = t - img * td.tiles_per_img;
    const int ty = trem / td.ntx, tx = trem - ty * td.ntx;
    const int px0 = tx * kTW, py0 = ty * kTH;
    if (n_p == 0) {                                                  // no RoI touches the tile: its gradient is zero.  Every
        const int gy = py0 + y, gx = px0 + lane;                     // element of the map is written by exactly one CTA, so
        if (gy >= H) return;                                         // the map needs no memset (1.07 GB at config 2)
        if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(grad_feat) & 15) == 0) {
            const int x4 = lane & 7;
            if (px0 + 4 * x4 < W)
                for (int c = lane >> 3; c < kTC; c += 4)
                    *reinterpret_cast<float4*>(grad_feat + (((size_t)img * C + (size_t)g * kTC + c) * H + gy) * W + px0 + 4 * x4) =
                        make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (gx < W) {
            for (int c = 0; c < kTC; ++c)
                grad_feat[(((size_t)img * C + (size_t)g * kTC + c) * H + gy) * W + gx] = 0.f;
        }
        return;
    }
    float acc[kTW];
#pragma unroll
    for (int x = 0; x < kTW; ++x) acc[x] = 0.f;

    // The fill kernel hands out list positions with an atomic, so the list order of a tile's pieces changes from run to
    // run; summing them in ascending RoI order (a RoI has at most one piece per tile) makes the result bit-reproducible.
    const int off = tile_off[t];
    const bool sorted = n_p <= kBwdMaxSort;
    if (sorted) {
        for (int i = tid; i < n_p; i += kBwdThreads) s_ids[i] = __ldg(&list[2 * (size_t)(off + i)].x);
        __syncthreads();
        for (int i = tid; i < n_p; i += kBwdThreads) {
            const int me = s_ids[i];
            int rank = 0;
            for (int j = 0; j < n_p; ++j) rank += s_ids[j] < me;
            s_perm[rank] = i;
        }
    }

    for (int base = 0; base < n_p; base += kChunk) {
        const int cnt = min(kChunk, n_p - base);
        __syncthreads();                                             // previous chunk's tables are no longer read (and s_perm is complete)
        auto pos_of = [&](int p) { return off + (sorted ? s_perm[base + p] : base + p); };
        if (tid < 2 * cnt) s_desc[tid] = __ldg(list + 2 * (size_t)pos_of(tid >> 1) + (tid & 1));
        for (int i = tid; i < cnt * kTH; i += kBwdThreads) s_wy[i] = __ldg(list_wy + (size_t)pos_of(i / kTH) * kTH + i % kTH);
        for (int i = tid; i < cnt * RR_POOL * kTW; i += kBwdThreads) s_wx[i] = 0.f;
        __syncthreads();
        for (int i = tid; i < cnt * RR_POOL * kTW; i += kBwdThreads) {   // list_wx[pos][pw][i] belongs to tile column c0 + i
            const int p = i / (RR_POOL * kTW), r = i - p * (RR_POOL * kTW), pw = r / kTW, k = r - pw * kTW;
            const int4 d0 = s_desc[2 * p], d1 = s_desc[2 * p + 1];
            const int cols = pw == 0 ? d0.w : (pw == 1 ? d1.x : d1.y);
            const int c0 = cols & 0xff, ncols = (cols >> 8) & 0xff;
            if (k < ncols) s_wx[(p * RR_POOL + pw) * kTW + c0 + k] = __ldg(list_wx + ((size_t)pos_of(p) * RR_POOL + pw) * kTW + k);
        }
        for (int i = tid; i < cnt * kTC * 9; i += kBwdThreads) {         // 288 contiguous floats per (RoI, channel group)
            const int p = i / (kTC * 9), e = i - p * (kTC * 9);
            const int roi = s_desc[2 * p].x;
            s_g[i] = __ldg(grad_out + ((size_t)roi * C + (size_t)g * kTC) * 9 + e);
        }
        if (tid < cnt) {
            const int4 d0 = s_desc[2 * tid], d1 = s_desc[2 * tid + 1];
            s_inv[tid] = 1.0f / __ldg(cnt_arr + d0.x);
            int cmin = kTW, cmax = -1;
            const int colsv[RR_POOL] = {d0.w, d1.x, d1.y};
#pragma unroll
            for (int pw = 0; pw < RR_POOL; ++pw) {
                const int c0 = colsv[pw] & 0xff, nc = (colsv[pw] >> 8) & 0xff;
                if (nc > 0) { cmin = min(cmin, c0); cmax = max(cmax, c0 + nc - 1); }
            }
            s_rng[tid] = cmin | ((cmax + 1) << 8) | ((d0.z & 0xffff) << 16);      // cmax + 1: 0 for a piece without columns; r0 | nrows << 8 on top
        }
        __syncthreads();

        for (int p = 0; p < cnt; ++p) {
            const unsigned rng = (unsigned)s_rng[p];                 // one broadcast word per (row, piece): most pairs stop here
            const int r0 = (rng >> 16) & 0xff, nrows = rng >> 24;
            if ((unsigned)(y - r0) >= (unsigned)nrows) continue;     // warp-uniform: this piece does not cover row y
            const float4 wy = s_wy[p * kTH + (y - r0)];
            const float inv = s_inv[p];
            const float* gp = s_g + (p * kTC + lane) * 9;
            float s3[RR_POOL];
#pragma unroll
            for (int pw = 0; pw < RR_POOL; ++pw)
                s3[pw] = fmaf(wy.x, gp[pw], fmaf(wy.y, gp[3 + pw], wy.z * gp[6 + pw])) * inv;
            const int cmin = rng & 0xff, cmax = (int)((rng >> 8) & 0xff) - 1;
            const float4* w0 = reinterpret_cast<const float4*>(s_wx + (p * RR_POOL + 0) * kTW);
            const float4* w1 = reinterpret_cast<const float4*>(s_wx + (p * RR_POOL + 1) * kTW);
            const float4* w2 = reinterpret_cast<const float4*>(s_wx + (p * RR_POOL + 2) * kTW);
#pragma unroll
            for (int q = 0; q < kTW / 4; ++q) {
                if (4 * q + 3 >= cmin && 4 * q <= cmax) {            // warp-uniform
                    const float4 a = w0[q], b = w1[q], c = w2[q];
                    acc[4 * q + 0] = fmaf(a.x, s3[0], fmaf(b.x, s3[1], fmaf(c.x, s3[2], acc[4 * q + 0])));
                    acc[4 * q + 1] = fmaf(a.y, s3[0], fmaf(b.y, s3[1], fmaf(c.y, s3[2], acc[4 * q + 1])));
                    acc[4 * q + 2] = fmaf(a.z, s3[0], fmaf(b.z, s3[1], fmaf(c.z, s3[2], acc[4 * q + 2])));
                    acc[4 * q + 3] = fmaf(a.w, s3[0], fmaf(b.w, s3[1], fmaf(c.w, s3[2], acc[4 * q + 3])));
                }
            }
        }
    }
    // ---- write-out: transpose the warp's [channel][x] block through shared memory, 128-byte rows, ReLU mask ----
    float* st = s_t + (size_t)y * kTC * 33;
#pragma unroll
    for (int x = 0; x < kTW; ++x) st[lane * 33 + x] = acc[x];
    __syncwarp();
    const int gy = py0 + y, gx = px0 + lane;
    if (gy >= H) return;
    if (RR_BWD_VEC && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(grad_feat)) & 15) == 0) {
        // four pixels per lane: eight lanes cover the 128-byte row of one channel, a warp four channels per round.  The
        // four scalar reads of the [c][33] scratch are conflict free (bank = c + 4 x4 + j over 4 channels x 8 chunks).
        const int x4 = lane & 7, cq = lane >> 3;
        if (px0 + 4 * x4 < W) {                                        // W % 4 == 0: a chunk is inside the map or outside it
            float4 f[kTC / 4];                                         // all eight mask loads in flight before the first store
            const size_t idx0 = (((size_t)img * C + (size_t)g * kTC + cq) * H + gy) * W + px0 + 4 * x4;
            const size_t cstep = (size_t)4 * H * W;
            if (relu) {
#pragma unroll
                for (int k = 0; k < kTC / 4; ++k) f[k] = __ldg(reinterpret_cast<const float4*>(feat + idx0 + k * cstep));
            }
#pragma unroll
            for (int k = 0; k < kTC / 4; ++k) {
                const float* sp = st + (cq + 4 * k) * 33 + 4 * x4;
                float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
                if (relu) {
                    if (!(f[k].x > 0.f)) v.x = 0.f;
                    if (!(f[k].y > 0.f)) v.y = 0.f;
                    if (!(f[k].z > 0.f)) v.z = 0.f;
                    if (!(f[k].w > 0.f)) v.w = 0.f;
                }
                *reinterpret_cast<float4*>(grad_feat + idx0 + k * cstep) = v;
            }
        }
        return;
    }
    if (gx < W) {
        for (int c = 0; c < kTC; ++c) {
            const size_t idx = (((size_t)img * C + (size_t)g * kTC + c) * H + gy) * W + gx;
            float v = st[c * 33 + lane];
            if (relu && !(__ldg(feat + idx) > 0.f)) v = 0.f;
            grad_feat[idx] = v;
        }
    }
}

// direct-path RoIs: one CTA per RoI, one atomicAdd per (channel, window pixel)
__global__ void __launch_bounds__(kRoiThreads)
roi_direct_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois, const float* __restrict__ grad_out,
                      const int* __restrict__ direct_list, const int* __restrict__ ctl,
                      int B, int C, int H, int W, int relu, float* __restrict__ grad_feat) {
    __shared__ float s_ax[RR_POOL][kMaxWin];
    __shared__ float s_ay[RR_POOL][kMaxWin];
    __shared__ AxisGeom s_g[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int n_direct = ctl[kCtlDirect];
    for (int i = blockIdx.x; i < n_direct; i += gridDim.x) {
        const int n = direct_list[i];
        const float* r = rois + (size_t)n * 5;
        const int bi = (int)r[0];
        __syncthreads();
        if (bi < 0 || bi >= B) continue;
        if (tid == 0) s_g[0] = axis_geom(r[1], r[3], W);
        if (tid == 32) s_g[1] = axis_geom(r[2], r[4], H);
        __syncthreads();
        const AxisGeom gx = s_g[0], gy = s_g[1];
        if (gx.n == 0 || gy.n == 0) continue;
        const float inv = 1.0f / (float)max(gx.grid * gy.grid, 1);
        const bool fits = gx.n <= kMaxWin && gy.n <= kMaxWin;
        if (fits) {
            for (int t = tid; t < RR_POOL * gx.n; t += blockDim.x) { const int p = t / gx.n, k = t - p * gx.n; s_ax[p][k] = axis_weight(gx, p, k, W); }
            for (int t = tid; t < RR_POOL * gy.n; t += blockDim.x) { const int p = t / gy.n, k = t - p * gy.n; s_ay[p][k] = axis_weight(gy, p, k, H); }
        }
        __syncthreads();
        for (int c = warp; c < C; c += nwarp) {
            const float* gp = grad_out + ((size_t)n * C + c) * 9;
            float G[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) G[k] = __ldg(gp + k) * inv;
            const size_t plane = (((size_t)bi * C + c) * H + gy.lo) * W + gx.lo;
            for (int yy = 0; yy < gy.n; ++yy) {
                const float a0 = fits ? s_ay[0][yy] : axis_weight(gy, 0, yy, H);
                const float a1 = fits ? s_ay[1][yy] : axis_weight(gy, 1, yy, H);
                const float a2 = fits ? s_ay[2][yy] : axis_weight(gy, 2, yy, H);
                const float t0 = a0 * G[0] + a1 * G[3] + a2 * G[6];
                const float t1 = a0 * G[1] + a1 * G[4] + a2 * G[7];
                const float t2 = a0 * G[2] + a1 * G[5] + a2 * G[8];
                for (int x = lane; x < gx.n; x += 32) {
                    const float b0 = fits ? s_ax[0][x] : axis_weight(gx, 0, x, W);
                    const float b1 = fits ? s_ax[1][x] : axis_weight(gx, 1, x, W);
                    const float b2 = fits ? s_ax[2][x] : axis_weight(gx, 2, x, W);
                    const float v = b0 * t0 + b1 * t1 + b2 * t2;
                    const size_t idx = plane + (size_t)yy * W + x;
                    if (v != 0.f && (!relu || __ldg(feat + idx) > 0.f)) atomicAdd(grad_feat + idx, v);
                }
            }
        }
    }
}

int roi_align_backward_launch(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                              int B, int C, int H, int W, int relu, int algo, const float* grad_out, float* grad_feat,
                              void* ws, cudaStream_t st) {
    int rc = 0;
    RoiWs w = carve_roi(ws, n_cap, B, C, H, W);
    RR_CUDA(cudaMemsetAsync(w.zeroed, 0, w.zeroed_bytes, st), rc);
    const int force_direct = algo == 1;
    if (force_direct || C % kTC != 0)              // the tile kernel (which writes every element, zeros included) does not run
        RR_CUDA(cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, st), rc);
    if (rc) return rc;
    launch_pdl(roi_prep_kernel, dim3((n_cap + 7) / 8), dim3(256), 0, st, rois, n_rois_dev, n_cap, B, C, H, W, force_direct, w.td,
               w.n_tiles, w.slot_cap, w.prep, w.meta, w.slot, w.cnt, w.wx, w.wy4, w.tile_count, w.tile_off, w.items,
               w.direct_list, w.ctl);
    RR_LAUNCHED_K(rc, "roi_prep_kernel", st);
    if (!force_direct && C % kTC == 0) {
        launch_pdl(roi_fill_kernel, dim3((n_cap + 7) / 8), dim3(256), 0, st, w.prep, w.slot, w.wx, w.wy4, n_rois_dev, n_cap, w.td,
                                                        w.tile_off, w.tile_fill, w.list, w.list_wx, w.list_wy);
        RR_LAUNCHED_K(rc, "roi_fill_kernel", st);
        static OncePerDevice attr_once; int attr_dev;
        if (attr_once.need(&attr_dev)) {
            RR_CUDA(cudaFuncSetAttribute(roi_tile_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem), rc);
            if (rc == 0) attr_once.mark(attr_dev);
        }
        roi_tile_bwd_kernel<<<w.n_tiles * (C / kTC), kBwdThreads, kBwdSmem, st>>>(feat, grad_out, w.list, w.list_wx, w.list_wy,
                                                                              w.cnt, w.tile_off, w.tile_fill, C, H, W, relu,
                                                                              w.td, grad_feat);
        RR_LAUNCHED_K(rc, "roi_tile_bwd_kernel", st);
    }
    roi_direct_bwd_kernel<<<8 * kSMs, kRoiThreads, 0, st>>>(feat, rois, grad_out, w.direct_list, w.ctl, B, C, H, W, relu, grad_feat);
    RR_LAUNCHED_K(rc, "roi_direct_bwd_kernel", st);
    return rc;
}

}  // namespace rr

using namespace rr;

RR_API size_t rr_roi_align_workspace_bytes(int n_cap, int B, int C, int H, int W) {
    if (n_cap <= 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return roi_align_ws_bytes(n_cap, B, C, H, W);
}

RR_API int rr_roi_align(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                        int B, int C, int H, int W, int relu, int algo, float* out,
                        void* ws, size_t ws_bytes, void* stream) {
    if (n_cap == 0) return 0;
    if (!feat || !rois || !out || !ws) return RR_E_BADARG;
    if (n_cap < 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0 || algo < 0 || algo > 2) return RR_E_BADARG;
    if (C > 1024) return RR_E_RANGE;                 // roi_combine stages one RoI (C*9 floats) in shared memory
    if (ws_bytes < roi_align_ws_bytes(n_cap, B, C, H, W) || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    return roi_align_launch(feat, rois, n_rois_dev, n_cap, B, C, H, W, relu, algo, 1, out, ws, (cudaStream_t)stream, nullptr);
}

RR_API int rr_roi_align_backward(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                                 int B, int C, int H, int W, int relu, int algo, const float* grad_out, float* grad_feat,
                                 void* ws, size_t ws_bytes, void* stream) {
    if (!grad_feat || B <= 0 || C <= 0 || H <= 0 || W <= 0 || n_cap < 0 || algo < 0 || algo > 1) return RR_E_BADARG;
    if (n_cap == 0) return (int)cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, (cudaStream_t)stream);
    if (!feat || !rois || !grad_out || !ws) return RR_E_BADARG;
    if (ws_bytes < roi_align_ws_bytes(n_cap, B, C, H, W) || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    return roi_align_backward_launch(feat, rois, n_rois_dev, n_cap, B, C, H, W, relu, algo, grad_out, grad_feat, ws,
                                     (cudaStream_t)stream);
}

#ifdef RR_TILE_TRACE
RR_API int rr_debug_tile_trace(unsigned long long* host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, rr::g_tile_trace, sizeof(unsigned long long) * (size_t)n_words);
}
#endif
