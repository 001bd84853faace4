// C-ABI glue of librrnet_b200: version / error strings / launch counter and the one-call
// eval path (decode -> stage-1 NMS -> RoIAlign+ReLU -> head -> generate_bbox), i.e.
// RRNet.forward after forward_stage1 (models/rrnet.py:31-54) + RRNetOperator.generate_bbox
// (operators/rrnet_operator.py:188-209) for the whole batch with no host synchronisation.
#include "rr_common.cuh"

namespace rr {

std::atomic<uint64_t> g_launches{0};
thread_local int g_sm_reserve = 0;
extern int g_select_single_cta;   // rr_decode.cu
int g_combine_in_tile = 0;        // RR_OPT_COMBINE_IN_TILE_KERNEL: the RoIAlign tile kernel combines a RoI's slots itself (measured slower in total)
int g_pdl_enabled = 0;          // measured on the B200: no gain for one batch at a time (0.771 ms either way), -7 % with two batches in flight
thread_local KernelTrace g_ktrace = {nullptr, nullptr, 0, 0};

// implemented in the per-kernel translation units
int decode_launch(const float*, const float*, const float*, int, int, int, int, int, int, float*, int64_t*,
                  void*, cudaStream_t);
size_t decode_ws_bytes(int B);
int stage1_nms_launch(const float*, int, int, int, double, float*, float*, float*, int32_t*, void*, cudaStream_t);
size_t stage1_nms_ws_bytes(int B, int K, int C);
int roi_align_launch(const float*, const float*, const int32_t*, int, int, int, int, int, int, int, int, float*, void*,
                     cudaStream_t, int*);
void roi_align_ws_views(void*, int, int, int, int, int, const float**, const int**, const int**, const float**);
int head_forward_launch_partial(float*, const float*, const int*, const int*, const float*, const int32_t*, int,
                                const float*, float*, int, int, cudaStream_t);
size_t roi_align_ws_bytes(int n_cap, int B, int C, int H, int W);
int head_forward_launch(const float*, const int32_t*, int, const float*, float*, int, cudaStream_t);
int generate_bbox_launch(const float*, const float*, const float*, const float*, const int32_t*, int, float,
                         float*, float*, cudaStream_t);

struct EvalWs {
    void* decode; void* nms; void* roi; float* roi_feat; size_t bytes;
};
static EvalWs carve_eval(void* ws, int B, int K, int C, int H, int W, int feat_ch) {
    Carver cv(ws);
    EvalWs w;
    w.decode = cv.take<char>(decode_ws_bytes(B));
    w.nms = cv.take<char>(stage1_nms_ws_bytes(B, K, C));
    w.roi = cv.take<char>(roi_align_ws_bytes(B * K, B, feat_ch, H, W));
    w.roi_feat = cv.take<float>((size_t)B * K * feat_ch * RR_POOL * RR_POOL);
    w.bytes = cv.off;
    return w;
}

}  // namespace rr

using namespace rr;

RR_API int rr_version(void) { return 100; }

RR_API uint64_t rr_launch_count(void) { return g_launches.load(); }
RR_API int rr_set_sm_reserve(int n_sms) {
    if (n_sms < 0 || n_sms >= kSMs) return RR_E_BADARG;
    g_sm_reserve = n_sms;
    return 0;
}

RR_API int rr_set_pdl(int enabled) {
    g_pdl_enabled = enabled ? 1 : 0;
    return 0;
}
RR_API int rr_set_option(int option, int value) {
    switch (option) {
        case RR_OPT_PDL: g_pdl_enabled = value ? 1 : 0; return 0;
        case RR_OPT_SELECT_SINGLE_CTA: g_select_single_cta = value ? 1 : 0; return 0;
        case RR_OPT_COMBINE_IN_TILE_KERNEL: g_combine_in_tile = value == 2 ? 2 : (value ? 1 : 0); return 0;
        default: return RR_E_BADARG;
    }
}

RR_API int rr_kernel_trace_begin(void* const* events, const char** names, int capacity, void* stream) {
    if (!events || capacity <= 0) return RR_E_BADARG;
    g_ktrace = {events, names, capacity, 0};
    ktrace_mark("begin", (cudaStream_t)stream);
    return 0;
}
RR_API int rr_kernel_trace_end(void) {
    const int n = g_ktrace.n;
    g_ktrace = {nullptr, nullptr, 0, 0};
    return n;
}

RR_API const char* rr_error_string(int code) {
    switch (code) {
        case RR_OK: return "ok";
        case RR_E_BADARG: return "rrnet_b200: bad argument (null pointer or non-positive size)";
        case RR_E_WORKSPACE: return "rrnet_b200: workspace too small or not 256-byte aligned";
        case RR_E_RANGE: return "rrnet_b200: size outside the supported range";
        case RR_E_ALIGN: return "rrnet_b200: pointer not 16-byte aligned";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "rrnet_b200: unknown error";
}

RR_API size_t rr_eval_workspace_bytes(int B, int C, int H, int W, int K, int feat_ch) {
    if (B <= 0 || C <= 0 || K <= 0 || feat_ch <= 0 || H <= 0 || W <= 0) return 0;
    return carve_eval(nullptr, B, K, C, H, W, feat_ch).bytes;
}

RR_API int rr_eval_forward(const float* hm, const float* wh, const float* off, const float* feat,
                           int B, int C, int H, int W, int K, int feat_ch, int pool, double nms_thr,
                           int roi_algo, const float* head_folded, float scale,
                           float* out_dets, int64_t* out_inds,
                           float* out_bxyxy, float* out_scores, float* out_clses, int32_t* out_counts,
                           float* out_reg, float* out_s1, float* out_s2, float* roi_feat,
                           void* ws, size_t ws_bytes, void* stream, void* const* stage_events) {
    if (!hm || !wh || !off || !feat || !head_folded || !out_dets || !out_bxyxy || !out_scores || !out_clses ||
        !out_counts || !out_reg || !out_s1 || !out_s2 || !ws)
        return RR_E_BADARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0) return RR_E_BADARG;
    if (feat_ch != RR_HEAD_CH) return RR_E_RANGE;               // the head is 256-channel (fasterrcnn_detector.py:9)
    if ((pool != 0 && pool != 3 && pool != RR_DECODE_PRECOLLECTED) || roi_algo < 0 || roi_algo > 15 || (roi_algo & 5) == 5) return RR_E_BADARG;
    const int head_algo = (roi_algo >> 1) & 1;       // bit 1: fp32 FFMA head instead of the tcgen05 one
    const int relu = (roi_algo & 8) ? 0 : 1;         // bit 3: `feat` already went through the ReLU (models/rrnet.py:144 makes that
                                                     // tensor for the stage-1 heads anyway): RoIAlign skips its own
    roi_algo = (roi_algo & 1) ? 1 : ((roi_algo & 4) ? 2 : 0);   // bit 0: direct-gather RoIAlign, bit 2: tile path staged by loads (no TMA)
    if (C > RR_MAX_CLASSES || K > RR_MAX_TOPK || (long long)K > (long long)H * W) return RR_E_RANGE;
    if ((long long)C * H * W >= (1LL << 31)) return RR_E_RANGE;
    if (ws_bytes < carve_eval(nullptr, B, K, C, H, W, feat_ch).bytes || ((uintptr_t)ws & 255)) return RR_E_WORKSPACE;
    if (((uintptr_t)out_reg & 15) || ((uintptr_t)head_folded & 15)) return RR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    EvalWs w = carve_eval(ws, B, K, C, H, W, feat_ch);
    float* rf = roi_feat ? roi_feat : w.roi_feat;
    const int n_cap = B * K;
    int rc = 0;
    auto mark = [&](int i) {
        if (stage_events && stage_events[i]) RR_CUDA(cudaEventRecord((cudaEvent_t)stage_events[i], st), rc);
    };
    mark(0);
    rc = decode_launch(hm, wh, off, B, C, H, W, K, pool, out_dets, out_inds, w.decode, st);
    if (rc) return rc;
    mark(1);
    rc = stage1_nms_launch(out_dets, B, K, C, nms_thr, out_bxyxy, out_scores, out_clses, out_counts, w.nms, st);
    if (rc) return rc;
    mark(2);
    const int32_t* n_dev = out_counts + B;
    // roi_feat requested: materialise it (combine) and feed the head from it; otherwise the head sums the
    // tile-path partial slots itself and only direct-path RoIs go through the buffer
    const int fused = roi_feat == nullptr && roi_algo != 1;
    // fused: the head sums the RoIAlign partial slots (rows_mode = 0, default), or - RR_OPT_COMBINE_IN_TILE_KERNEL - the tile
    // kernel combines a RoI's slots itself and leaves the RoI's [9][256] row in rf (rows_mode = 1; needs the TMA kernel)
    int rows_mode = 0;
    rc = roi_align_launch(feat, out_bxyxy, n_dev, n_cap, B, feat_ch, H, W, relu, roi_algo,
                          fused ? (g_combine_in_tile == 2 ? 3 : (g_combine_in_tile ? 2 : 0)) : 1, rf, w.roi, st, &rows_mode);
    if (rc) return rc;
    mark(3);
    if (fused) {
        const float* partial; const int* slot; const int* pieces; const float* count;
        roi_align_ws_views(w.roi, n_cap, B, feat_ch, H, W, &partial, &slot, &pieces, &count);
        rc = head_forward_launch_partial(rf, partial, slot, pieces, count, n_dev, n_cap, head_folded, out_reg, head_algo, rows_mode, st);
    } else {
        rc = head_forward_launch(rf, n_dev, n_cap, head_folded, out_reg, head_algo, st);
    }
    if (rc) return rc;
    mark(4);
    rc = generate_bbox_launch(out_bxyxy, out_reg, out_scores, out_clses, n_dev, n_cap, scale, out_s1, out_s2, st);
    mark(5);
    return rc;
}
