/*
 * rrnet_b200.h -- C ABI of librrnet_b200.so: RRNet's post-backbone detection hot path
 * as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary.  The reference (ouc-ocean-group/RRNet) has exactly one
 * native ABI on this path -- `_nms` in ext/nms/nms/gpu_nms.hpp:1-2 -- and otherwise calls
 * torch / torchvision ops from Python (models/rrnet.py, operators/rrnet_operator.py,
 * modules/loss/functional.py, datasets/transforms/functional.py).  Each entry point below
 * names the reference interface it replaces (file:line relative to the reference root).
 * INTEGRATION.md shows the reference-side binding (Cython `cdef extern` / ctypes) for each.
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars; no torch / CUDA types in signatures
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - tensors are dense row-major fp32 (NCHW for maps) unless stated; indices int32/int64.
 *   - return value: 0 = ok, >0 = cudaError_t of a failed CUDA call,
 *     <0 = argument error (RR_E_*).  rr_error_string() describes either.
 *   - device-pointer entry points never allocate, never synchronise and never touch the host:
 *     the caller owns every buffer and provides scratch sized by the matching
 *     *_workspace_bytes() query (256-byte aligned).  They are CUDA-graph capturable.
 *   - `*_host` entry points are synchronous conveniences that own their device scratch
 *     (grow-only cache per device, mutex protected).
 *   - variable-length results use fixed-capacity outputs plus device-side counts.
 */
#ifndef RRNET_B200_H_
#define RRNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_OK              0
#define RR_E_BADARG       (-1)   /* null pointer / non-positive size */
#define RR_E_WORKSPACE    (-2)   /* workspace too small or misaligned */
#define RR_E_RANGE        (-3)   /* size outside the supported range (see each function) */
#define RR_E_ALIGN        (-4)   /* pointer not aligned as required */

#define RR_MAX_TOPK        16384 /* decode: K <= RR_MAX_TOPK and K <= H*W */
#define RR_MAX_CLASSES     64    /* stage-1 NMS: num_classes <= RR_MAX_CLASSES */
#define RR_HEAD_CH         256   /* re-regression head input channels (fixed by the reference) */
#define RR_HEAD_MID        64
#define RR_POOL            3     /* RoIAlign output is 3x3 (models/rrnet.py:51) */
#define RR_DECODE_RAW_SCORES 0x100
#define RR_DECODE_PRECOLLECTED 0x200 /* the candidate lists in `ws` were filled by rr_hm_tail_collect: skip sample + collect */

int         rr_version(void);
const char* rr_error_string(int code);
/* number of kernel launches issued by this library since load (all entry points) */
uint64_t    rr_launch_count(void);
/* Number of SMs (0..147, default 0) that the persistent kernels (tile RoIAlign, tensor-core head) leave free, so that
 * the short latency-bound kernels of ANOTHER batch on another stream can run next to them (bench.py --streams 2).
 * Per calling host thread (thread-local); read when that thread issues / captures the next launch. */
int         rr_set_sm_reserve(int n_sms);
/* Per-kernel device times for benchmarks (no reference counterpart; the reference has no timers on this path).
 * After rr_kernel_trace_begin every kernel launched by the CALLING THREAD through this library records the next of
 * the caller's `capacity` CUDA events (cudaEvent_t, timing enabled) on its launch stream; events[0] is recorded on
 * `stream` by the call itself.  names[i] (optional, static strings) = kernel that ended at events[i].
 * rr_kernel_trace_end() stops the trace and returns the number of events recorded.  Not for use under graph capture. */
/* Programmatic dependent launch between the kernels of the eval path (default on): 0 = ordinary stream order. */
int         rr_set_pdl(int enabled);
/* Process-wide switches between equivalent implementations (same results), for measurements:
 *   RR_OPT_PDL                as rr_set_pdl
 *   RR_OPT_SELECT_SINGLE_CTA  1 = the decode's top-K selection by ONE CTA per image instead of a cluster of 8
 *   RR_OPT_COMBINE_IN_TILE_KERNEL  1 = in rr_eval_forward the RoIAlign tile kernel combines a RoI's partial slots itself as soon
 *                             as its last piece is done (deterministic, bit-identical; default 0: the head sums the slots);
 *                             2 = no slots at all: the tile kernel's units add their bins into the RoI's zeroed row with float
 *                             reductions (red.global.add.f32).  NOT bit-reproducible for RoIs cut into three or more pieces
 *                             (the order of the additions is the hardware's); within 1e-6 of the default */
#define RR_OPT_PDL 1
#define RR_OPT_SELECT_SINGLE_CTA 2
#define RR_OPT_COMBINE_IN_TILE_KERNEL 3
int         rr_set_option(int option, int value);
int         rr_kernel_trace_begin(void* const* events, const char** names, int capacity, void* stream);
int         rr_kernel_trace_end(void);

/* ------------------------------------------------------------------------------------------
 * Decode: replaces RRNet.transform_bbox + RRNet._topk + _gather_feat /
 * _transpose_and_gather_feat (models/rrnet.py:83-138).
 *   hm  [B,C,H,W] heat-map LOGITS (sigmoid is applied inside, :119)
 *   wh  [B,2,H,W] (ch0 = w, ch1 = h), off [B,2,H,W] (ch0 = x, ch1 = y)
 *   pool: 0 = RRNet's path (no peak suppression); 3 = 3x3 max-pool peak keep
 *         (operators/centernet_operator.py:204-210 semantics).  OR-ing RR_DECODE_RAW_SCORES
 *         into it makes the call a plain RRNet._topk (models/rrnet.py:93-109): hm then holds
 *         scores (no sigmoid is applied) and wh / off may be NULL (x1=x2=x, y1=y2=y).
 *   out_dets [B,K,6] = x1,y1,x2,y2,score,cls  sorted by score descending (ties: flat index asc)
 *   out_inds [B,K]   = y*W+x (int64, as the reference's `inds`), may be NULL
 * One global top-K over C*H*W per image == the reference's two-stage top-K on tie-free input.
 * Requires 0 < K <= min(H*W, RR_MAX_TOPK), C*H*W < 2^31.
 * ---------------------------------------------------------------------------------------- */
size_t rr_decode_workspace_bytes(int B, int C, int H, int W, int K);
int rr_decode_topk(const float* hm, const float* wh, const float* off,
                   int B, int C, int H, int W, int K, int pool,
                   float* out_dets, int64_t* out_inds,
                   void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage-1 head tail fused with the decode's candidate collection: replaces the LAST layer of the
 * heat-map head, nn.Conv2d(256, planes, (1,1)) with bias (detectors/centernet_detector.py:11-15, called from
 * RRNet.forward_stage1, models/rrnet.py:140-153), plus decode's sample + streaming pass over the logits.
 *   t       [B,Cin,H,W]  output of the head's 3x3 conv + ReLU (BasicCov, :81-93; stays with cuDNN)
 *   weight  [Cout,Cin]   (the 1x1 kernel), bias [Cout];  Cout <= 16, Cin <= 1024
 *   hm_out  [B,Cout,H,W] the heat-map logits (what the reference's head returns)
 *   decode_ws: a workspace of rr_decode_workspace_bytes(B,..) bytes - or the START of an rr_eval_forward workspace
 *           (its first region is the decode workspace); on return it holds the per-image candidate lists, and
 *           rr_decode_topk / rr_eval_forward called with pool = RR_DECODE_PRECOLLECTED on (hm_out, same ws) go
 *           straight to the top-K selection.  One pass over t; fp32 FMA (1e-5 against the reference's convolution).
 * ---------------------------------------------------------------------------------------- */
int rr_hm_tail_collect(const float* t, const float* weight, const float* bias, int B, int Cin, int Cout,
                       int H, int W, int K, float* hm_out, void* decode_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage-1 NMS: replaces RRNet.nms (default per-class branch, models/rrnet.py:56-72, i.e.
 * torchvision.ops.nms per class: no "+1", suppress iff IoU > thr) and the per-image loop /
 * concatenation of RRNet.forward (:37-49), for the whole batch in one call.
 *   dets [B,K,6] as produced by rr_decode_topk (score-descending per image)
 *   out_bxyxy  [B*K,5] = (image index as float, x1,y1,x2,y2)   image-major, class-ascending,
 *   out_scores [B*K], out_clses [B*K] (float, 0-based)          score-descending within class
 *   out_counts [B+1] int32: rows kept per image, then the total N in out_counts[B]
 * ---------------------------------------------------------------------------------------- */
size_t rr_stage1_nms_workspace_bytes(int B, int K, int num_classes);
int rr_stage1_nms(const float* dets, int B, int K, int num_classes, double thr,
                  float* out_bxyxy, float* out_scores, float* out_clses, int32_t* out_counts,
                  void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Generic segmented hard NMS: replaces ext.nms.nms_wrapper.nms / gpu_nms / cpu_nms /
 * py_cpu_nms (ext/nms/nms_wrapper.py:23-33, nms/gpu_nms.pyx:16-31, nms/cpu_nms.pyx:122-173,
 * nms/py_cpu_nms.py:4-32) and torchvision.ops.nms, batched over S independent segments.
 *   boxes [M,4] x1,y1,x2,y2; scores [M]; seg_offsets [S+1] int32 (device), ascending,
 *   seg_offsets[0]=0, seg_offsets[S]=M.  Sorting (stable, score desc) is done on the device.
 *   pixel_offset: 0 torchvision, 1 legacy "+1" areas;  ge_cmp: 0 suppress iff IoU > thr,
 *   1 iff IoU >= thr (cpu_nms.pyx:170).  thr is compared as a double against the fp32 IoU.
 *   keep_idx [M] int32: for segment s, entries [seg_offsets[s], +keep_count[s]) are the kept
 *   row indices (global, into boxes) in acceptance order; keep_count [S] int32.
 *   max_seg = upper bound of any segment length (host-known; M is always valid).
 * ---------------------------------------------------------------------------------------- */
size_t rr_nms_workspace_bytes(int M, int S);
int rr_nms_batched(const float* boxes, const float* scores, const int32_t* seg_offsets,
                   int M, int S, double thr, int pixel_offset, int ge_cmp,
                   int32_t* keep_idx, int32_t* keep_count,
                   void* ws, size_t ws_bytes, void* stream);

/* The reference's only native ABI, signature-identical to `_nms`
 * (ext/nms/nms/gpu_nms.hpp:1-2, nms_kernel.cu:91-144): HOST buffers, rows already sorted by
 * score (gpu_nms.pyx:25-28), "+1" areas, suppress iff IoU > thresh.  Synchronous.
 * keep_out must hold boxes_num ints.  Unlike the reference it reports errors (return code)
 * and reduces the bit-mask on the device (no mask D2H, no host scan). */
int rr_nms_legacy_host(int* keep_out_host, int* num_out_host, const float* boxes_host,
                       int boxes_num, int boxes_dim, float nms_overlap_thresh, int device_id);

/* ------------------------------------------------------------------------------------------
 * Soft-NMS: replaces ext.nms.nms_wrapper.soft_nms -> cpu_soft_nms
 * (ext/nms/nms_wrapper.py:13-19, nms/cpu_nms.pyx:17-120) as called per class by
 * RRNetOperator._ext_nms (operators/rrnet_operator.py:211-232), batched over S segments.
 *   boxes [M,5] x1,y1,x2,y2,score IN/OUT: on return the first keep_count[s] rows of each
 *   segment are the survivors in selection order with decayed scores.
 *   src_idx [M] int32 (optional): original row (global) now sitting at each kept position.
 *   method: 1 linear, 2 gaussian, else hard (cpu_nms.pyx:92-103).
 *   ws is only needed when M > 6144 (segments larger than the shared-memory capacity run in
 *   place in global memory); it may be NULL otherwise.
 * ---------------------------------------------------------------------------------------- */
size_t rr_soft_nms_workspace_bytes(int M);
int rr_soft_nms_batched(float* boxes, const int32_t* seg_offsets, int M, int S,
                        float sigma, float Nt, float threshold, int method,
                        int32_t* src_idx, int32_t* keep_count,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * RoIAlign(+ReLU): replaces torchvision.ops.roi_align(torch.relu(feat), rois, (3,3))
 * (models/rrnet.py:51): spatial_scale 1, sampling_ratio -1 (adaptive), aligned False.
 *   feat [B,C,H,W]; rois [n_cap,5] = (image index as float, x1,y1,x2,y2)
 *   n_rois_dev: device int32 holding the live row count (rows >= it are skipped), or NULL
 *   to process all n_cap rows.  relu != 0 applies max(v,0) to every tap (fused ReLU).
 *   algo: 0 = tile-centric (each feature tile is staged once in shared memory and serves every
 *   RoI piece that crosses it; RoIs wider/taller than 64 pixels or over the partial-slot budget
 *   take the direct path; the tiles arrive by TMA when W % 4 == 0 and feat is 16-byte aligned),
 *   1 = direct gather for every RoI, 2 = tile-centric with tiles staged by ordinary loads (the
 *   kernel used when TMA cannot be).  All are deterministic.
 *   out [n_cap,C,3,3].  C <= 1024.
 * ---------------------------------------------------------------------------------------- */
size_t rr_roi_align_workspace_bytes(int n_cap, int B, int C, int H, int W);
int rr_roi_align(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                 int B, int C, int H, int W, int relu, int algo, float* out,
                 void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Re-regression head (eval mode): replaces RRNet.forward_stage2 ->
 * FasterRCNNDetector.forward -> Bottleneck.forward (models/rrnet.py:155-157,
 * detectors/fasterrcnn_detector.py:13-18, backbones/resnet.py:33-53).
 * rr_head_fold folds the three BatchNorms (running stats, eps 1e-5) into the convolutions
 * once and also prepares the (hi, lo) tf32 weight tiles of the tensor-core kernel; parameters use
 * the reference's state_dict layouts:
 *   w1 [64,256]  bn1 [4,64]  (rows gamma,beta,running_mean,running_var)
 *   w2 [64,64,3,3] bn2 [4,64]   w3 [256,64] bn3 [4,256]   wr [4,256]  br [4]
 * folded: rr_head_folded_floats() floats, opaque layout, 16-byte aligned.
 *   roi_feat [n_cap,256,3,3] -> reg [n_cap,4]
 *   algo: 0 = tcgen05 tensor cores, fp32 accuracy through a 3xTF32 split (default),
 *         1 = fp32 FFMA kernel.
 * ---------------------------------------------------------------------------------------- */
size_t rr_head_folded_floats(void);
int rr_head_fold(const float* w1, const float* bn1, const float* w2, const float* bn2,
                 const float* w3, const float* bn3, const float* wr, const float* br,
                 float* folded, void* stream);
int rr_head_forward(const float* roi_feat, const int32_t* n_rois_dev, int n_cap,
                    const float* folded, int algo, float* reg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Box decode of stage 2: replaces RRNetOperator.generate_bbox
 * (operators/rrnet_operator.py:188-209) for ALL rows at once (the reference handles one
 * image per call; rows are image-major so per-image results are row ranges given by the
 * stage-1 counts).  scale = cfg.Train.scale_factor (4).
 *   s1 [n_cap,6] = X1,Y1,w,h,score,0      s2 [n_cap,6] = x,y,w,h,score,cls+1
 * ---------------------------------------------------------------------------------------- */
int rr_generate_bbox(const float* bxyxy, const float* reg, const float* scores, const float* clses,
                     const int32_t* n_rois_dev, int n_cap, float scale,
                     float* s1, float* s2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole eval path in one call: decode -> stage-1 NMS -> RoIAlign(+ReLU) -> head ->
 * generate_bbox, i.e. RRNet.forward after forward_stage1 (models/rrnet.py:31-54) plus
 * RRNetOperator.generate_bbox for every image.  All launches go to `stream`, no host sync.
 * Outputs have capacity B*K rows; counts [B+1] as in rr_stage1_nms.
 * roi_feat may be NULL (the workspace then holds it).  roi_algo: bit 0 as in rr_roi_align (1 = direct
 * gather), bit 1 selects the head kernel (0 = tcgen05, 2 = fp32 FFMA), bit 2 (4) = rr_roi_align's algo 2,
 * bit 3 (8) = `feat` is ALREADY relu(pre_feat[-1]) - the tensor forward_stage1 builds for the stage-1 heads
 * (models/rrnet.py:144) - so RoIAlign does not apply the ReLU again (same results, fewer instructions).
 * stage_events: NULL, or 6 cudaEvent_t handles (as void*) recorded on `stream` before decode and
 * after decode, stage-1 NMS, RoIAlign, head and generate_bbox (a per-stage timing hook; recording
 * an event does not synchronise).
 * ---------------------------------------------------------------------------------------- */
size_t rr_eval_workspace_bytes(int B, int C, int H, int W, int K, int feat_ch);
int rr_eval_forward(const float* hm, const float* wh, const float* off, const float* feat,
                    int B, int C, int H, int W, int K, int feat_ch, int pool, double nms_thr,
                    int roi_algo, const float* head_folded, float scale,
                    float* out_dets, int64_t* out_inds,
                    float* out_bxyxy, float* out_scores, float* out_clses, int32_t* out_counts,
                    float* out_reg, float* out_s1, float* out_s2, float* roi_feat,
                    void* ws, size_t ws_bytes, void* stream, void* const* stage_events);

/* ------------------------------------------------------------------------------------------
 * Training targets: replaces to_heatmap / gaussian_radius / gaussian2d / draw_umich_gaussian
 * (datasets/transforms/functional.py:177-262) and the zero-padding of collate_fn_ctnet
 * (datasets/drones_det.py:70-94), for a whole batch on the device.
 *   annos [B,max_n,8] = x,y,w,h,score,cls(1-based),.. input pixels; n_obj [B] int32 live rows
 *   hm [B,cls_num,img_h/sf,img_w/sf] (zeroed inside), wh [B,max_n,2], ind [B,max_n,1] (float),
 *   offset [B,max_n,2], reg_mask [B,max_n,1] (float 0/1); padded rows are written as 0.
 * ---------------------------------------------------------------------------------------- */
int rr_render_targets(const float* annos, const int32_t* n_obj, int B, int max_n,
                      int img_h, int img_w, int scale_factor, int cls_num,
                      float* hm, float* wh, float* ind, float* offset, float* reg_mask,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Heat-map focal loss: replaces focal_loss_for_hm / FocalLossHM
 * (modules/loss/functional.py:25-51, modules/loss/focalloss.py:15-20) together with the
 * caller's clamp(sigmoid(logits),1e-4,1-1e-4) (operators/rrnet_operator.py:55).
 *   logits, gt: n fp32 elements each (any shape, dense)
 *   stats [4] fp32 (device): loss, pos_sum, neg_sum, num_pos
 *   forward : writes stats.                     backward: grad[i] = upstream * dloss/dlogit_i
 *   fwd_bwd : both in one launch (cooperative grid sync; logits/gt are re-read from L2).
 * ---------------------------------------------------------------------------------------- */
size_t rr_focal_workspace_bytes(int64_t n);
int rr_focal_forward(const float* logits, const float* gt, int64_t n, float* stats,
                     void* ws, size_t ws_bytes, void* stream);
int rr_focal_backward(const float* logits, const float* gt, int64_t n, const float* stats,
                      float upstream, float* grad, void* stream);
int rr_focal_fwd_bwd(const float* logits, const float* gt, int64_t n, float upstream,
                     float* stats, float* grad, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused target render + heat-map focal loss: rr_render_targets' heat-map (to_heatmap,
 * datasets/transforms/functional.py:230-262) is rendered tile by tile in shared memory inside the focal loss
 * (modules/loss/functional.py:25-51 + operators/rrnet_operator.py:55) and never written to
 * memory -- the training step's heat-map term straight from the padded annotations.
 *   logits [B,cls_num,img_h/sf,img_w/sf] (img_w/sf a multiple of 4), annos [B,max_n,8], n_obj [B]
 *   forward : stats [4] = loss, pos_sum, neg_sum, num_pos; gt_out (optional, may be NULL) receives
 *             the rendered heat-map as well.      backward: grad = upstream * dloss/dlogits.
 *   fwd_bwd : both in ONE pass over the logits (stats and grad = upstream * dloss/dlogits): num_pos, which scales
 *             the gradient, is counted from the annotations first (distinct (class, centre cell) pairs).
 * ---------------------------------------------------------------------------------------- */
size_t rr_focal_render_workspace_bytes(int B, int cls_num, int img_h, int img_w, int scale_factor);
int rr_focal_render_forward(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                            int img_h, int img_w, int scale_factor, int cls_num,
                            float* stats, float* gt_out, void* ws, size_t ws_bytes, void* stream);
int rr_focal_render_backward(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                             int img_h, int img_w, int scale_factor, int cls_num,
                             const float* stats, float upstream, float* grad, void* stream);
int rr_focal_render_fwd_bwd(const float* logits, const float* annos, const int32_t* n_obj, int B, int max_n,
                            int img_h, int img_w, int scale_factor, int cls_num, float upstream,
                            float* stats, float* grad, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward of rr_roi_align (training; SURVEY 8b `_backward`): grad_feat [B,C,H,W] = d loss / d feat of
 * roi_align(relu(feat), rois, (3,3)) given grad_out [n_cap,C,3,3].  Same arguments, algo and workspace size as the
 * forward (the RoI bookkeeping is recomputed).  Tile path: a gather, every element written once with a plain store, pieces
 * summed in ascending RoI order (bit-reproducible); direct-path RoIs (windows over 64 pixels) are added with atomicAdd.  grad_feat is fully written.
 * ---------------------------------------------------------------------------------------- */
int rr_roi_align_backward(const float* feat, const float* rois, const int32_t* n_rois_dev, int n_cap,
                          int B, int C, int H, int W, int relu, int algo, const float* grad_out, float* grad_feat,
                          void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * RegL1Loss of a regression map (wh or offset): modules/loss/regl1loss.py:9-17 and its backward, without the
 * NHWC permute copy of the whole map.
 *   output [B,c,H,W]  mask [B,max_n] (0/1 as float)  ind [B,max_n] (flat y*W+x, as FLOAT like the collate pads it)
 *   target [B,max_n,c]
 *   loss [1]  = sum |pred*mask - target*mask| / (c*sum(mask) + 1e-4),  pred[b,k,ch] = output[b,ch,ind[b,k]]
 *   grad [B,c,H,W] or NULL: d loss / d output * grad_scale (zero-filled, then the <= B*max_n*c entries scattered)
 * One memset + one launch; the sums are fixed-order doubles (bit-reproducible loss).
 * ---------------------------------------------------------------------------------------- */
int rr_regl1_fwd_bwd(const float* output, const float* mask, const float* ind, const float* target,
                     int B, int c, int H, int W, int max_n, float grad_scale,
                     float* loss, float* grad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage-2 regression loss of RRNetOperator.criterion (operators/rrnet_operator.py:64-84) and its backward, one CTA
 * per image: box_iou(bxyxy*scale, gt) -> max over the ground truth -> IoU > 0.5 (no positive in an image: that image
 * contributes 0) -> generate_bbox_target (:86-102) -> smooth_l1_loss(mean) / B.
 *   bxyxy [N,5] image-major, seg_offsets [B+1] row ranges of the images, s2_reg [N,4],
 *   gt_xyxy [B,max_n,gt_stride] (columns 0..3 = x1,y1,x2,y2 in input pixels, AFTER the in-place xywh -> xyxy of :67;
 *   zero-padded rows take part like in the reference)
 *   loss_parts [B] (their sum is the loss); grad_reg [N,4] / grad_box [N,4] (d loss / d bxyxy[:,1:5], through the
 *   targets, which the reference does not detach) or NULL, both times grad_scale.
 * ---------------------------------------------------------------------------------------- */
int rr_stage2_loss(const float* bxyxy, const int32_t* seg_offsets, const float* s2_reg, const float* gt_xyxy,
                   int B, int max_n, int gt_stride, float scale, float grad_scale,
                   float* loss_parts, float* grad_reg, float* grad_box, void* stream);

/* ------------------------------------------------------------------------------------------
 * True-positive matching of the evaluation, utils/metrics/metrics.py:51-136 (`get_tp`) for B images in one launch.
 *   pred [B,M,6] = x,y,w,h,score,cls (n_pred [B] valid rows each), target [B,N,6] = x,y,w,h,*,cls with cls 0 = ignore
 *   region (n_tgt [B]), thresholds [T] IoU thresholds; M, N <= 768, T <= 16, cls_num <= 32 (the reference: 500, 500, 10, 11)
 *   order [B,M]        index into pred of the p-th best detection (-1 past n_pred)
 *   tp [B,M,T]         1 where that detection is a true positive at the threshold
 *   out_cls [B,M]      its class, or -1 when the reference would not emit it (inside an ignore region, or no ground
 *                      truth of its class in the image)
 *   target_count, in_img [B,cls_num-1]   ground-truth boxes per class / 1 if the class occurs (after the ignore filter)
 * ---------------------------------------------------------------------------------------- */
int rr_ap_match(const float* pred, const int32_t* n_pred, const float* target, const int32_t* n_tgt,
                const float* thresholds, int B, int M, int N, int T, int cls_num,
                int32_t* order, float* tp, int32_t* out_cls, float* target_count, float* in_img, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRNET_B200_H_ */
