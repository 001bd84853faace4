/*
 * rr_oracle.c -- CPU restatement of RRNet's post-backbone hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing here is on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the timed CPU baseline.
 *
 * Parity pinning: every function below is checked (tests/test_oracle_golden.py)
 * against fixtures in tests/golden/ that were produced by importing the reference's
 * own Python from /root/reference (tests/golden/make_golden.py) and by the reference's
 * compiled Cython NMS (oracle/_ref), plus the reference's two in-tree known answers:
 * the 5-box vector of ext/nms/nms_wrapper.py:37-55 and the demo annotation heat-map
 * (data/demo/annotations/0000364_01765_d_0000782.txt).
 *
 * All arithmetic is IEEE fp32 evaluated in the reference's operation order; build with
 * -ffp-contract=off (see oracle/Makefile) so no multiply-add is contracted.
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Third-party steps (torch.topk, torchvision.ops.nms / roi_align)
 * are restated from their published algorithms and checked against torch 2.11 /
 * torchvision 0.26 CPU outputs in the same fixtures.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_version(void) { return 1; }

ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------
 * Ordering key for fp32 values: monotone map float -> uint32 (larger float, larger key).
 * Same transform torch's radix select uses (ATen/native/cuda/SortingRadixSelect.cuh:27).
 * -0.0 sorts below +0.0; NaNs are not expected on this path.
 * ---------------------------------------------------------------------------------- */
static inline uint32_t f2key(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

static inline float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

/* ====================================================================================
 * (a2,a3) Decode: models/rrnet.py:93-138 (RRNet._topk, RRNet.transform_bbox).
 *
 * The reference takes top-K per class over H*W (:96) and then top-K over the C*K
 * concatenation (:102); on tie-free input this equals ONE top-K over C*H*W per image
 * (SURVEY B.1).  The reference ranks sigmoid(hm); sigmoid is monotone, so ranking is done
 * on the logit itself (ties between *different* logits whose fp32 sigmoid collide are
 * outside the bit-exact contract; the canonical order is logit desc, flat index asc).
 *
 *   pool = 0 : RRNet's actual path (no peak suppression).
 *   pool = 3 : operators/centernet_operator.py:204-210 (_ctnet_nms): keep iff value equals
 *              the 3x3 max (padding -inf); others become score 0 (logit -inf).
 *
 * out_dets [B,K,6] = x1,y1,x2,y2,score,cls   (models/rrnet.py:133-137 operation order)
 * out_inds [B,K]   = flat y*W+x (int64)       (:98,:104)
 * out_flat [B,K]   = flat index into C*H*W (int64), i.e. cls*H*W + ind (for tests)
 * ================================================================================== */
typedef struct { uint32_t key; uint32_t idx; } cand_t;

/* a ranks before b: larger key first, then lower flat index */
static inline int cand_before(cand_t a, cand_t b) {
    return (a.key > b.key) || (a.key == b.key && a.idx < b.idx);
}

static void heap_sift_down(cand_t* h, int n, int i) {
    /* min-heap w.r.t. ranking: root = the WORST of the kept candidates */
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_before(h[m], h[l])) m = l;
        if (r < n && cand_before(h[m], h[r])) m = r;
        if (m == i) return;
        cand_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

static int cand_cmp_desc(const void* pa, const void* pb) {
    cand_t a = *(const cand_t*)pa, b = *(const cand_t*)pb;
    if (cand_before(a, b)) return -1;
    if (cand_before(b, a)) return 1;
    return 0;
}

static inline float pooled_logit(const float* plane, int H, int W, int y, int x, int pool) {
    float v = plane[(size_t)y * W + x];
    if (pool != 3) return v;
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            if (plane[(size_t)yy * W + xx] > v) return -INFINITY;
        }
    }
    return v;
}

ORC_API int orc_decode(const float* hm, const float* wh, const float* off,
                       int B, int C, int H, int W, int K, int pool,
                       float* out_dets, int64_t* out_inds, int64_t* out_flat) {
    const size_t HW = (size_t)H * W, CHW = (size_t)C * HW;
    if (K <= 0 || (size_t)K > HW) return -1;   /* torch.topk would raise (:96) */
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (size_t)K);
        if (!heap) { rc = -2; continue; }
        int n = 0;
        const float* hb = hm + (size_t)b * CHW;
        for (int c = 0; c < C; ++c) {
            const float* plane = hb + (size_t)c * HW;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    cand_t cd;
                    cd.key = f2key(pooled_logit(plane, H, W, y, x, pool));
                    cd.idx = (uint32_t)((size_t)c * HW + (size_t)y * W + x);
                    if (n < K) {
                        heap[n++] = cd;
                        if (n == K)
                            for (int i = K / 2 - 1; i >= 0; --i) heap_sift_down(heap, K, i);
                    } else if (cand_before(cd, heap[0])) {
                        heap[0] = cd;
                        heap_sift_down(heap, K, 0);
                    }
                }
        }
        qsort(heap, (size_t)K, sizeof(cand_t), cand_cmp_desc);
        for (int k = 0; k < K; ++k) {
            uint32_t flat = heap[k].idx;
            int cls = (int)(flat / HW);
            int ind = (int)(flat % HW);
            int yi = ind / W, xi = ind % W;                       /* :99-100 */
            float logit = pooled_logit(hb + (size_t)cls * HW, H, W, yi, xi, pool);
            float score = (logit == -INFINITY) ? 0.0f : sigmoidf_(logit);   /* :119 */
            const float* whb = wh + (size_t)b * 2 * HW;
            const float* ofb = off + (size_t)b * 2 * HW;
            float xs = (float)xi + ofb[ind];                      /* :126 offset ch0 = x */
            float ys = (float)yi + ofb[HW + ind];                 /* :127 offset ch1 = y */
            float w = whb[ind];      if (!(w >= 0.0f)) w = (w != w) ? w : 0.0f;  /* :128 clamp(min=0) */
            float h = whb[HW + ind]; if (!(h >= 0.0f)) h = (h != h) ? h : 0.0f;
            float px = xs - w / 2.0f;                             /* :133 */
            float py = ys - h / 2.0f;                             /* :134 */
            float* o = out_dets + ((size_t)b * K + k) * 6;
            o[0] = px; o[1] = py; o[2] = w + px; o[3] = h + py;    /* :137 */
            o[4] = score; o[5] = (float)cls;
            if (out_inds) out_inds[(size_t)b * K + k] = ind;
            if (out_flat) out_flat[(size_t)b * K + k] = flat;
        }
        free(heap);
    }
    return rc;
}

/* ====================================================================================
 * (a5,a10) Hard NMS, the greedy rule shared by the three variants in the reference:
 *   torchvision.ops.nms           (models/rrnet.py:69,78)        pixel_offset=0, '>'
 *   ext/nms nms_kernel.cu:24-32,71 / py_cpu_nms.py:4-32          pixel_offset=1, '>'
 *   ext/nms cpu_nms.pyx:122-173 (:170)                           pixel_offset=1, '>='
 * Order = stable sort by score descending (ties -> lower index first).  The overlap is an
 * fp32 value compared against the threshold as a double, as the CPU implementations do.
 * boxes [n,4] (x1,y1,x2,y2), keep_out [n] receives ORIGINAL indices in acceptance order.
 * ================================================================================== */
typedef struct { float s; int i; } sidx_t;
static int sidx_cmp(const void* pa, const void* pb) {
    const sidx_t* a = (const sidx_t*)pa; const sidx_t* b = (const sidx_t*)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) - (a->i < b->i);
}

static inline float iou_f32(const float* a, const float* b, float area_a, float area_b, float o) {
    float xx1 = a[0] > b[0] ? a[0] : b[0];
    float yy1 = a[1] > b[1] ? a[1] : b[1];
    float xx2 = a[2] < b[2] ? a[2] : b[2];
    float yy2 = a[3] < b[3] ? a[3] : b[3];
    float w = xx2 - xx1 + o; if (!(w > 0.0f)) w = 0.0f;
    float h = yy2 - yy1 + o; if (!(h > 0.0f)) h = 0.0f;
    float inter = w * h;
    return inter / (area_a + area_b - inter);
}

ORC_API int orc_nms(const float* boxes, const float* scores, int n, double thr,
                    int pixel_offset, int ge_cmp, int32_t* keep_out) {
    if (n <= 0) return 0;
    const float o = pixel_offset ? 1.0f : 0.0f;
    sidx_t* ord = (sidx_t*)malloc(sizeof(sidx_t) * (size_t)n);
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) {
        ord[i].s = scores[i]; ord[i].i = i;
        const float* b = boxes + (size_t)i * 4;
        area[i] = (b[2] - b[0] + o) * (b[3] - b[1] + o);
    }
    qsort(ord, (size_t)n, sizeof(sidx_t), sidx_cmp);
    int nk = 0;
    for (int _i = 0; _i < n; ++_i) {
        int i = ord[_i].i;
        if (sup[i]) continue;
        keep_out[nk++] = i;
        const float* bi = boxes + (size_t)i * 4;
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = ord[_j].i;
            if (sup[j]) continue;
            float ovr = iou_f32(bi, boxes + (size_t)j * 4, area[i], area[j], o);
            if (ge_cmp ? ((double)ovr >= thr) : ((double)ovr > thr)) sup[j] = 1;
        }
    }
    free(ord); free(area); free(sup);
    return nk;
}

/* ====================================================================================
 * (a5) RRNet.nms default branch, models/rrnet.py:56-72: per class present in the image
 * (ascending, `unique()` :60) run torchvision-style NMS (thr given, normally 0.7) on the
 * rows of that class and concatenate.  dets [K,6]; out [<=K,6]; returns rows kept.
 * src_idx (optional) receives the source row of each kept row.
 * ================================================================================== */
ORC_API int orc_stage1_nms(const float* dets, int K, int num_classes, double thr,
                           float* out, int32_t* src_idx) {
    float* bx = (float*)malloc(sizeof(float) * 4 * (size_t)(K > 0 ? K : 1));
    float* sc = (float*)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
    int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
    int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
    int n_out = 0;
    for (int c = 0; c < num_classes; ++c) {
        int m = 0;
        for (int r = 0; r < K; ++r)
            if (dets[(size_t)r * 6 + 5] == (float)c) {
                memcpy(bx + (size_t)m * 4, dets + (size_t)r * 6, 16);
                sc[m] = dets[(size_t)r * 6 + 4];
                rows[m++] = r;
            }
        int nk = orc_nms(bx, sc, m, thr, 0, 0, keep);
        for (int t = 0; t < nk; ++t) {
            int r = rows[keep[t]];
            memcpy(out + (size_t)n_out * 6, dets + (size_t)r * 6, 24);
            if (src_idx) src_idx[n_out] = r;
            ++n_out;
        }
    }
    free(bx); free(sc); free(rows); free(keep);
    return n_out;
}

/* ====================================================================================
 * (a9) Soft-NMS, ext/nms/nms/cpu_nms.pyx:17-120 (cpu_soft_nms), in place on boxes [n,5]
 * (x1,y1,x2,y2,score).  method: 1 linear, 2 gaussian, else hard.  The gaussian weight is
 * exp() evaluated in double on the fp32 quotient -(ov*ov)/sigma and rounded to fp32
 * (:97 calls np.exp on a C float).  Returns N, the number of leading rows kept.
 * ================================================================================== */
ORC_API int orc_soft_nms(float* boxes, int n, float sigma, float Nt, float threshold, int method) {
    int N = n;
    for (int i = 0; i < N; ++i) {
        float* bi = boxes + (size_t)i * 5;
        float maxscore = bi[4];
        int maxpos = i;
        float tx1 = bi[0], ty1 = bi[1], tx2 = bi[2], ty2 = bi[3], ts = bi[4];
        for (int pos = i + 1; pos < N; ++pos)                         /* :46-50 */
            if (maxscore < boxes[(size_t)pos * 5 + 4]) { maxscore = boxes[(size_t)pos * 5 + 4]; maxpos = pos; }
        float* bm = boxes + (size_t)maxpos * 5;                       /* :53-64 swap */
        bi[0] = bm[0]; bi[1] = bm[1]; bi[2] = bm[2]; bi[3] = bm[3]; bi[4] = bm[4];
        bm[0] = tx1; bm[1] = ty1; bm[2] = tx2; bm[3] = ty2; bm[4] = ts;
        tx1 = bi[0]; ty1 = bi[1]; tx2 = bi[2]; ty2 = bi[3];
        int pos = i + 1;
        while (pos < N) {                                             /* :74-118 */
            float* bp = boxes + (size_t)pos * 5;
            float x1 = bp[0], y1 = bp[1], x2 = bp[2], y2 = bp[3];
            float area = (x2 - x1 + 1) * (y2 - y1 + 1);
            float iw = ((tx2 < x2 ? tx2 : x2) - (tx1 > x1 ? tx1 : x1) + 1);
            if (iw > 0) {
                float ih = ((ty2 < y2 ? ty2 : y2) - (ty1 > y1 ? ty1 : y1) + 1);
                if (ih > 0) {
                    float ua = (float)((tx2 - tx1 + 1) * (ty2 - ty1 + 1) + area - iw * ih);
                    float ov = iw * ih / ua;
                    float weight;
                    if (method == 1) weight = (ov > Nt) ? 1 - ov : 1;
                    else if (method == 2) weight = (float)exp((double)(-(ov * ov) / sigma));
                    else weight = (ov > Nt) ? 0 : 1;
                    bp[4] = weight * bp[4];
                    if (bp[4] < threshold) {                          /* :108-115 */
                        float* bl = boxes + (size_t)(N - 1) * 5;
                        bp[0] = bl[0]; bp[1] = bl[1]; bp[2] = bl[2]; bp[3] = bl[3]; bp[4] = bl[4];
                        N -= 1;
                        pos -= 1;
                    }
                }
            }
            pos += 1;
        }
    }
    return N;
}

/* ====================================================================================
 * (a6) torchvision.ops.roi_align(relu(feat), rois, (PH,PW)) as called at
 * models/rrnet.py:51: spatial_scale 1, sampling_ratio -1 (adaptive), aligned False.
 * Restated from torchvision's published CPU kernel (roi_align_kernel.cpp: pre-computed
 * bilinear taps, then per channel a running fp32 sum of w1*f1+w2*f2+w3*f3+w4*f4, divided by
 * the sample count).  feat [B,C,H,W]; rois [N,5] = (batch as float, x1,y1,x2,y2);
 * out [N,C,PH,PW].  relu != 0 applies max(v,0) to every tap (the reference feeds relu(feat)).
 * ================================================================================== */
typedef struct { int p1, p2, p3, p4; float w1, w2, w3, w4; } tap_t;

ORC_API int orc_roi_align(const float* feat, const float* rois, int N, int B, int C, int H, int W,
                          int PH, int PW, int relu, float* out) {
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int n = 0; n < N; ++n) {
        const float* r = rois + (size_t)n * 5;
        int bi = (int)r[0];
        float* o = out + (size_t)n * C * PH * PW;
        if (bi < 0 || bi >= B) { rc = -1; memset(o, 0, sizeof(float) * (size_t)C * PH * PW); continue; }
        float sw = r[1], sh = r[2], ew = r[3], eh = r[4];
        float rw = ew - sw, rh = eh - sh;
        if (!(rw > 1.0f)) rw = 1.0f;
        if (!(rh > 1.0f)) rh = 1.0f;
        float bh = rh / (float)PH, bw = rw / (float)PW;
        int gh = (int)ceilf(rh / (float)PH), gw = (int)ceilf(rw / (float)PW);
        int cnt_i = gh * gw; if (cnt_i < 1) cnt_i = 1;
        float count = (float)cnt_i;
        size_t ntap = (size_t)PH * PW * (gh > 0 ? gh : 0) * (gw > 0 ? gw : 0);
        tap_t* taps = (tap_t*)malloc(sizeof(tap_t) * (ntap ? ntap : 1));
        size_t t = 0;
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw)
                for (int iy = 0; iy < gh; ++iy) {
                    float yy = sh + ph * bh + (float)(iy + .5f) * bh / (float)gh;
                    for (int ix = 0; ix < gw; ++ix) {
                        float xx = sw + pw * bw + (float)(ix + .5f) * bw / (float)gw;
                        float x = xx, y = yy;
                        tap_t tp;
                        if (y < -1.0 || y > H || x < -1.0 || x > W) {
                            memset(&tp, 0, sizeof(tp));
                            taps[t++] = tp;
                            continue;
                        }
                        if (y <= 0) y = 0;
                        if (x <= 0) x = 0;
                        int yl = (int)y, xl = (int)x, yh, xh;
                        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                        float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
                        tp.w1 = hy * hx; tp.w2 = hy * lx; tp.w3 = ly * hx; tp.w4 = ly * lx;
                        tp.p1 = yl * W + xl; tp.p2 = yl * W + xh; tp.p3 = yh * W + xl; tp.p4 = yh * W + xh;
                        taps[t++] = tp;
                    }
                }
        for (int c = 0; c < C; ++c) {
            const float* f = feat + ((size_t)bi * C + c) * H * W;
            size_t ti = 0;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.f;
                    for (int s = 0; s < gh * gw; ++s) {
                        tap_t tp = taps[ti++];
                        float f1 = f[tp.p1], f2 = f[tp.p2], f3 = f[tp.p3], f4 = f[tp.p4];
                        if (relu) {
                            f1 = f1 > 0.f ? f1 : 0.f; f2 = f2 > 0.f ? f2 : 0.f;
                            f3 = f3 > 0.f ? f3 : 0.f; f4 = f4 > 0.f ? f4 : 0.f;
                        }
                        acc += tp.w1 * f1 + tp.w2 * f2 + tp.w3 * f3 + tp.w4 * f4;
                    }
                    acc /= count;
                    o[((size_t)c * PH + ph) * PW + pw] = acc;
                }
        }
        free(taps);
    }
    return rc;
}

/* ====================================================================================
 * (a7) Re-regression head in eval mode: detectors/fasterrcnn_detector.py:13-18 ->
 * backbones/resnet.py:33-53 (Bottleneck(256,64)) -> adaptive_avg_pool2d(1) -> 1x1 conv.
 * x [N,256,3,3].  Weights in the reference's state_dict layout:
 *   w1 [64,256] (conv1 1x1, no bias)   bn1 {gamma,beta,mean,var}[64]
 *   w2 [64,64,3,3] (conv2 pad 1)       bn2 [64]
 *   w3 [256,64]                         bn3 [256]
 *   wr [4,256], br [4]                  (regressor)
 * bn(z) = (z-mean)/sqrt(var+eps)*gamma+beta with eps=1e-5 (nn.BatchNorm2d default).
 * Sums are accumulated in double (the oracle is the accurate answer; the reference's fp32
 * mkldnn/cuDNN sums differ from it at ~1e-6 relative, inside the 1e-5 contract).
 * ================================================================================== */
static inline float bn_apply(double z, const float* bn, int c, int nch) {
    /* bn = [gamma | beta | mean | var], each nch long */
    double g = bn[c], b = bn[nch + c], m = bn[2 * nch + c], v = bn[3 * nch + c];
    return (float)((z - m) / sqrt(v + 1e-5) * g + b);
}

ORC_API int orc_head(const float* x, int N, const float* w1, const float* bn1,
                     const float* w2, const float* bn2, const float* w3, const float* bn3,
                     const float* wr, const float* br, float* out) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const float* xi = x + (size_t)n * 256 * 9;
        float t1[64 * 9], t2[64 * 9], pooled[256];
        for (int o = 0; o < 64; ++o)
            for (int p = 0; p < 9; ++p) {
                double acc = 0;
                for (int c = 0; c < 256; ++c) acc += (double)w1[o * 256 + c] * xi[c * 9 + p];
                float v = bn_apply(acc, bn1, o, 64);
                t1[o * 9 + p] = v > 0.f ? v : 0.f;
            }
        for (int o = 0; o < 64; ++o)
            for (int py = 0; py < 3; ++py)
                for (int px = 0; px < 3; ++px) {
                    double acc = 0;
                    for (int c = 0; c < 64; ++c)
                        for (int ky = 0; ky < 3; ++ky) {
                            int yy = py + ky - 1;
                            if (yy < 0 || yy > 2) continue;
                            for (int kx = 0; kx < 3; ++kx) {
                                int xx = px + kx - 1;
                                if (xx < 0 || xx > 2) continue;
                                acc += (double)w2[((o * 64 + c) * 3 + ky) * 3 + kx] * t1[c * 9 + yy * 3 + xx];
                            }
                        }
                    float v = bn_apply(acc, bn2, o, 64);
                    t2[o * 9 + py * 3 + px] = v > 0.f ? v : 0.f;
                }
        for (int o = 0; o < 256; ++o) {
            double ps = 0;
            for (int p = 0; p < 9; ++p) {
                double acc = 0;
                for (int c = 0; c < 64; ++c) acc += (double)w3[o * 64 + c] * t2[c * 9 + p];
                float v = bn_apply(acc, bn3, o, 256) + xi[o * 9 + p];     /* resnet.py:49 */
                v = v > 0.f ? v : 0.f;
                ps += v;
            }
            pooled[o] = (float)(ps / 9.0);
        }
        for (int j = 0; j < 4; ++j) {
            double acc = br[j];
            for (int c = 0; c < 256; ++c) acc += (double)wr[j * 256 + c] * pooled[c];
            out[(size_t)n * 4 + j] = (float)acc;
        }
    }
    return 0;
}

/* ====================================================================================
 * (a8) RRNetOperator.generate_bbox, operators/rrnet_operator.py:188-209, for the rows of
 * one image (bxyxy[:,0] == batch_idx).  s1 = [X1,Y1,w,h,score,0] built BEFORE the +1
 * (:196-198); s2 decodes with w+1,h+1 (:200-208), class emitted 1-based.
 * Returns the number of rows written to s1/s2 ([n,6] each).
 * ================================================================================== */
ORC_API int orc_generate_bbox(const float* bxyxy, const float* reg, const float* scores,
                              const float* clses, int N, int batch_idx, float scale,
                              float* s1, float* s2) {
    int m = 0;
    for (int i = 0; i < N; ++i) {
        const float* r = bxyxy + (size_t)i * 5;
        if (r[0] != (float)batch_idx) continue;
        float X1 = r[1] * scale, Y1 = r[2] * scale, X2 = r[3] * scale, Y2 = r[4] * scale;
        float w = X2 - X1, h = Y2 - Y1;
        float* a = s1 + (size_t)m * 6;
        a[0] = X1; a[1] = Y1; a[2] = w; a[3] = h; a[4] = scores[i]; a[5] = 0.f;
        float w1 = w + 1.f, h1 = h + 1.f;
        const float* g = reg + (size_t)i * 4;
        float cx = g[0] * w1 + X1 + w1 / 2.f;
        float cy = g[1] * h1 + Y1 + h1 / 2.f;
        float ow = expf(g[2]) * w1, oh = expf(g[3]) * h1;
        float* q = s2 + (size_t)m * 6;
        q[0] = cx - ow / 2.f; q[1] = cy - oh / 2.f; q[2] = ow; q[3] = oh;
        q[4] = scores[i]; q[5] = clses[i] + 1.f;
        ++m;
    }
    return m;
}

/* ====================================================================================
 * (a11) Target render: datasets/transforms/functional.py:177-262 (gaussian_radius,
 * gaussian2d, draw_umich_gaussian, to_heatmap).  annos [n,8] = x,y,w,h,score,cls,.. in
 * input pixels (cls 1-based); img_h,img_w input size; hm [cls_num, img_h/sf, img_w/sf]
 * must be zero-initialised by the caller (it is max-accumulated, so several calls can
 * splat into the same map).  Side outputs wh[n,2], ind[n], offset[n,2], reg_mask[n].
 * radius_out (optional) receives the per-object integer radius.
 * ================================================================================== */
static float gaussian_radius_f32(float height, float width) {
    /* functional.py:177-198 with min_overlap = 0.7; python scalars enter as fp32 */
    const float c_1m = (float)(1 - 0.7), c_1p = (float)(1 + 0.7);
    float b1 = height + width;
    float c1 = width * height * c_1m / c_1p;
    float sq1 = sqrtf(b1 * b1 - 4.0f * 1.0f * c1);       /* 4*a1*c1 with a1 = 1 */
    float r1 = (b1 + sq1) / 2.0f;
    float b2 = 2.0f * (height + width);
    float c2 = c_1m * width * height;
    float sq2 = sqrtf(b2 * b2 - 16.0f * c2);             /* 4*a2*c2 with a2 = 4 */
    float r2 = (b2 + sq2) / 2.0f;
    const float a3x4 = (float)(4 * (4 * 0.7));            /* 4*a3, a3 = 4*0.7 */
    float b3 = (float)(-2 * 0.7) * (height + width);
    float c3 = (float)(0.7 - 1) * width * height;
    float sq3 = sqrtf(b3 * b3 - a3x4 * c3);
    float r3 = (b3 + sq3) / 2.0f;
    float r = r1 < r2 ? r1 : r2;
    return r < r3 ? r : r3;
}

ORC_API int orc_render(const float* annos, int n, int img_h, int img_w, int scale_factor, int cls_num,
                       float* hm, float* wh, float* ind, float* offset, float* reg_mask,
                       float* radius_out) {
    const int Hh = img_h / scale_factor, Wh = img_w / scale_factor;
    const float sf = (float)scale_factor;
    for (int k = 0; k < n; ++k) {
        const float* a = annos + (size_t)k * 8;
        float x1 = a[0], y1 = a[1];
        float x2 = a[2] + a[0], y2 = a[3] + a[1];                 /* :246-247 */
        x1 = x1 / sf; y1 = y1 / sf; x2 = x2 / sf; y2 = y2 / sf;   /* :248 */
        float bh = y2 - y1, bw = x2 - x1;                         /* :250 */
        float cx = (x1 + x2) / 2.f, cy = (y1 + y2) / 2.f;         /* :254 */
        float cxi = floorf(cx), cyi = floorf(cy);
        if (wh) { wh[2 * k] = bw; wh[2 * k + 1] = bh; }
        if (offset) { offset[2 * k] = cx - cxi; offset[2 * k + 1] = cy - cyi; }
        if (reg_mask) reg_mask[k] = (bh > 0.f && bw > 0.f) ? 1.f : 0.f;
        if (ind) ind[k] = cyi * (float)(img_w / 4) + cxi;          /* :257 hard-coded 4 */
        float rad = floorf(gaussian_radius_f32(ceilf(bh), ceilf(bw)));
        if (!(rad > 0.f)) rad = 0.f;                              /* :259 clamp(min=0) */
        if (radius_out) radius_out[k] = rad;
        int cls = (int)(a[5] - 1.f);                              /* :249 */
        if (!hm) continue;
        if (cls < 0) cls += cls_num;                              /* python negative index */
        if (cls < 0 || cls >= cls_num) return -1;
        float* plane = hm + (size_t)cls * Hh * Wh;
        /* draw_umich_gaussian :212-227 */
        float diameter = 2.f * rad + 1.f;
        float sigma = diameter / 6.f;
        float left = cxi < rad ? cxi : rad;
        float right = ((float)Wh - cxi) < (rad + 1.f) ? ((float)Wh - cxi) : (rad + 1.f);
        float top = cyi < rad ? cyi : rad;
        float bottom = ((float)Hh - cyi) < (rad + 1.f) ? ((float)Hh - cyi) : (rad + 1.f);
        int ya = (int)(cyi - top), yb = (int)(cyi + bottom);
        int xa = (int)(cxi - left), xb = (int)(cxi + right);
        /* python slicing semantics: negative starts wrap, stops clip; an empty slice draws
         * nothing (:225).  Centres are inside the map on this path, so only clipping matters. */
        if (ya < 0 || xa < 0) continue;
        if (yb > Hh) yb = Hh;
        if (xb > Wh) xb = Wh;
        float denom = 2.f * sigma * sigma;                        /* :205 (2*sigma)*sigma */
        for (int y = ya; y < yb; ++y)
            for (int x = xa; x < xb; ++x) {
                float dx = (float)x - cxi, dy = (float)y - cyi;
                float g = expf(-(dx * dx + dy * dy) / denom);
                if (g > plane[(size_t)y * Wh + x]) plane[(size_t)y * Wh + x] = g;
            }
    }
    return 0;
}

/* ====================================================================================
 * (a12) Heat-map focal loss: modules/loss/functional.py:25-51 (focal_loss_for_hm) fed by
 * operators/rrnet_operator.py:55 (p = clamp(sigmoid(z), 1e-4, 1-1e-4)).
 * sums_out = {pos_loss_sum, neg_loss_sum, num_pos}; returns the loss in *loss_out;
 * grad_out (optional) = d loss / d z  (zero where sigmoid(z) is outside the clamp range,
 * as autograd through clamp gives).  Element terms are fp32, sums are accumulated in double.
 * ================================================================================== */
ORC_API int orc_focal(const float* logits, const float* gt, int64_t n, double* sums_out,
                      double* loss_out, float* grad_out) {
    const float lo = 1e-4f, hi = 1.0f - 1e-4f;
    double pos = 0, neg = 0, npos = 0;
#pragma omp parallel for reduction(+ : pos, neg, npos) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float s = sigmoidf_(logits[i]);
        float p = s < lo ? lo : (s > hi ? hi : s);
        float g = gt[i];
        if (g == 1.0f) {
            float q = 1.f - p;
            pos += (double)(logf(p) * (q * q));
            npos += 1;
        } else if (g < 1.0f) {
            float q = 1.f - g;
            float w = (q * q) * (q * q);
            neg += (double)(logf(1.f - p) * (p * p) * w);
        }
    }
    double loss = (npos == 0) ? -neg : -(pos + neg) / npos;
    if (sums_out) { sums_out[0] = pos; sums_out[1] = neg; sums_out[2] = npos; }
    if (loss_out) *loss_out = loss;
    if (grad_out) {
        const double scale = (npos == 0) ? -1.0 : -1.0 / npos;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            double s = 1.0 / (1.0 + exp(-(double)logits[i]));
            float sf = sigmoidf_(logits[i]);
            double d = 0;
            if (sf >= lo && sf <= hi) {          /* clamp passes gradient inside [lo,hi] */
                double p = s, g = gt[i];
                if (g == 1.0) d = (1 - p) * (1 - p) / p - 2 * (1 - p) * log(p);
                else if (g < 1.0) {
                    double q = 1 - g;
                    d = q * q * q * q * (2 * p * log(1 - p) - p * p / (1 - p));
                }
                d *= p * (1 - p);
            }
            grad_out[i] = (float)(d * scale);
        }
    }
    return 0;
}

/* ====================================================================================
 * Legacy ext/nms GPU-kernel layout helper (nms_kernel.cu:34-78,127-139): the 64-wide
 * suppression bit-mask of score-sorted boxes and the host greedy reduce, used to check the
 * legacy-ABI entry point bit-for-bit.  boxes [n,dim>=4] sorted by the caller.
 * ================================================================================== */
ORC_API int orc_nms_sorted(const float* boxes, int n, int dim, double thr, int pixel_offset,
                           int ge_cmp, int32_t* keep_out) {
    if (n <= 0) return 0;
    const float o = pixel_offset ? 1.0f : 0.0f;
    unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (sup[i]) continue;
        keep_out[nk++] = i;
        const float* a = boxes + (size_t)i * dim;
        float area_a = (a[2] - a[0] + o) * (a[3] - a[1] + o);
        for (int j = i + 1; j < n; ++j) {
            if (sup[j]) continue;
            const float* b = boxes + (size_t)j * dim;
            float area_b = (b[2] - b[0] + o) * (b[3] - b[1] + o);
            float ovr = iou_f32(a, b, area_a, area_b, o);
            if (ge_cmp ? ((double)ovr >= thr) : ((double)ovr > thr)) sup[j] = 1;
        }
    }
    free(sup);
    return nk;
}
