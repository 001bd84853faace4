"""Seeded synthetic inputs for the post-backbone path.

Everything is generated on the CPU with ``torch.Generator().manual_seed(seed)`` so the
oracle, the golden fixtures and the CUDA path all see identical bytes.  The heat-map
logits are *de-tied*: the top (K+64) logits of every image are strictly decreasing with
a gap that keeps their fp32 sigmoids strictly decreasing too, so "top-K indices
bit-exact" is well defined (torch.topk's tie order is arbitrary).
"""
import math

import torch

# Seeds fixed by SURVEY 8d.
SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 101, 202, 303, 404, 505

_GAP = 2e-5  # logit gap that keeps fp32 sigmoids distinct for logits <= ~4.5


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def heatmap_logits(B, C, H, W, K, seed, peak_frac=0.6):
    """[B,C,H,W] fp32 logits: background 1.5*N(0,1)-4 plus ~peak_frac*K planted 3x3 peaks per
    image at clustered centres (peak logit U(-1,4)), then de-tied over the top (K+64)."""
    g = _gen(seed)
    hm = torch.randn(B, C, H, W, generator=g) * 1.5 - 4.0
    hm.clamp_(max=3.0)
    P = max(1, int(peak_frac * K))
    n_clusters = max(1, P // 12)
    for b in range(B):
        cc = torch.rand(n_clusters, 2, generator=g) * torch.tensor([H - 1.0, W - 1.0])
        which = torch.randint(0, n_clusters, (P,), generator=g)
        jitter = torch.randn(P, 2, generator=g) * torch.tensor([H / 24.0 + 1.0, W / 24.0 + 1.0])
        ctr = (cc[which] + jitter).round()
        ys = ctr[:, 0].clamp(0, H - 1).long()
        xs = ctr[:, 1].clamp(0, W - 1).long()
        cls = torch.randint(0, C, (P,), generator=g)
        peak = torch.rand(P, generator=g) * 5.0 - 1.0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy = (ys + dy).clamp(0, H - 1)
                xx = (xs + dx).clamp(0, W - 1)
                fall = peak - 1.5 * (abs(dy) + abs(dx))
                cur = hm[b, cls, yy, xx]
                hm[b].index_put_((cls, yy, xx), torch.maximum(cur, fall))
    detie_(hm, K)
    return hm


def detie_(hm, K, extra=64):
    """In place: make each image's top (K+extra) logits strictly decreasing by >= _GAP (values
    are only ever raised, so they stay above everything outside the set), then check that the
    CPU fp32 sigmoid of the set is strictly decreasing as well."""
    B = hm.shape[0]
    flat = hm.view(B, -1)
    n = min(K + extra, flat.shape[1])
    for b in range(B):
        vals, idx = torch.topk(flat[b], n)
        # v'[i] = max(v[i], v'[i+1] + gap)  <=>  reverse running max of v[i] - (n-1-i)*gap
        ramp = torch.arange(n - 1, -1, -1, dtype=torch.float64) * _GAP
        w = vals.double() - ramp
        w = torch.flip(torch.cummax(torch.flip(w, [0]), 0).values, [0])
        v32 = (w + ramp).float()
        flat[b, idx] = v32
        s = torch.sigmoid(flat[b, idx])
        if not bool((v32[:-1] > v32[1:]).all()) or not bool((s[:-1] > s[1:]).all()):
            raise RuntimeError("de-tie failed: fp32 sigmoids not strictly decreasing (seed needs a resample)")
    return hm


def wh_offset(B, H, W, seed, neg_frac=0.02, wh_range=(1.5, 40.0)):
    """wh [B,2,H,W] = U(wh_range) in stride-4 units with neg_frac of entries negative
    (exercises clamp(min=0) -> degenerate boxes); offset [B,2,H,W] = U(0,1)."""
    g = _gen(seed + 1)
    wh = torch.rand(B, 2, H, W, generator=g) * (wh_range[1] - wh_range[0]) + wh_range[0]
    neg = torch.rand(B, 2, H, W, generator=g) < neg_frac
    wh = torch.where(neg, -wh, wh)
    off = torch.rand(B, 2, H, W, generator=g)
    return wh, off


def features(B, C, H, W, seed):
    """[B,C,H,W] = N(0,1): about half negative so the fused ReLU matters."""
    g = _gen(seed + 2)
    return torch.randn(B, C, H, W, generator=g)


def head_params(seed=0):
    """Re-regression head parameters in the reference's state_dict layout
    (detectors/fasterrcnn_detector.py:9-11, backbones/resnet.py:22-29), BN running stats
    randomised (mean N(0,.1), var U(.5,1.5)) so folding them is a real test."""
    g = _gen(seed + 3)

    def conv(o, i, k):
        fan = i * k * k
        return torch.randn(o, i, k, k, generator=g) * math.sqrt(2.0 / fan)

    def bn(ch):
        gamma = torch.rand(ch, generator=g) * 0.5 + 0.75
        beta = torch.randn(ch, generator=g) * 0.1
        mean = torch.randn(ch, generator=g) * 0.1
        var = torch.rand(ch, generator=g) + 0.5
        return torch.stack([gamma, beta, mean, var]).contiguous()

    return {
        "w1": conv(64, 256, 1).view(64, 256).contiguous(), "bn1": bn(64),
        "w2": conv(64, 64, 3).contiguous(), "bn2": bn(64),
        "w3": conv(256, 64, 1).view(256, 64).contiguous(), "bn3": bn(256),
        "wr": (torch.randn(4, 256, generator=g) * 0.05).contiguous(),
        "br": (torch.randn(4, generator=g) * 0.1).contiguous(),
    }


def head_state_dict(p, prefix="head_detector."):
    """Map head_params() to the reference's state_dict keys."""
    sd = {
        prefix + "top_layer.conv1.weight": p["w1"].view(64, 256, 1, 1),
        prefix + "top_layer.conv2.weight": p["w2"],
        prefix + "top_layer.conv3.weight": p["w3"].view(256, 64, 1, 1),
        prefix + "regressor.weight": p["wr"].view(4, 256, 1, 1),
        prefix + "regressor.bias": p["br"],
    }
    for i, k in ((1, "bn1"), (2, "bn2"), (3, "bn3")):
        sd[prefix + "top_layer.bn%d.weight" % i] = p[k][0]
        sd[prefix + "top_layer.bn%d.bias" % i] = p[k][1]
        sd[prefix + "top_layer.bn%d.running_mean" % i] = p[k][2]
        sd[prefix + "top_layer.bn%d.running_var" % i] = p[k][3]
    return sd


def eval_inputs(B, H, W, K, seed, C=10, feat_ch=256):
    """All inputs of the eval path for one batch: dict(hm, wh, off, feat)."""
    wh, off = wh_offset(B, H, W, seed)
    return {
        "hm": heatmap_logits(B, C, H, W, K, seed),
        "wh": wh, "off": off,
        "feat": features(B, feat_ch, H, W, seed),
    }


def train_annos(B, img_h, img_w, seed, n_range=(20, 150), num_classes=10):
    """Per-image annotation lists [n,8] = x,y,w,h,score,cls(1-based),trunc,occl in input pixels:
    n in U{n_range}, clustered centres, w,h in U(4,120) px, clipped to the image.  Objects whose
    gaussian radius falls within 1e-4 of an integer are resampled (torch's AVX sqrt is not
    correctly rounded, SURVEY B.6, so floor() there is machine dependent)."""
    g = _gen(seed)
    out = []
    for _ in range(B):
        n = int(torch.randint(n_range[0], n_range[1] + 1, (1,), generator=g))
        ncl = max(1, n // 10)
        cc = torch.rand(ncl, 2, generator=g) * torch.tensor([img_w * 1.0, img_h * 1.0])
        rows = []
        while len(rows) < n:
            c = cc[int(torch.randint(0, ncl, (1,), generator=g))]
            ctr = c + torch.randn(2, generator=g) * 40.0
            w, h = (torch.rand(2, generator=g) * 116.0 + 4.0).tolist()
            x = min(max(float(ctr[0]) - w / 2, 0.0), img_w - 2.0)
            y = min(max(float(ctr[1]) - h / 2, 0.0), img_h - 2.0)
            w = min(w, img_w - 1.0 - x)
            h = min(h, img_h - 1.0 - y)
            # integer pixel boxes like VisDrone annotations
            x, y, w, h = float(int(x)), float(int(y)), float(max(int(w), 1)), float(max(int(h), 1))
            cls = float(int(torch.randint(1, num_classes + 1, (1,), generator=g)))
            if _radius_near_integer(h / 4.0, w / 4.0):
                continue
            rows.append([x, y, w, h, 1.0, cls, 0.0, 0.0])
        out.append(torch.tensor(rows, dtype=torch.float32))
    return out


def _radius_near_integer(bh, bw, tol=1e-4):
    H, W = math.ceil(bh), math.ceil(bw)
    b1 = H + W
    r1 = (b1 + math.sqrt(max(b1 * b1 - 4 * W * H * 0.3 / 1.7, 0.0))) / 2
    b2 = 2 * (H + W)
    r2 = (b2 + math.sqrt(max(b2 * b2 - 16 * 0.3 * W * H, 0.0))) / 2
    b3 = -1.4 * (H + W)
    r3 = (b3 + math.sqrt(max(b3 * b3 + 11.2 * 0.3 * W * H, 0.0))) / 2
    r = min(r1, r2, r3)
    return abs(r - round(r)) < tol


def pad_annos(annos_list):
    """The reference collate layout (datasets/drones_det.py:70-94): zero-padded [B,max_n,8] + counts."""
    B = len(annos_list)
    max_n = max(a.shape[0] for a in annos_list)
    out = torch.zeros(B, max_n, 8)
    cnt = torch.zeros(B, dtype=torch.int32)
    for i, a in enumerate(annos_list):
        out[i, : a.shape[0]] = a[:, :8]
        cnt[i] = a.shape[0]
    return out, cnt


def nms_stress_boxes(n, seed, cluster=20):
    """n boxes in n/cluster clusters, centre jitter N(0,6), wh U(16,96)*U(.8,1.25), scores a random
    permutation of n distinct values -> dets [n,5] (x1,y1,x2,y2,score)."""
    g = _gen(seed)
    ncl = max(1, n // cluster)
    span = 60.0 * math.sqrt(ncl)
    cc = torch.rand(ncl, 2, generator=g) * span
    base = torch.rand(ncl, 2, generator=g) * 80.0 + 16.0
    which = torch.randint(0, ncl, (n,), generator=g)
    ctr = cc[which] + torch.randn(n, 2, generator=g) * 6.0
    wh = base[which] * (torch.rand(n, 2, generator=g) * 0.45 + 0.8)
    scores = (torch.randperm(n, generator=g).float() + 1.0) / (n + 1.0)
    x1y1 = ctr - wh / 2
    x2y2 = ctr + wh / 2
    return torch.cat([x1y1, x2y2, scores[:, None]], dim=1).contiguous()


AP_MATCH_CASES = ((30, 20, 2, 1), (60, 40, 0, 2), (5, 3, 1, 3), (80, 100, 4, 4), (200, 150, 6, 5), (1, 0, 0, 6))


def ap_match_case(n_gt, n_fp, n_ign, seed):
    """One image for the evaluation's true-positive matching (utils/metrics/metrics.py:get_tp): ground truth
    [n,6] = x,y,w,h,1,cls (cls 0 = ignore region) and detections [m,6] = x,y,w,h,score,cls: two jittered copies of every
    box (10 % with a wrong class), n_fp random false positives, distinct scores."""
    g = torch.Generator().manual_seed(7000 + seed)
    xy = torch.rand(n_gt, 2, generator=g) * 400
    wh = torch.rand(n_gt, 2, generator=g) * 60 + 8
    cls = torch.randint(1, 11, (n_gt,), generator=g).float()
    tgt = torch.cat([xy, wh, torch.ones(n_gt, 1), cls[:, None]], 1)
    if n_ign:
        ixy = torch.rand(n_ign, 2, generator=g) * 300
        iwh = torch.rand(n_ign, 2, generator=g) * 120 + 60
        tgt = torch.cat([tgt, torch.cat([ixy, iwh, torch.zeros(n_ign, 1), torch.zeros(n_ign, 1)], 1)])
        tgt = tgt[torch.randperm(tgt.shape[0], generator=g)]
    rep = torch.randint(0, n_gt, (2 * n_gt,), generator=g)
    jit = (torch.rand(2 * n_gt, 4, generator=g) - 0.5) * torch.tensor([12.0, 12.0, 10.0, 10.0])
    det = torch.cat([xy[rep], wh[rep]], 1) + jit
    det[:, 2:] = det[:, 2:].clamp(min=2)
    dcls = cls[rep].clone()
    flip = torch.rand(2 * n_gt, generator=g) < 0.1
    dcls[flip] = torch.randint(1, 11, (int(flip.sum()),), generator=g).float()
    fp = torch.cat([torch.rand(n_fp, 2, generator=g) * 400, torch.rand(n_fp, 2, generator=g) * 60 + 8], 1)
    boxes = torch.cat([det, fp])
    c = torch.cat([dcls, torch.randint(1, 11, (n_fp,), generator=g).float()])
    score = torch.randperm(boxes.shape[0], generator=g).float() / boxes.shape[0] * 0.98 + 0.01
    return torch.cat([boxes, score[:, None], c[:, None]], 1).contiguous(), tgt.contiguous()
