"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the reference's outputs; inputs are stored when small and re-generated
from `rrnet_b200.synth` seeds (with a sha1 of their bytes stored for drift detection) when
large.  The reference code is imported from /root/reference through tests/golden/ref_loader.py;
third-party arithmetic is torch 2.11.0 / torchvision 0.26.0 CPU (the container's versions, the
reference pins none).
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from rrnet_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402


def sha1(*tensors):
    h = hashlib.sha1()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy() if isinstance(t, torch.Tensor) else t).tobytes())
    return h.hexdigest()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-18s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


KNOWN_5 = np.array([[10, 9, 20, 19, 0.5], [10, 10, 15, 30, 0.45], [10, 10, 26, 26, 0.7],
                    [8, 9, 14, 16, 0.3], [8, 8, 15, 15, 0.1]], dtype=np.float32)


def gold_nms(ref):
    # (i) the reference's in-tree known answer, ext/nms/nms_wrapper.py:37-55
    keep_cpu = np.asarray(ref.cpu_nms.cpu_nms(KNOWN_5.copy(), 0.3), np.int64)
    keep_py = np.asarray(ref.py_cpu_nms(KNOWN_5.copy(), 0.3), np.int64)
    soft = ref.NW.soft_nms(KNOWN_5.copy(), Nt=0.4, sigma=0.3)
    assert keep_cpu.tolist() == [2, 3] and keep_py.tolist() == [2, 3] and soft.shape[0] == 5
    keep_tv = torchvision.ops.nms(torch.from_numpy(KNOWN_5[:, :4]), torch.from_numpy(KNOWN_5[:, 4]), 0.3).numpy()
    # (ii) clustered random boxes, all three semantics, plus degenerate / duplicate rows
    d = synth.nms_stress_boxes(700, 11).numpy()
    d[50:60, 2] = d[50:60, 0]          # zero width  -> area 0 (0/0 = NaN never suppresses when o=0)
    d[60:64, :4] = d[64:68, :4]        # exact duplicates with different scores
    d[70:74, 2:4] = d[70:74, 0:2] - 3  # x2<x1: negative extents
    res = {"known": KNOWN_5, "known_cpu_nms": keep_cpu, "known_py_cpu_nms": keep_py,
           "known_soft_rows": soft, "known_torchvision": keep_tv, "boxes": d}
    for thr in (0.3, 0.5, 0.7):
        t = "%02d" % int(thr * 10)
        res["tv_" + t] = torchvision.ops.nms(torch.from_numpy(d[:, :4].copy()), torch.from_numpy(d[:, 4].copy()), thr).numpy()
        res["cpu_" + t] = np.asarray(ref.cpu_nms.cpu_nms(d.copy(), thr), np.int64)
        res["py_" + t] = np.asarray(ref.py_cpu_nms(d.copy(), thr), np.int64)
    save("nms", **res)


def gold_soft_nms(ref):
    d = synth.nms_stress_boxes(400, 12).numpy()
    res = {"boxes": d}
    for method in (0, 1, 2):
        rows = ref.NW.soft_nms(d.copy(), sigma=0.5, Nt=0.7, threshold=0.1, method=method)
        res["rows_m%d" % method] = rows
    # as called by RRNetOperator._ext_nms (rrnet_operator.py:211-232): xywh rows with class column
    g = torch.Generator().manual_seed(13)
    cls = torch.randint(1, 4, (400, 1), generator=g).float()
    xywh = torch.from_numpy(d.copy())
    xywh[:, 2:4] -= xywh[:, 0:2]
    pred = torch.cat([xywh, cls], dim=1)
    res["ext_in"] = pred.clone()
    res["ext_out"] = ref.O.RRNetOperator._ext_nms(pred.clone())
    save("soft_nms", **res)


def gold_decode(ref):
    B, C, H, W, K = 2, 10, 40, 56, 100
    seed = 21
    hm = synth.heatmap_logits(B, C, H, W, K, seed)
    wh, off = synth.wh_offset(B, H, W, seed)
    net = ref.make_net()
    with torch.no_grad():
        scores, inds, clses, ys, xs = net._topk(torch.sigmoid(hm), K)
        dets = net.transform_bbox(hm, wh, off, K)
    save("decode", shape=np.array([B, C, H, W, K]), seed=seed, sha_in=sha1(hm, wh, off),
         dets=dets, inds=inds, clses=clses, ys=ys, xs=xs, scores=scores)


def gold_pipeline(ref):
    """RRNet.forward post-backbone (models/rrnet.py:25-54) + generate_bbox + _ext_nms."""
    B, C, H, W, K = 2, 10, 48, 64, 200
    seed = 31
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(seed)
    net = ref.make_net((x["hm"], x["wh"], x["off"]))
    missing = net.load_state_dict(synth.head_state_dict(hp), strict=False)
    assert not missing.unexpected_keys
    assert not [k for k in missing.missing_keys if k.startswith("head_detector") and "num_batches" not in k]
    with torch.no_grad():
        outs = net([x["feat"], x["feat"]], k=K)
        hms, whs, offs, s2_reg, bxyxy, scores, clses = outs
        roi_feat = torchvision.ops.roi_align(torch.relu(x["feat"]), bxyxy, (3, 3))
        res = dict(shape=np.array([B, C, H, W, K]), seed=seed,
                   sha_in=sha1(x["hm"], x["wh"], x["off"], x["feat"]),
                   sha_head=sha1(*[hp[k] for k in sorted(hp)]),
                   s2_reg=s2_reg, bxyxy=bxyxy, scores=scores, clses=clses,
                   roi_feat_every8=roi_feat[::8].contiguous(),
                   roi_feat_sum=roi_feat.double().sum(dim=(1, 2, 3)))
        for b in range(B):
            outs_c = (hms, whs, offs, s2_reg.clone(), bxyxy.clone(), scores.clone(), clses.clone())
            s1, s2 = ref.O.RRNetOperator.generate_bbox(ref.fake_op, outs_c, b)
            res["s1_b%d" % b] = s1
            res["s2_b%d" % b] = s2
            res["final_b%d" % b] = ref.O.RRNetOperator._ext_nms(s2.clone())
    save("pipeline", **res)


def gold_roi_align(ref):
    g = torch.Generator().manual_seed(41)
    B, C, H, W = 2, 8, 20, 24
    feat = torch.randn(B, C, H, W, generator=g)
    rois = [
        [0, 2.3, 3.1, 9.7, 12.2], [1, 0.0, 0.0, 24.0, 20.0], [0, -5.0, -4.0, 6.0, 5.0],
        [1, 18.5, 15.25, 30.0, 28.0], [0, 5.0, 5.0, 5.0, 5.0], [1, 7.0, 7.0, 7.4, 7.9],
        [0, -30.0, -30.0, -20.0, -20.0], [1, 30.0, 2.0, 40.0, 9.0], [0, 3.0, 2.0, 2.0, 1.0],
        [1, 0.5, 0.5, 3.49, 3.51], [0, 22.9, 18.9, 23.9, 19.9], [1, -1.0, -1.0, 2.0, 2.0],
        [0, 10.0, 1.0, 50.0, 19.0], [1, 1.0, 10.0, 4.0, 70.0],
    ]
    more = torch.rand(40, 4, generator=g) * torch.tensor([W * 1.2, H * 1.2, 14.0, 14.0]) - torch.tensor([2.0, 2.0, 0, 0])
    for i in range(40):
        x1, y1, w, h = more[i].tolist()
        rois.append([i % B, x1, y1, x1 + w, y1 + h])
    rois = torch.tensor(rois, dtype=torch.float32)
    out_relu = torchvision.ops.roi_align(torch.relu(feat), rois, (3, 3))
    out_raw = torchvision.ops.roi_align(feat, rois, (3, 3))
    save("roi_align", feat=feat, rois=rois, out_relu=out_relu, out_raw=out_raw)


def gold_head(ref):
    seed = 51
    hp = synth.head_params(seed)
    net = ref.make_net()
    net.load_state_dict(synth.head_state_dict(hp), strict=False)
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(24, 256, 3, 3, generator=g)) * 1.3
    with torch.no_grad():
        y = net.forward_stage2(x)
    save("head", seed=seed, x=x, y=y, sha_head=sha1(*[hp[k] for k in sorted(hp)]))


def gold_render(ref):
    res = {}
    # (ii) of SURVEY 8c: the demo annotation (class-0 rows dropped), 960x540 image
    ann_path = os.path.join(ref_loader.REF, "data/demo/annotations/0000364_01765_d_0000782.txt")
    rows = [[float(v) for v in ln.strip().rstrip(",").split(",")[:8]] for ln in open(ann_path) if ln.strip()]
    demo = torch.tensor(rows, dtype=torch.float32)
    demo = demo[demo[:, 5] != 0]
    img = torch.zeros(3, 540, 960)
    _, _, hm, wh, ind, off, msk = ref.TF.to_heatmap((img, demo.clone()))
    assert hm.shape == (10, 135, 240) and int((hm == 1).sum()) == 81
    res.update(demo_annos=demo, demo_hm=hm, demo_wh=wh, demo_ind=ind, demo_off=off,
               demo_mask=msk.float(), demo_sum=np.float64(hm.double().sum().item()))
    # random clustered objects incl. border cases, 2 images 256x320
    annos = synth.train_annos(2, 256, 320, 61, n_range=(30, 60))
    annos[0][0, :4] = torch.tensor([0.0, 0.0, 9.0, 7.0])          # top-left corner
    annos[0][1, :4] = torch.tensor([300.0, 240.0, 19.0, 15.0])    # bottom-right corner
    annos[1][0, :4] = torch.tensor([100.0, 100.0, 1.0, 1.0])      # tiny -> radius 0
    for b, a in enumerate(annos):
        _, _, hm, wh, ind, off, msk = ref.TF.to_heatmap((torch.zeros(3, 256, 320), a.clone()))
        res.update({"r%d_annos" % b: a, "r%d_hm" % b: hm, "r%d_wh" % b: wh, "r%d_ind" % b: ind,
                    "r%d_off" % b: off, "r%d_mask" % b: msk.float()})
    save("render", **res)


def gold_focal(ref):
    g = torch.Generator().manual_seed(71)
    annos = synth.train_annos(2, 128, 128, 72, n_range=(5, 12))
    gts = []
    for a in annos:
        gts.append(ref.TF.to_heatmap((torch.zeros(3, 128, 128), a.clone()))[2])
    gt = torch.stack(gts)                                   # [2,10,32,32]
    z = torch.randn(2, 10, 32, 32, generator=g) * 3.0 - 2.0
    z[0, 0, 0, :8] = torch.tensor([-12.0, -9.5, -9.3, -9.2, 9.2, 9.3, 9.5, 12.0])   # straddle the clamp
    z.requires_grad_(True)
    p = torch.clamp(torch.sigmoid(z), min=1e-4, max=1 - 1e-4)          # rrnet_operator.py:55
    loss = ref.LF.focal_loss_for_hm(p, gt)
    loss.backward()
    res = dict(logits=z.detach(), gt=gt, loss=np.float64(loss.item()), grad=z.grad.clone())
    # no-positive branch (functional.py:47-48)
    z2 = z.detach().clone().requires_grad_(True)
    gt0 = gt.clone()
    gt0[gt0 == 1] = 0.95
    p2 = torch.clamp(torch.sigmoid(z2), min=1e-4, max=1 - 1e-4)
    loss2 = ref.LF.focal_loss_for_hm(p2, gt0)
    loss2.backward()
    res.update(gt_nopos=gt0, loss_nopos=np.float64(loss2.item()), grad_nopos=z2.grad.clone())
    save("focal", **res)


def gold_criterion(ref):
    """RRNetOperator.criterion (rrnet_operator.py:42-84) on the pipeline inputs: the four losses and the
    gradient of (hm + 0.1 wh + off + s2) w.r.t. the stage-1 maps of stack 0."""
    B, C, H, W, K = 2, 10, 48, 64, 200
    seed = 31
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(seed)
    hm = x["hm"].clone().requires_grad_(True)
    wh = x["wh"].clone().requires_grad_(True)
    off = x["off"].clone().requires_grad_(True)
    net = ref.make_net((hm, wh, off))
    net.load_state_dict(synth.head_state_dict(hp), strict=False)
    annos = synth.train_annos(B, H * 4, W * 4, 81, n_range=(10, 25))
    # make a few ground-truth boxes coincide with stage-1 boxes so that the stage-2 loss has positives
    with torch.no_grad():
        dets = net.transform_bbox(x["hm"], x["wh"], x["off"], K)
    for b in range(B):
        for j in range(4):
            d = dets[b, 3 * j]
            x1, y1, x2, y2 = [float(v) * 4 for v in d[:4]]
            if x2 - x1 > 4 and y2 - y1 > 4 and x1 > 2 and y1 > 2 and x2 * 1.05 < W * 4 - 2 and y2 < H * 4 - 2:
                annos[b][j, :4] = torch.tensor([x1 + 0.5, y1 - 0.5, (x2 - x1) * 1.05, (y2 - y1) * 0.97])
    from datasets.drones_det import DronesDET
    batch = []
    for b, a in enumerate(annos):
        img, an, g_hm, g_wh, g_ind, g_off, g_msk = ref.TF.to_heatmap((torch.zeros(3, H * 4, W * 4), a.clone()))
        batch.append((img, an, g_hm, g_wh, g_ind, g_off, g_msk, "img%d" % b))
    imgs, c_annos, g_hms, g_whs, g_inds, g_offs, g_msks, _ = DronesDET.collate_fn_ctnet(batch)
    outs = net([x["feat"], x["feat"]], k=K)
    targets = (g_hms, g_whs, g_inds, g_offs, g_msks, c_annos.clone())
    hm_l, wh_l, off_l, s2_l = ref.O.RRNetOperator.criterion(ref.fake_op, outs, targets)
    total = hm_l + 0.1 * wh_l + off_l + s2_l
    total.backward()
    save("criterion", shape=np.array([B, C, H, W, K]), seed=seed, annos=c_annos, n_obj=np.array([a.shape[0] for a in annos]),
         gt_hms=g_hms, gt_whs=g_whs, gt_inds=g_inds, gt_offs=g_offs, gt_masks=g_msks,
         losses=np.array([float(hm_l), float(wh_l), float(off_l), float(s2_l)], np.float64),
         grad_hm=hm.grad, grad_wh=wh.grad, grad_off=off.grad)


def gold_ap_match(ref):
    """utils/metrics/metrics.py:get_tp of the reference on the seeded images of synth.AP_MATCH_CASES."""
    import utils.metrics.metrics as M
    thr = torch.arange(0.5, 1.0, 0.05)
    out = {"thresholds": thr}
    for k, case in enumerate(synth.AP_MATCH_CASES):
        pred, tgt = synth.ap_match_case(*case)
        flags = [torch.zeros(0, thr.numel()) for _ in range(10)]
        confs = [torch.zeros(0) for _ in range(10)]
        f, cf, tc, ii = M.get_tp(pred.clone(), tgt.clone(), flags, confs, torch.zeros(10), torch.zeros(10), thr)
        out["tp_%d" % k] = torch.cat(f)
        out["conf_%d" % k] = torch.cat(cf)
        out["sizes_%d" % k] = torch.tensor([x.shape[0] for x in f])
        out["target_count_%d" % k] = tc
        out["in_img_%d" % k] = ii
        out["sha_%d" % k] = np.frombuffer(bytes.fromhex(sha1(pred, tgt)), dtype=np.uint8)
    save("ap_match", **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = ref_loader.load()
    print("torch", torch.__version__, "torchvision", torchvision.__version__)
    gold_nms(ref)
    gold_soft_nms(ref)
    gold_decode(ref)
    gold_roi_align(ref)
    gold_head(ref)
    gold_pipeline(ref)
    gold_render(ref)
    gold_focal(ref)
    gold_criterion(ref)
    gold_ap_match(ref)


if __name__ == "__main__":
    main()
