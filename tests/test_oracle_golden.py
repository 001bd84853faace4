"""Pin the CPU oracle (oracle/rr_oracle.c) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py) and the reference's in-tree known answers.
CPU only -- runs in the authoring container and on the GPU box alike."""
import hashlib

import numpy as np
import pytest
import torch

from rrnet_b200 import synth
from tests.conftest import load_golden, rel_err

TOL = 1e-5


def sha1(*tensors):
    h = hashlib.sha1()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy() if isinstance(t, torch.Tensor) else t).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------ NMS (a5, a10)
def test_nms_known_answer(oracle_mod):
    g = load_golden("nms")
    k = g["known"]
    # ext/nms/nms_wrapper.py:54-56: nms(thresh=0.3) -> [2, 3]
    assert oracle_mod.nms(k[:, :4], k[:, 4], 0.3, pixel_offset=1, ge_cmp=True).tolist() == [2, 3]
    assert oracle_mod.nms(k[:, :4], k[:, 4], 0.3, pixel_offset=1, ge_cmp=False).tolist() == [2, 3]
    assert g["known_cpu_nms"].tolist() == [2, 3] and g["known_py_cpu_nms"].tolist() == [2, 3]
    # torchvision semantics differ on the same vector
    assert oracle_mod.nms(k[:, :4], k[:, 4], 0.3).tolist() == g["known_torchvision"].tolist() == [2, 1, 3]
    # ext/nms/nms_wrapper.py:47-50: soft_nms keeps all five
    rows = oracle_mod.soft_nms(k, sigma=0.3, Nt=0.4, threshold=0.001, method=1)
    assert rows.shape[0] == 5
    np.testing.assert_array_equal(rows, g["known_soft_rows"])


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
def test_nms_three_semantics(oracle_mod, thr):
    g = load_golden("nms")
    d = g["boxes"]
    t = "%02d" % int(thr * 10)
    assert oracle_mod.nms(d[:, :4], d[:, 4], thr, 0, False).tolist() == g["tv_" + t].tolist()
    assert oracle_mod.nms(d[:, :4], d[:, 4], thr, 1, True).tolist() == g["cpu_" + t].tolist()
    assert oracle_mod.nms(d[:, :4], d[:, 4], thr, 1, False).tolist() == g["py_" + t].tolist()


def test_nms_empty_and_single(oracle_mod):
    assert oracle_mod.nms(np.zeros((0, 4)), np.zeros((0,)), 0.5).tolist() == []
    assert oracle_mod.nms(np.array([[0, 0, 1, 1.0]]), np.array([0.3]), 0.5).tolist() == [0]


# ------------------------------------------------------------------ soft-NMS (a9)
@pytest.mark.parametrize("method", [0, 1, 2])
def test_soft_nms(oracle_mod, method):
    g = load_golden("soft_nms")
    rows = oracle_mod.soft_nms(g["boxes"], sigma=0.5, Nt=0.7, threshold=0.1, method=method)
    ref = g["rows_m%d" % method]
    assert rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, :4], ref[:, :4])
    assert rel_err(rows[:, 4], ref[:, 4]) < TOL


# ------------------------------------------------------------------ decode (a2, a3)
def test_decode(oracle_mod):
    g = load_golden("decode")
    B, C, H, W, K = g["shape"].tolist()
    seed = int(g["seed"])
    hm = synth.heatmap_logits(B, C, H, W, K, seed)
    wh, off = synth.wh_offset(B, H, W, seed)
    assert sha1(hm, wh, off) == str(g["sha_in"]), "synthetic generator drifted from the golden inputs"
    dets, inds, flat = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), K)
    np.testing.assert_array_equal(inds, g["inds"])                     # bit-exact top-K indices
    np.testing.assert_array_equal(flat // (H * W), g["clses"])
    np.testing.assert_array_equal(dets[..., 5], g["dets"][..., 5])
    np.testing.assert_array_equal(dets[..., :4], g["dets"][..., :4])   # box arithmetic is exact
    assert rel_err(dets[..., 4], g["dets"][..., 4]) < TOL              # sigmoid: 1e-5 rel


def test_decode_pool3_matches_torch(oracle_mod):
    """pool=3 follows operators/centernet_operator.py:204-210 (_ctnet_nms) restated with torch."""
    B, C, H, W, K = 1, 3, 17, 23, 40
    hm = synth.heatmap_logits(B, C, H, W, K, 5)
    wh, off = synth.wh_offset(B, H, W, 5)
    heat = torch.sigmoid(hm)
    keep = (torch.nn.functional.max_pool2d(heat, 3, stride=1, padding=1) == heat).float()
    ref_scores, ref_idx = torch.topk((heat * keep).view(B, -1), K)
    dets, inds, flat = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), K, pool=3)
    npos = int((ref_scores[0] > 0).sum())
    assert npos > 5
    np.testing.assert_array_equal(flat[0, :npos], ref_idx[0, :npos].numpy())
    assert rel_err(dets[0, :npos, 4], ref_scores[0, :npos].numpy()) < TOL
    assert (dets[0, npos:, 4] == 0).all()


def test_decode_rejects_bad_k(oracle_mod):
    z = np.zeros((1, 2, 4, 4), np.float32)
    with pytest.raises(ValueError):
        oracle_mod.decode(z, z[:, :2], z[:, :2], 17)


# ------------------------------------------------------------------ RoIAlign (a6)
def test_roi_align(oracle_mod):
    g = load_golden("roi_align")
    out = oracle_mod.roi_align(g["feat"], g["rois"], relu=True)
    assert rel_err(out, g["out_relu"]) < TOL
    out = oracle_mod.roi_align(g["feat"], g["rois"], relu=False)
    assert rel_err(out, g["out_raw"]) < TOL


# ------------------------------------------------------------------ head (a7)
def test_head(oracle_mod):
    g = load_golden("head")
    hp = synth.head_params(int(g["seed"]))
    assert sha1(*[hp[k] for k in sorted(hp)]) == str(g["sha_head"])
    y = oracle_mod.head(g["x"], {k: v.numpy() for k, v in hp.items()})
    assert rel_err(y, g["y"], floor=1.0) < TOL             # relative to max|y| (dot products cancel)


# ------------------------------------------------------------------ full eval path (a1..a9)
def test_pipeline(oracle_mod):
    g = load_golden("pipeline")
    B, C, H, W, K = g["shape"].tolist()
    seed = int(g["seed"])
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(seed)
    assert sha1(x["hm"], x["wh"], x["off"], x["feat"]) == str(g["sha_in"])
    dets, _, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    rows, scores, clses = [], [], []
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(dets[b], C, 0.7)
        rows.append(np.concatenate([np.full((kept.shape[0], 1), b, np.float32), kept[:, :4]], 1))
        scores.append(kept[:, 4])
        clses.append(kept[:, 5])
    bxyxy = np.concatenate(rows)
    scores = np.concatenate(scores)
    clses = np.concatenate(clses)
    np.testing.assert_array_equal(bxyxy, g["bxyxy"])                   # keep-lists + order bit-exact
    np.testing.assert_array_equal(clses, g["clses"])
    assert rel_err(scores, g["scores"]) < TOL
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy, relu=True)
    assert rel_err(roi[::8], g["roi_feat_every8"]) < TOL
    assert rel_err(roi.astype(np.float64).sum(axis=(1, 2, 3)), g["roi_feat_sum"], floor=1e-4) < 1e-4
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    assert rel_err(reg, g["s2_reg"], floor=1.0) < TOL       # sums with cancellation: relative to max|y|
    for b in range(B):
        s1, s2 = oracle_mod.generate_bbox(bxyxy, g["s2_reg"], g["scores"], clses, b, 4.0)
        assert rel_err(s1, g["s1_b%d" % b]) < TOL
        assert rel_err(s2, g["s2_b%d" % b]) < TOL


def test_ext_nms_final(oracle_mod):
    """RRNetOperator._ext_nms (rrnet_operator.py:211-232): per class soft-NMS of xywh rows."""
    g = load_golden("soft_nms")
    pred = g["ext_in"]
    outs = []
    for c in np.unique(pred[:, 5]):
        rows = pred[pred[:, 5] == c].copy()
        rows[:, 2] += rows[:, 0]
        rows[:, 3] += rows[:, 1]
        kept = oracle_mod.soft_nms(rows[:, :5], sigma=0.5, Nt=0.7, threshold=0.1, method=2)
        kept = np.concatenate([kept, np.full((kept.shape[0], 1), c, np.float32)], 1)
        kept[:, 2:4] -= kept[:, 0:2]
        outs.append(kept)
    out = np.concatenate(outs)
    assert out.shape == g["ext_out"].shape
    assert rel_err(out, g["ext_out"]) < TOL


# ------------------------------------------------------------------ render (a11)
def test_render_demo_known_answer(oracle_mod):
    g = load_golden("render")
    r = oracle_mod.render(g["demo_annos"], 540, 960)
    assert r["hm"].shape == (10, 135, 240)
    assert int((r["hm"] == 1).sum()) == 81
    assert abs(float(r["hm"].astype(np.float64).sum()) - 294.5373923947336) < 1e-3
    np.testing.assert_array_equal(r["hm"] == 1, g["demo_hm"] == 1)
    np.testing.assert_array_equal(r["hm"] > 0, g["demo_hm"] > 0)
    assert rel_err(r["hm"], g["demo_hm"]) < TOL
    np.testing.assert_array_equal(r["wh"], g["demo_wh"])
    np.testing.assert_array_equal(r["ind"], g["demo_ind"])
    np.testing.assert_array_equal(r["offset"], g["demo_off"])
    np.testing.assert_array_equal(r["reg_mask"], g["demo_mask"])


@pytest.mark.parametrize("b", [0, 1])
def test_render_random(oracle_mod, b):
    g = load_golden("render")
    r = oracle_mod.render(g["r%d_annos" % b], 256, 320)
    np.testing.assert_array_equal(r["hm"] > 0, g["r%d_hm" % b] > 0)
    np.testing.assert_array_equal(r["hm"] == 1, g["r%d_hm" % b] == 1)
    assert rel_err(r["hm"], g["r%d_hm" % b]) < TOL
    np.testing.assert_array_equal(r["wh"], g["r%d_wh" % b])
    np.testing.assert_array_equal(r["ind"], g["r%d_ind" % b])
    np.testing.assert_array_equal(r["offset"], g["r%d_off" % b])
    np.testing.assert_array_equal(r["reg_mask"], g["r%d_mask" % b])


# ------------------------------------------------------------------ focal loss (a12)
def test_focal(oracle_mod):
    g = load_golden("focal")
    loss, sums, grad = oracle_mod.focal(g["logits"], g["gt"], want_grad=True)
    assert sums[2] == float((g["gt"] == 1).sum())
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(grad, g["grad"], floor=1e-5) < 1e-4
    loss, sums, grad = oracle_mod.focal(g["logits"], g["gt_nopos"], want_grad=True)
    assert sums[2] == 0
    assert abs(loss - float(g["loss_nopos"])) / abs(float(g["loss_nopos"])) < TOL
    assert rel_err(grad, g["grad_nopos"], floor=1e-5) < 1e-4


# ------------------------------------------------------------------ evaluation: true-positive matching
def _sha(*tensors):
    h = hashlib.sha1()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return np.frombuffer(bytes.fromhex(h.hexdigest()), dtype=np.uint8)


@pytest.mark.parametrize("k", range(len(synth.AP_MATCH_CASES)))
def test_ap_match_oracle_vs_reference_get_tp(k):
    """oracle/metrics_np.get_tp_image against the outputs of the reference's own get_tp (utils/metrics/metrics.py:51-136)."""
    from oracle import metrics_np
    g = load_golden("ap_match")
    pred, tgt = synth.ap_match_case(*synth.AP_MATCH_CASES[k])
    np.testing.assert_array_equal(_sha(pred, tgt), g["sha_%d" % k])          # same bytes the reference saw
    o = metrics_np.get_tp_image(pred.numpy(), tgt.numpy(), g["thresholds"])
    np.testing.assert_array_equal(o["target_count"], g["target_count_%d" % k])
    np.testing.assert_array_equal(o["in_img"], g["in_img_%d" % k])
    tp, conf, sizes = [], [], []
    for c in range(1, 11):
        sel = (o["cls"] == c) & o["emit"]
        tp.append(o["tp"][sel]); conf.append(o["conf"][sel]); sizes.append(int(sel.sum()))
    np.testing.assert_array_equal(np.array(sizes), g["sizes_%d" % k])
    np.testing.assert_array_equal(np.concatenate(tp), g["tp_%d" % k])
    np.testing.assert_array_equal(np.concatenate(conf), g["conf_%d" % k])
