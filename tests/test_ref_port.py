"""The torch-CPU restatement of the reference's call sequence (oracle/ref_port.py, the timed CPU
baseline) reproduces the goldens the unmodified reference produced.  CPU only."""
import numpy as np
import torch

from oracle import build_ref, ref_port
from rrnet_b200 import synth
from tests.conftest import load_golden, rel_err


def test_ref_port_pipeline_matches_reference_golden():
    g = load_golden("pipeline")
    B, C, H, W, K = g["shape"].tolist()
    seed = int(g["seed"])
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(seed)
    r = ref_port.post_backbone(x["hm"], x["wh"], x["off"], x["feat"], hp, K)
    np.testing.assert_array_equal(r["bxyxy"].numpy(), g["bxyxy"])
    np.testing.assert_array_equal(r["scores"].numpy(), g["scores"])
    np.testing.assert_array_equal(r["clses"].numpy(), g["clses"])
    assert rel_err(r["reg"].numpy(), g["s2_reg"], floor=1.0) < 1e-6
    base = 0
    for b in range(B):
        n = r["counts"][b]
        assert rel_err(r["s1"][base:base + n].numpy(), g["s1_b%d" % b]) < 1e-6
        assert rel_err(r["s2"][base:base + n].numpy(), g["s2_b%d" % b], floor=1e-3) < 1e-5
        base += n
    mod = build_ref.load()
    if mod is not None:
        n0 = r["counts"][0]
        final = ref_port.final_soft_nms(torch.from_numpy(g["s2_b0"]), mod)
        assert final.shape == g["final_b0"].shape
        assert rel_err(final.numpy(), g["final_b0"], floor=1e-3) < 1e-5


def test_ref_port_decode_matches_reference_golden():
    g = load_golden("decode")
    B, C, H, W, K = g["shape"].tolist()
    hm = synth.heatmap_logits(B, C, H, W, K, int(g["seed"]))
    wh, off = synth.wh_offset(B, H, W, int(g["seed"]))
    rows, ind = ref_port.topk_decode(hm, wh, off, K)
    np.testing.assert_array_equal(rows.numpy(), g["dets"])
    np.testing.assert_array_equal(ind.numpy(), g["inds"])


def test_ref_port_focal_matches_reference_golden():
    g = load_golden("focal")
    loss = ref_port.focal_loss(torch.from_numpy(g["logits"]), torch.from_numpy(g["gt"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-6 * abs(float(g["loss"]))
