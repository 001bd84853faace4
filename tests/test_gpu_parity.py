"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.
Bar: bit-exact top-K indices, keep-lists, classes and box arithmetic; 1e-5 relative (fp32) for
scores, RoI features, head outputs and losses.  Run on the B200 box:  pytest -m gpu"""
import numpy as np
import pytest
import torch

from rrnet_b200 import synth
from tests.conftest import box_rel_err, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rrnet_b200 import ops as _ops
    _ops._lib.lib()
    return _ops


def dev(t):
    return (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).cuda()


def npy(t):
    return t.detach().cpu().numpy()


# ===================================================================== decode (a2-a4)
@pytest.mark.parametrize("B,C,H,W,K,seed", [
    (2, 10, 40, 56, 100, 21),       # the golden case
    (1, 10, 128, 128, 500, 101),    # config 1
    (3, 7, 33, 47, 64, 5),          # odd sizes: C*H*W not a multiple of 4 -> scalar path
    (1, 2, 8, 8, 64, 9),            # K == H*W (largest legal K)
    (2, 10, 272, 480, 1500, 202),   # config 2 shape, sample-threshold fast path
])
def test_decode_vs_oracle(ops, oracle_mod, B, C, H, W, K, seed):
    hm = synth.heatmap_logits(B, C, H, W, K, seed)
    wh, off = synth.wh_offset(B, H, W, seed)
    dets, inds = ops.decode_topk(dev(hm), dev(wh), dev(off), K)
    o_dets, o_inds, o_flat = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), K)
    np.testing.assert_array_equal(npy(inds), o_inds)
    d = npy(dets)
    np.testing.assert_array_equal(d[..., 5], o_dets[..., 5])
    np.testing.assert_array_equal(d[..., :4], o_dets[..., :4])
    assert rel_err(d[..., 4], o_dets[..., 4]) < TOL


def test_decode_golden(ops):
    g = load_golden("decode")
    B, C, H, W, K = g["shape"].tolist()
    hm = synth.heatmap_logits(B, C, H, W, K, int(g["seed"]))
    wh, off = synth.wh_offset(B, H, W, int(g["seed"]))
    dets, inds = ops.decode_topk(dev(hm), dev(wh), dev(off), K)
    np.testing.assert_array_equal(npy(inds), g["inds"])
    np.testing.assert_array_equal(npy(dets)[..., :4], g["dets"][..., :4])
    np.testing.assert_array_equal(npy(dets)[..., 5], g["dets"][..., 5])
    assert rel_err(npy(dets)[..., 4], g["dets"][..., 4]) < TOL


@pytest.mark.parametrize("case", ["constant", "two_level", "dense_hits"])
def test_decode_ties_and_fallback(ops, oracle_mod, case):
    """Massive ties / threshold-estimate failures take the exact in-CTA path; the canonical order
    (logit desc, flat index asc) must still match the oracle exactly."""
    B, C, H, W, K = 2, 4, 96, 120, 300           # N = 46080 > candidate capacity -> sampling active
    g = torch.Generator().manual_seed(3)
    if case == "constant":
        hm = torch.full((B, C, H, W), -1.25)
    elif case == "two_level":
        hm = torch.where(torch.rand(B, C, H, W, generator=g) < 0.4, torch.tensor(2.0), torch.tensor(-3.0))
    else:                                         # > capacity elements above any sampled threshold
        hm = torch.randn(B, C, H, W, generator=g).round()
    wh, off = synth.wh_offset(B, H, W, 3)
    dets, inds = ops.decode_topk(dev(hm), dev(wh), dev(off), K)
    o_dets, o_inds, _ = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), K)
    np.testing.assert_array_equal(npy(inds), o_inds)
    np.testing.assert_array_equal(npy(dets)[..., [0, 1, 2, 3, 5]], o_dets[..., [0, 1, 2, 3, 5]])


@pytest.mark.parametrize("B,H,W,K", [(2, 64, 96, 700), (1, 272, 480, 5000), (3, 40, 40, 1600)])
def test_decode_cluster_selection_matches_single_cta(ops, oracle_mod, B, H, W, K):
    """The cluster selection (8 CTAs per image: sorted runs exchanged through distributed shared memory, rank by
    bisection) and the one-CTA-per-image selection write identical rows; K = H*W hits the exact fall-back in both."""
    x = synth.eval_inputs(B, H, W, K, seed=B * 7 + K)
    hm, wh, off = dev(x["hm"]), dev(x["wh"]), dev(x["off"])
    d_c, i_c = ops.decode_topk(hm, wh, off, K)
    ops.set_option(ops.OPT_SELECT_SINGLE_CTA, 1)
    try:
        d_s, i_s = ops.decode_topk(hm, wh, off, K)
    finally:
        ops.set_option(ops.OPT_SELECT_SINGLE_CTA, 0)
    np.testing.assert_array_equal(npy(i_c), npy(i_s))
    np.testing.assert_array_equal(npy(d_c), npy(d_s))
    o_dets, o_inds, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    np.testing.assert_array_equal(npy(i_c), o_inds)


def test_decode_pool3(ops, oracle_mod):
    B, C, H, W, K = 2, 10, 64, 80, 120
    hm = synth.heatmap_logits(B, C, H, W, K, 8)
    wh, off = synth.wh_offset(B, H, W, 8)
    dets, inds = ops.decode_topk(dev(hm), dev(wh), dev(off), K, pool=3)
    o_dets, o_inds, _ = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), K, pool=3)
    np.testing.assert_array_equal(npy(inds), o_inds)
    np.testing.assert_array_equal(npy(dets)[..., [0, 1, 2, 3, 5]], o_dets[..., [0, 1, 2, 3, 5]])
    assert rel_err(npy(dets)[..., 4], o_dets[..., 4]) < TOL


def test_decode_rejects_bad_arguments(ops):
    z = torch.zeros(1, 2, 4, 4).cuda()
    with pytest.raises(ops.RRNetB200Error):
        ops.decode_topk(z, z, z, 17)              # K > H*W: torch.topk raises in the reference
    with pytest.raises(ops.RRNetB200Error):
        ops.decode_topk(z, z, z, 4, pool=5)


# ===================================================================== hard NMS (a5, a10)
def test_nms_known_answer_all_entry_points(ops):
    g = load_golden("nms")
    k = g["known"]
    bx, sc = dev(k[:, :4].copy()), dev(k[:, 4].copy())
    assert npy(ops.nms(bx, sc, 0.3, 1, True)).tolist() == [2, 3]       # cpu_nms semantics
    assert npy(ops.nms(bx, sc, 0.3, 1, False)).tolist() == [2, 3]      # gpu_nms / py_cpu_nms
    assert npy(ops.nms(bx, sc, 0.3, 0, False)).tolist() == [2, 1, 3]   # torchvision
    order = np.argsort(-k[:, 4], kind="stable")
    keep = ops.nms_legacy_host(k[order], 0.3)                           # `_nms` ABI: sorted rows
    assert order[keep].tolist() == [2, 3]


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
def test_nms_three_semantics_golden(ops, thr):
    g = load_golden("nms")
    d = g["boxes"]
    bx, sc = dev(d[:, :4].copy()), dev(d[:, 4].copy())
    t = "%02d" % int(thr * 10)
    assert npy(ops.nms(bx, sc, thr, 0, False)).tolist() == g["tv_" + t].tolist()
    assert npy(ops.nms(bx, sc, thr, 1, True)).tolist() == g["cpu_" + t].tolist()
    assert npy(ops.nms(bx, sc, thr, 1, False)).tolist() == g["py_" + t].tolist()


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1500, 5000])
def test_nms_sizes_vs_oracle(ops, oracle_mod, n):
    d = synth.nms_stress_boxes(n, 100 + n).numpy()
    for po, ge in ((0, False), (1, True)):
        keep = npy(ops.nms(dev(d[:, :4].copy()), dev(d[:, 4].copy()), 0.7, po, ge))
        assert keep.tolist() == oracle_mod.nms(d[:, :4], d[:, 4], 0.7, po, ge).tolist()


def test_nms_score_ties_are_stable(ops, oracle_mod):
    d = synth.nms_stress_boxes(400, 77).numpy()
    d[:, 4] = np.round(d[:, 4] * 8) / 8                                 # heavy score ties
    keep = npy(ops.nms(dev(d[:, :4].copy()), dev(d[:, 4].copy()), 0.5))
    assert keep.tolist() == oracle_mod.nms(d[:, :4], d[:, 4], 0.5).tolist()


def test_nms_batched_segments(ops, oracle_mod):
    sizes = [0, 5, 130, 0, 64, 700, 1]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    d = synth.nms_stress_boxes(int(offs[-1]), 55).numpy()
    keep_idx, keep_cnt = ops.nms_batched(dev(d[:, :4].copy()), dev(d[:, 4].copy()), dev(offs), 0.6)
    keep_idx, keep_cnt = npy(keep_idx), npy(keep_cnt)
    for s, n in enumerate(sizes):
        a, b = offs[s], offs[s + 1]
        ref = oracle_mod.nms(d[a:b, :4], d[a:b, 4], 0.6) + a
        assert keep_cnt[s] == len(ref)
        assert keep_idx[a:a + keep_cnt[s]].tolist() == ref.tolist()


def test_nms_empty(ops):
    e = torch.zeros(0, 4).cuda()
    assert ops.nms(e, torch.zeros(0).cuda(), 0.5).numel() == 0
    assert len(ops.nms_legacy_host(np.zeros((0, 5), np.float32), 0.5)) == 0


def test_nms_legacy_host_matches_reference_layout(ops, oracle_mod):
    d = synth.nms_stress_boxes(3000, 9).numpy()
    order = np.argsort(-d[:, 4], kind="stable")
    rows = d[order]
    keep = ops.nms_legacy_host(rows, 0.7)
    assert keep.tolist() == oracle_mod.nms_sorted(rows, float(np.float32(0.7)), 1, False).tolist()


def test_nms_idempotent_20k(ops):
    """Full-size stress (config 5): NMS of the kept set keeps everything, kept boxes are pairwise
    below the threshold, and every dropped box overlaps a higher-scored kept box."""
    d = synth.nms_stress_boxes(20000, synth.SEED_C5)
    bx, sc = d[:, :4].cuda().contiguous(), d[:, 4].cuda().contiguous()
    keep = ops.nms(bx, sc, 0.7, 1, False)
    assert 0 < keep.numel() < 20000
    again = ops.nms(bx[keep], sc[keep], 0.7, 1, False)
    assert again.tolist() == list(range(keep.numel()))
    assert bool((sc[keep][:-1] >= sc[keep][1:]).all())


# ===================================================================== stage-1 NMS (a5 + forward loop)
@pytest.mark.parametrize("B,H,W,K,seed", [(2, 48, 64, 200, 31), (3, 64, 96, 700, 12)])
def test_stage1_nms_vs_oracle(ops, oracle_mod, B, H, W, K, seed):
    C = 10
    hm = synth.heatmap_logits(B, C, H, W, K, seed)
    wh, off = synth.wh_offset(B, H, W, seed, wh_range=(1.5, 14.0))
    dets, _ = ops.decode_topk(dev(hm), dev(wh), dev(off), K)
    bxyxy, scores, clses, counts = ops.stage1_nms(dets, C, 0.7)
    counts = npy(counts)
    d = npy(dets)
    base = 0
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(d[b], C, 0.7)
        n = kept.shape[0]
        assert counts[b] == n
        np.testing.assert_array_equal(npy(bxyxy[base:base + n, 1:]), kept[:, :4])
        assert (npy(bxyxy[base:base + n, 0]) == b).all()
        np.testing.assert_array_equal(npy(scores[base:base + n]), kept[:, 4])
        np.testing.assert_array_equal(npy(clses[base:base + n]), kept[:, 5])
        base += n
    assert counts[B] == base
    assert base < B * K                                   # something was actually suppressed


def test_stage1_single_class_worst_case(ops, oracle_mod):
    """All K boxes in one class (one long segment spanning many 64-wide tiles)."""
    K = 1500
    d = synth.nms_stress_boxes(K, 4).numpy()
    order = np.argsort(-d[:, 4], kind="stable")
    dets = np.concatenate([d[order], np.full((K, 1), 3, np.float32)], 1)[None]
    bxyxy, scores, clses, counts = ops.stage1_nms(dev(dets.copy()), 10, 0.7)
    kept, _ = oracle_mod.stage1_nms(dets[0], 10, 0.7)
    n = int(npy(counts)[0])
    assert n == kept.shape[0]
    np.testing.assert_array_equal(npy(bxyxy[:n, 1:]), kept[:, :4])


# ===================================================================== RoIAlign (a6)
ALGOS = [0, 1, 2]   # 0 = tile-centric kernel, TMA-staged tiles when W % 4 == 0 (+ direct fallback), 1 = direct gather for every RoI,
                    # 2 = tile-centric kernel with tiles staged by ordinary loads


@pytest.mark.parametrize("algo", ALGOS)
def test_roi_align_golden_edge_cases(ops, algo):
    g = load_golden("roi_align")
    out = ops.roi_align(dev(g["feat"]), dev(g["rois"]), relu=True, algo=algo)
    assert rel_err(npy(out), g["out_relu"], floor=1e-3) < TOL
    out = ops.roi_align(dev(g["feat"]), dev(g["rois"]), relu=False, algo=algo)
    assert rel_err(npy(out), g["out_raw"], floor=1.0) < TOL          # signed taps cancel: relative to max


def _random_rois(n, B, H, W, seed, wh_max=39.0):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * torch.tensor([W * 1.1, H * 1.1]) - 3.0
    wh = torch.rand(n, 2, generator=g) * wh_max + 0.2
    rois = torch.cat([torch.randint(0, B, (n, 1), generator=g).float(), xy, xy + wh], 1)
    rois[::17, 3:] = rois[::17, 1:3]                                  # degenerate (w = h = 0)
    return rois


@pytest.mark.parametrize("algo", ALGOS)
def test_roi_align_vs_oracle_random(ops, oracle_mod, algo):
    B, C, H, W = 2, 256, 48, 64
    feat = synth.features(B, C, H, W, 3)
    rois = _random_rois(300, B, H, W, 4)
    out = npy(ops.roi_align(dev(feat), dev(rois), algo=algo))
    ref = oracle_mod.roi_align(feat.numpy(), rois.numpy(), relu=True)
    assert rel_err(out, ref, floor=1e-3) < TOL


def test_roi_align_tile_path_mixed_sizes(ops, oracle_mod):
    """Odd map size (partial edge tiles, W not a multiple of 32), RoIs from sub-pixel to larger than the
    64-pixel tile-path limit (those take the direct path inside the same call), unsorted image indices."""
    B, C, H, W = 3, 64, 77, 150
    feat = synth.features(B, C, H, W, 9)
    rois = _random_rois(500, B, H, W, 10, wh_max=90.0)
    rois[5] = torch.tensor([1.0, -200.0, -200.0, -150.0, -150.0])      # entirely outside -> zeros
    rois[6] = torch.tensor([7.0, 1.0, 1.0, 9.0, 9.0])                  # invalid image index -> zeros
    rois[7] = torch.tensor([0.0, 31.5, 23.5, 32.5, 24.5])              # straddles a tile corner
    rois[8] = torch.tensor([2.0, 0.0, 0.0, 63.0, 63.0])                # widest tile-path window
    out = ops.roi_align(dev(feat), dev(rois), algo=0)
    ok = np.ones(500, bool); ok[6] = False
    ref = oracle_mod.roi_align(feat.numpy(), rois.numpy()[ok], relu=True)
    assert rel_err(npy(out)[ok], ref, floor=1e-3) < TOL
    assert float(out[6].abs().max()) == 0.0 and float(out[5].abs().max()) == 0.0
    # the two algorithms agree to rounding and each is deterministic run to run
    out1 = ops.roi_align(dev(feat), dev(rois), algo=1)
    assert rel_err(npy(out), npy(out1), floor=1e-3) < TOL
    assert torch.equal(out, ops.roi_align(dev(feat), dev(rois), algo=0))


@pytest.mark.parametrize("relu", [True, False])
def test_roi_align_tma_tiles_edge_tiles_and_long_lists(ops, oracle_mod, relu):
    """W % 4 == 0 -> the TMA-staged tile kernel: partial edge tiles (zero filled by the TMA), W not a multiple of
    the tile width, tiles with more than 32 pieces (several work items per tile), every column alignment of a
    unit inside its 16-byte chunks; against the oracle, against the load-staged tile kernel, run to run."""
    B, C, H, W = 3, 64, 77, 152
    feat = synth.features(B, C, H, W, 21)
    rois = _random_rois(900, B, H, W, 22, wh_max=60.0)
    rois[:300, 1:3] = rois[:300, 1:3] * 0.2 + 30.0                       # crowd one corner: long piece lists
    rois[:300, 3:] = rois[:300, 1:3] + (rois[:300, 3:] - rois[:300, 3:].floor()) * 25.0 + 1.0
    rois[7] = torch.tensor([0.0, 31.5, 23.5, 32.5, 24.5])                # straddles a tile corner
    rois[8] = torch.tensor([2.0, 88.0, 13.0, 151.9, 76.9])               # reaches the map's last column and row
    rois[9] = torch.tensor([1.0, 140.0, 70.0, 170.0, 90.0])              # hangs over the map edge
    out = ops.roi_align(dev(feat), dev(rois), relu=relu, algo=0)
    ref = oracle_mod.roi_align(feat.numpy(), rois.numpy(), relu=relu)
    assert rel_err(npy(out), ref, floor=1e-3 if relu else 1.0) < TOL
    out2 = ops.roi_align(dev(feat), dev(rois), relu=relu, algo=2)
    assert rel_err(npy(out), npy(out2), floor=1e-3 if relu else 1.0) < TOL
    assert torch.equal(out, ops.roi_align(dev(feat), dev(rois), relu=relu, algo=0))


def test_roi_align_channels_not_multiple_of_32_and_slot_overflow(ops, oracle_mod):
    feat = synth.features(1, 8, 40, 50, 11)                               # C=8: every RoI takes the direct path
    rois = _random_rois(64, 1, 40, 50, 12)
    out = npy(ops.roi_align(dev(feat), dev(rois), algo=0))
    assert rel_err(out, oracle_mod.roi_align(feat.numpy(), rois.numpy(), relu=True), floor=1e-3) < TOL
    # 60x60 RoIs over a tiny-tile grid: 9-12 pieces each, more than the 4-per-RoI slot budget allows ->
    # the RoIs past the budget fall back to the direct path, results unchanged
    feat = synth.features(1, 32, 96, 128, 13)
    g = torch.Generator().manual_seed(14)
    xy = torch.rand(200, 2, generator=g) * torch.tensor([60.0, 30.0]) + 1.0
    rois = torch.cat([torch.zeros(200, 1), xy, xy + 60.0], 1)
    out = npy(ops.roi_align(dev(feat), dev(rois), algo=0))
    assert rel_err(out, oracle_mod.roi_align(feat.numpy(), rois.numpy(), relu=True), floor=1e-3) < TOL


@pytest.mark.parametrize("algo", ALGOS)
def test_roi_align_device_count_and_huge_roi(ops, oracle_mod, algo):
    B, C, H, W = 1, 8, 120, 150
    feat = synth.features(B, C, H, W, 6)
    rois = torch.tensor([[0, -10.0, -10.0, 160.0, 130.0], [0, 3.0, 4.0, 140.0, 9.0], [0, 1.0, 1.0, 5.0, 5.0]])
    n_dev = torch.tensor([2], dtype=torch.int32).cuda()
    out = ops.roi_align(dev(feat), dev(rois), n_dev=n_dev, algo=algo)
    ref = oracle_mod.roi_align(feat.numpy(), rois.numpy()[:2], relu=True)
    assert rel_err(npy(out[:2]), ref, floor=1e-3) < TOL


# ===================================================================== head (a7)
@pytest.mark.parametrize("algo", [0, 1])          # 0 = tcgen05 (3xTF32), 1 = fp32 FFMA
def test_head_golden_and_oracle(ops, oracle_mod, algo):
    g = load_golden("head")
    hp = synth.head_params(int(g["seed"]))
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    y = npy(ops.head_forward(dev(g["x"]), folded, algo=algo))
    assert rel_err(y, g["y"], floor=1.0) < TOL
    x = torch.relu(torch.randn(333, 256, 3, 3, generator=torch.Generator().manual_seed(1))) * 2
    y = npy(ops.head_forward(dev(x), folded, algo=algo))
    ref = oracle_mod.head(x.numpy(), {k: v.numpy() for k, v in hp.items()})
    assert rel_err(y, ref, floor=1.0) < TOL

@pytest.mark.parametrize("n", [1, 7, 8, 9, 127, 1184, 1185, 2501])
def test_head_tile_counts_and_device_count(ops, oracle_mod, n):
    """The tensor-core head is persistent: 148 CTAs walk over tiles of 8 RoIs.  RoI counts around the tile size,
    around one tile per CTA (148 * 8 = 1184) and several tiles per CTA; both kernels against the oracle, and the
    device-side row count (rows past it must stay untouched)."""
    hp = synth.head_params(5)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    gen = torch.Generator().manual_seed(100 + n)
    x = torch.relu(torch.randn(n, 256, 3, 3, generator=gen)) * 3 - 0.25     # a few negative inputs too
    ref = oracle_mod.head(x.numpy(), {k: v.numpy() for k, v in hp.items()})
    for algo in (0, 1):
        y = npy(ops.head_forward(dev(x), folded, algo=algo))
        assert rel_err(y, ref, floor=1.0) < TOL
    live = max(n - 3, 0)
    n_dev = torch.tensor([live], dtype=torch.int32, device="cuda")
    L = ops._lib.lib()
    reg = torch.full((n, 4), 12345.0, device="cuda")
    xd = dev(x)
    ops.check(L.rr_head_forward(xd.data_ptr(), n_dev.data_ptr(), n, folded.data_ptr(), 0, reg.data_ptr(),
                                torch.cuda.current_stream().cuda_stream), "rr_head_forward")
    reg = npy(reg)
    if live:
        assert rel_err(reg[:live], ref[:live], floor=1.0) < TOL
    assert (reg[live:] == 12345.0).all()


# ===================================================================== stage-1 head tail fused with decode (f4)
def _tail_inputs(B, Cin, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.relu(torch.randn(B, Cin, H, W, generator=g))                      # output of the head's 3x3 conv + ReLU
    w = torch.randn(C, Cin, 1, 1, generator=g) * (2.0 / Cin ** 0.5)
    b = torch.full((C,), -2.19) + 0.1 * torch.randn(C, generator=g)             # centernet_detector.py:18 bias init
    return t, w, b


@pytest.mark.parametrize("B,Cin,C,H,W,K", [
    (2, 256, 10, 48, 64, 200),       # the reference's head: 256 -> 10
    (1, 256, 10, 128, 128, 500),     # config-1 size
    (3, 40, 2, 13, 11, 50),          # H*W % 4 != 0 (scalar loads), Cin % 8 != 0, planes = 2 (the offset head's shape)
    (1, 64, 16, 20, 36, 300),        # the widest supported head, K close to H*W
    (2, 32, 7, 8, 8, 64),            # whole image fits the candidate list: no threshold
])
def test_hm_tail_collect_vs_reference_conv_and_decode(ops, oracle_mod, B, Cin, C, H, W, K):
    """rr_hm_tail_collect: logits against the reference layer itself (nn.Conv2d(256, planes, 1) on the CPU, fp32) at 1e-5,
    and the decode that starts from its candidate lists bit-exact against the oracle applied to the logits it wrote."""
    t, w, b = _tail_inputs(B, Cin, C, H, W, seed=B * 1000 + C)
    conv = torch.nn.Conv2d(Cin, C, (1, 1))
    with torch.no_grad():
        conv.weight.copy_(w); conv.bias.copy_(b)
        ref = conv(t).numpy()
    hm, ws = ops.hm_tail_collect(dev(t), dev(w), dev(b), K)
    # a logit is a 256-term sum that partly cancels (|logit| <= ~6, sum of |terms| ~ 10): the fp32 rounding error scales
    # with the terms, not with the result, so it is measured against max(|ref|, 0.1 max|ref|).  The CPU convolution sums in
    # blocked order, this kernel in channel order - like cuDNN's fp32 kernel, against which it is bit-identical
    assert rel_err(npy(hm), ref, floor=0.1) < TOL
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref_gpu = torch.nn.functional.conv2d(dev(t), dev(w), dev(b))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel_err(npy(hm), npy(ref_gpu), floor=0.1) < 2e-6
    x = synth.wh_offset(B, H, W, seed=7)
    dets, inds = ops.decode_topk(hm, dev(x[0]), dev(x[1]), K, precollected_ws=ws)
    o_dets, o_inds, _ = oracle_mod.decode(npy(hm), x[0].numpy(), x[1].numpy(), K)
    np.testing.assert_array_equal(npy(inds), o_inds)
    np.testing.assert_array_equal(npy(dets)[..., [0, 1, 2, 3, 5]], o_dets[..., [0, 1, 2, 3, 5]])
    assert rel_err(npy(dets)[..., 4], o_dets[..., 4]) < TOL
    # and it is what the unfused decode makes of the same map
    d2, i2 = ops.decode_topk(hm, dev(x[0]), dev(x[1]), K)
    np.testing.assert_array_equal(npy(i2), npy(inds))
    np.testing.assert_array_equal(npy(d2), npy(dets))


def test_hm_tail_rejects_bad_arguments(ops):
    t, w, b = _tail_inputs(1, 32, 17, 8, 8, seed=1)
    with pytest.raises(Exception):
        ops.hm_tail_collect(dev(t), dev(w), dev(b), 10)            # more than 16 output planes
    t, w, b = _tail_inputs(1, 32, 4, 8, 8, seed=1)
    with pytest.raises(Exception):
        ops.hm_tail_collect(dev(t), dev(w[:, :16]), dev(b), 10)    # weight does not match t


def test_eval_path_from_tail_full_size(ops):
    """Config-2 size: the fused entry (tail -> selection -> NMS -> RoIAlign -> head -> boxes) gives bit-identical
    results to the unfused path fed with the logit map the tail wrote."""
    B, C, H, W, K = 8, 10, 272, 480, 1500
    t, w, b = _tail_inputs(B, 256, C, H, W, seed=99)
    x = synth.eval_inputs(B, H, W, K, synth.SEED_C2)
    hp = synth.head_params(synth.SEED_C2)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    td, wd, bd = dev(t), dev(w), dev(b)
    del t
    whd, offd, featd = dev(x["wh"]), dev(x["off"]), dev(x["feat"])
    fused = ops.EvalPath(B, C, H, W, K, folded)
    hm = fused.forward_from_tail(td, wd, bd, whd, offd, featd)
    rf = fused.results()
    ref = torch.nn.functional.conv2d(td.cpu(), w, b).numpy()
    assert rel_err(npy(hm), ref, floor=0.1) < TOL
    plain = ops.EvalPath(B, C, H, W, K, folded)
    plain.forward(hm, whd, offd, featd)
    rp = plain.results()
    assert rf["n"] == rp["n"] and rf["counts"] == rp["counts"] and rf["n"] > 1000
    for key in ("bxyxy", "scores", "clses", "reg", "s1", "s2"):
        np.testing.assert_array_equal(npy(rf[key]), npy(rp[key]))
    np.testing.assert_array_equal(npy(fused.inds), npy(plain.inds))


# ===================================================================== whole eval path (a1..a8)
def test_eval_path_golden(ops):
    g = load_golden("pipeline")
    B, C, H, W, K = g["shape"].tolist()
    seed = int(g["seed"])
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(seed)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    path.forward(dev(x["hm"]), dev(x["wh"]), dev(x["off"]), dev(x["feat"]))
    r = path.results()
    assert r["n"] == g["bxyxy"].shape[0]
    np.testing.assert_array_equal(npy(r["bxyxy"]), g["bxyxy"])          # keep-lists and order bit-exact
    np.testing.assert_array_equal(npy(r["clses"]), g["clses"])
    assert rel_err(npy(r["scores"]), g["scores"]) < TOL
    roi = npy(path.roi_feat[: r["n"]])
    assert rel_err(roi[::8], g["roi_feat_every8"], floor=1e-3) < TOL
    assert rel_err(npy(r["reg"]), g["s2_reg"], floor=1.0) < TOL
    base = 0
    for b in range(B):
        n = r["counts"][b]
        assert rel_err(npy(r["s1"][base:base + n]), g["s1_b%d" % b]) < TOL
        assert box_rel_err(npy(r["s2"][base:base + n]), g["s2_b%d" % b]) < TOL
        base += n


def test_eval_path_config1_vs_oracle(ops, oracle_mod):
    """Config 1: B=1, 10x128x128, K=500 -- every stage against the oracle."""
    B, C, H, W, K = 1, 10, 128, 128, 500
    x = synth.eval_inputs(B, H, W, K, synth.SEED_C1)
    hp = synth.head_params(synth.SEED_C1)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    path.forward(dev(x["hm"]), dev(x["wh"]), dev(x["off"]), dev(x["feat"]))
    r = path.results()
    dets, _, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    kept, _ = oracle_mod.stage1_nms(dets[0], C, 0.7)
    bxyxy = np.concatenate([np.zeros((kept.shape[0], 1), np.float32), kept[:, :4]], 1)
    np.testing.assert_array_equal(npy(r["bxyxy"]), bxyxy)
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy, relu=True)
    assert rel_err(npy(path.roi_feat[: r["n"]]), roi, floor=1e-3) < TOL
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    assert rel_err(npy(r["reg"]), reg, floor=1.0) < TOL
    s1, s2 = oracle_mod.generate_bbox(bxyxy, npy(r["reg"]), kept[:, 4], kept[:, 5], 0, 4.0)
    assert rel_err(npy(r["s1"]), s1) < TOL
    assert box_rel_err(npy(r["s2"]), s2) < TOL


def test_eval_path_full_size_properties(ops, oracle_mod):
    """Config 2 (B=8, 10x272x480, K=1500): exact decode + keep-lists for every image, RoI features and
    head outputs of EVERY RoI (~11 k), and batch independence (image b alone == image b in batch)."""
    B, C, H, W, K = 8, 10, 272, 480, 1500
    x = synth.eval_inputs(B, H, W, K, synth.SEED_C2)
    hp = synth.head_params(synth.SEED_C2)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    xd = {k: dev(v) for k, v in x.items()}
    path.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    r = path.results()
    dets, inds, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    np.testing.assert_array_equal(npy(path.inds), inds)
    np.testing.assert_array_equal(npy(path.dets)[..., [0, 1, 2, 3, 5]], dets[..., [0, 1, 2, 3, 5]])
    # scores sorted descending per image
    s = npy(path.dets)[..., 4]
    assert (s[:, :-1] >= s[:, 1:]).all()
    rows = []
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(dets[b], C, 0.7)
        assert r["counts"][b] == kept.shape[0]
        rows.append(np.concatenate([np.full((kept.shape[0], 1), b, np.float32), kept[:, :4]], 1))
    bxyxy = np.concatenate(rows)
    np.testing.assert_array_equal(npy(r["bxyxy"]), bxyxy)
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy, relu=True)               # all RoIs
    assert rel_err(npy(path.roi_feat)[: r["n"]], roi, floor=1e-3) < TOL
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    # head outputs are O(1) regression deltas: error relative to max|reg| (absolute 1e-5 on the deltas), see conftest.rel_err
    assert rel_err(npy(r["reg"]), reg, floor=1.0) < TOL
    del roi
    # the fused form (head sums the RoIAlign partial slots itself, no RoI feature tensor) is bit-identical
    fused = ops.EvalPath(B, C, H, W, K, folded)
    fused.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    rf = fused.results()
    assert rf["n"] == r["n"]
    np.testing.assert_array_equal(npy(rf["reg"]), npy(r["reg"]))
    np.testing.assert_array_equal(npy(rf["s2"]), npy(r["s2"]))
    # the FFMA head agrees with the tensor-core head to rounding
    ffma = ops.EvalPath(B, C, H, W, K, folded, head_algo=1)
    ffma.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    assert rel_err(npy(ffma.results()["reg"]), npy(r["reg"]), floor=1.0) < TOL
    # the direct-gather RoIAlign agrees to rounding
    direct = ops.EvalPath(B, C, H, W, K, folded, roi_algo=1)
    direct.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    assert rel_err(npy(direct.results()["reg"]), npy(r["reg"]), floor=1.0) < TOL
    # image 5 alone gives exactly the rows it had inside the batch
    one = ops.EvalPath(1, C, H, W, K, folded)
    one.forward(xd["hm"][5:6], xd["wh"][5:6], xd["off"][5:6], xd["feat"][5:6])
    r1 = one.results()
    lo = sum(r["counts"][:5])
    hi = lo + r["counts"][5]
    assert r1["n"] == hi - lo
    np.testing.assert_array_equal(npy(r1["bxyxy"])[:, 1:], npy(r["bxyxy"])[lo:hi, 1:])
    np.testing.assert_array_equal(npy(r1["reg"]), npy(r["reg"])[lo:hi])
    np.testing.assert_array_equal(npy(r1["s2"]), npy(r["s2"])[lo:hi])


def _check_shard_against_oracle(ops, oracle_mod, B, K, seed, roi_stride):
    """One rank's shard of an image-sharded run: decode indices / boxes / classes and stage-1 keep-lists bit-exact
    for every image, RoI features + head + generate_bbox for every `roi_stride`-th RoI, through the fused path."""
    C, H, W = 10, 272, 480
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(synth.SEED_C2)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    path.forward(dev(x["hm"]), dev(x["wh"]), dev(x["off"]), dev(x["feat"]))
    r = path.results()
    dets, inds, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    np.testing.assert_array_equal(npy(path.inds), inds)
    np.testing.assert_array_equal(npy(path.dets)[..., [0, 1, 2, 3, 5]], dets[..., [0, 1, 2, 3, 5]])
    assert rel_err(npy(path.dets)[..., 4], dets[..., 4]) < TOL
    rows, sc, cl = [], [], []
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(dets[b], C, 0.7)
        assert r["counts"][b] == kept.shape[0]
        rows.append(np.concatenate([np.full((kept.shape[0], 1), b, np.float32), kept[:, :4]], 1))
        sc.append(kept[:, 4]); cl.append(kept[:, 5])
    bxyxy, sc, cl = np.concatenate(rows), np.concatenate(sc), np.concatenate(cl)
    np.testing.assert_array_equal(npy(r["bxyxy"]), bxyxy)
    np.testing.assert_array_equal(npy(r["clses"]), cl)
    sub = np.arange(0, bxyxy.shape[0], roi_stride)
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy[sub], relu=True)
    assert rel_err(npy(path.roi_feat)[sub], roi, floor=1e-3) < TOL
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    assert rel_err(npy(r["reg"])[sub], reg, floor=1.0) < TOL
    for b in (0, B - 1):
        sb = sub[bxyxy[sub, 0] == b]
        s1, s2 = oracle_mod.generate_bbox(bxyxy[sb], npy(r["reg"])[sb], sc[sb], cl[sb], b, 4.0)
        assert rel_err(npy(r["s1"])[sb], s1) < TOL
        assert box_rel_err(npy(r["s2"])[sb], s2) < TOL
    # what the rank contributes to the all-gather: padded rows + counts in one blob
    blob = npy(path.result_blob)
    np.testing.assert_array_equal(blob[B * K * 6:].view(np.int32)[:B], np.asarray(r["counts"], np.int32))
    assert int(blob[B * K * 6:].view(np.int32)[B]) == r["n"]
    return r


@pytest.mark.parametrize("world,rank", [(8, 0), (8, 3), (8, 7), (4, 1), (2, 1)])
def test_eval_path_config4_shards(ops, oracle_mod, world, rank):
    """Config 4 as BASELINE.json states it: batch 64 split 32 / 16 / 8 per GPU on 2 / 4 / 8 GPUs, seeds 404 + rank
    (SURVEY 8d); every shard shape, one or more ranks each."""
    B = 64 // world
    _check_shard_against_oracle(ops, oracle_mod, B, 1500, synth.SEED_C4 + rank, roi_stride=1 if world == 8 else 7)


def test_eval_path_combine_in_tile_kernel_is_bit_identical(ops):
    """RR_OPT_COMBINE_IN_TILE_KERNEL: the RoIAlign tile kernel sums a RoI's partial slots itself when its last piece is done
    (arrival counters, combine jobs) and hands the head one row per RoI.  Same order of summation as the head's own sum of
    the slots, so every output is bit-identical - also on a workspace full of garbage (a RoI without pieces has no row)."""
    B, C, H, W, K = 4, 10, 136, 240, 900
    x = {k: dev(v) for k, v in synth.eval_inputs(B, H, W, K, 4242).items()}
    folded = ops.head_fold({k: v.cuda() for k, v in synth.head_params(3).items()})
    ref = ops.EvalPath(B, C, H, W, K, folded)
    ref.forward(x["hm"], x["wh"], x["off"], x["feat"])
    r = ref.results()
    ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, 1)
    try:
        for fill in (0, 0x7f):
            p = ops.EvalPath(B, C, H, W, K, folded)
            p.ws.fill_(fill)
            for _ in range(2):                      # a second call on the used workspace
                p.forward(x["hm"], x["wh"], x["off"], x["feat"])
            got = p.results()
            assert got["n"] == r["n"] and got["n"] > 500
            for key in ("bxyxy", "reg", "s1", "s2"):
                np.testing.assert_array_equal(npy(got[key]), npy(r[key]))
    finally:
        ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, 0)


def test_eval_path_atomic_rows_within_tolerance(ops):
    """RR_OPT_COMBINE_IN_TILE_KERNEL = 2: no partial slots, the tile kernel's units add their bins into the RoI's zeroed row
    with float reductions.  The order of the additions is not fixed, so RoIs cut into three or more pieces may differ in the
    last bits from run to run: the results are held to 1e-6 of the default path (not bit-exact), boxes of the first stage and
    the kept set are untouched, and the FFMA head reads the same rows."""
    B, C, H, W, K = 4, 10, 136, 240, 900
    x = {k: dev(v) for k, v in synth.eval_inputs(B, H, W, K, 4242).items()}
    folded = ops.head_fold({k: v.cuda() for k, v in synth.head_params(3).items()})
    ref = ops.EvalPath(B, C, H, W, K, folded)
    ref.forward(x["hm"], x["wh"], x["off"], x["feat"])
    r = ref.results()
    ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, 2)
    try:
        for fill, algo in ((0, 0), (0x7f, 0), (0x7f, 1)):          # garbage workspace; tcgen05 head and FFMA head
            p = ops.EvalPath(B, C, H, W, K, folded, head_algo=algo)
            p.ws.fill_(fill)
            for _ in range(2):
                p.forward(x["hm"], x["wh"], x["off"], x["feat"])
            got = p.results()
            assert got["n"] == r["n"] and got["n"] > 500
            np.testing.assert_array_equal(npy(got["bxyxy"]), npy(r["bxyxy"]))
            np.testing.assert_array_equal(npy(got["s1"]), npy(r["s1"]))
            assert rel_err(npy(got["reg"]), npy(r["reg"]), floor=1.0) < (1e-6 if algo == 0 else TOL)
            assert box_rel_err(npy(got["s2"]), npy(r["s2"])) < TOL
    finally:
        ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, 0)


def test_eval_path_config5_full_batch(ops, oracle_mod):
    """Config 5 at its full size: B=16, K=5000 proposals per image (~75 k RoIs) through the whole path."""
    r = _check_shard_against_oracle(ops, oracle_mod, 16, 5000, synth.SEED_C5, roi_stride=5)
    assert r["n"] > 16 * 4000



def test_eval_path_config5_dense_scene(ops, oracle_mod):
    """Config 5 shape (K=5000 proposals per image at 10x272x480; 2 of the 16 images so that the oracle stays
    within seconds): decode and keep-lists bit-exact, RoI features / head / boxes for a strided subset."""
    B, C, H, W, K = 2, 10, 272, 480, 5000
    x = synth.eval_inputs(B, H, W, K, synth.SEED_C5)
    hp = synth.head_params(synth.SEED_C5)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    xd = {k: dev(v) for k, v in x.items()}
    path.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    r = path.results()
    dets, inds, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    np.testing.assert_array_equal(npy(path.inds), inds)
    np.testing.assert_array_equal(npy(path.dets)[..., [0, 1, 2, 3, 5]], dets[..., [0, 1, 2, 3, 5]])
    rows, sc, cl = [], [], []
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(dets[b], C, 0.7)
        assert r["counts"][b] == kept.shape[0]
        rows.append(np.concatenate([np.full((kept.shape[0], 1), b, np.float32), kept[:, :4]], 1))
        sc.append(kept[:, 4]); cl.append(kept[:, 5])
    bxyxy = np.concatenate(rows)
    np.testing.assert_array_equal(npy(r["bxyxy"]), bxyxy)
    np.testing.assert_array_equal(npy(r["clses"]), np.concatenate(cl))
    sub = np.arange(0, bxyxy.shape[0], 53)
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy[sub], relu=True)
    assert rel_err(npy(path.roi_feat)[sub], roi, floor=1e-3) < TOL
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    assert rel_err(npy(r["reg"])[sub], reg, floor=1.0) < TOL
    sub0 = sub[bxyxy[sub, 0] == 0]
    s1, s2 = oracle_mod.generate_bbox(bxyxy[sub0], npy(r["reg"])[sub0], np.concatenate(sc)[sub0],
                                      np.concatenate(cl)[sub0], 0, 4.0)
    assert rel_err(npy(r["s1"])[sub0], s1) < TOL
    assert box_rel_err(npy(r["s2"])[sub0], s2) < TOL
    # final stage at this density: per-class Gaussian soft-NMS of one image's ~4.7k stage-2 boxes vs the oracle
    n0 = r["counts"][0]
    s2_img = r["s2"][:n0]
    order = torch.sort(s2_img[:, 5], stable=True).indices
    d = s2_img[order]
    _, cnts = torch.unique_consecutive(d[:, 5], return_counts=True)
    seg = torch.zeros(cnts.numel() + 1, dtype=torch.int32, device=d.device)
    seg[1:] = torch.cumsum(cnts, 0).int()
    b5 = d[:, :5].clone()
    b5[:, 2:4] += b5[:, 0:2]
    out, _, kc = ops.soft_nms_batched(b5, seg, 0.5, 0.7, 0.1, 2)
    out, kc, seg_h = npy(out), npy(kc), npy(seg)
    for s_i in range(len(kc)):
        ref = oracle_mod.soft_nms(npy(b5)[seg_h[s_i]:seg_h[s_i + 1]], sigma=0.5, Nt=0.7, threshold=0.1, method=2)
        assert kc[s_i] == ref.shape[0]
        np.testing.assert_array_equal(out[seg_h[s_i]:seg_h[s_i] + kc[s_i], :4], ref[:, :4])
        assert rel_err(out[seg_h[s_i]:seg_h[s_i] + kc[s_i], 4], ref[:, 4]) < TOL


# ===================================================================== soft-NMS (a9)
@pytest.mark.parametrize("method", [0, 1, 2])
def test_soft_nms_golden(ops, method):
    g = load_golden("soft_nms")
    d = g["boxes"]
    seg = torch.tensor([0, d.shape[0]], dtype=torch.int32).cuda()
    rows, src, cnt = ops.soft_nms_batched(dev(d.copy()), seg, sigma=0.5, Nt=0.7, threshold=0.1, method=method)
    n = int(cnt.item())
    ref = g["rows_m%d" % method]
    assert n == ref.shape[0]
    np.testing.assert_array_equal(npy(rows[:n, :4]), ref[:, :4])
    assert rel_err(npy(rows[:n, 4]), ref[:, 4]) < TOL
    np.testing.assert_array_equal(d[npy(src[:n]), :4], ref[:, :4])       # src_idx maps back to input rows


def test_soft_nms_segments_vs_oracle(ops, oracle_mod):
    sizes = [300, 0, 1, 2500, 70]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    d = synth.nms_stress_boxes(int(offs[-1]), 66).numpy()
    rows, src, cnt = ops.soft_nms_batched(dev(d.copy()), dev(offs), sigma=0.5, Nt=0.7, threshold=0.1, method=2)
    rows, cnt = npy(rows), npy(cnt)
    for s, n in enumerate(sizes):
        a = offs[s]
        ref = oracle_mod.soft_nms(d[a:a + n], sigma=0.5, Nt=0.7, threshold=0.1, method=2)
        assert cnt[s] == ref.shape[0]
        np.testing.assert_array_equal(rows[a:a + cnt[s], :4], ref[:, :4])
        assert rel_err(rows[a:a + cnt[s], 4], ref[:, 4]) < TOL


def test_soft_nms_large_segment_global_path(ops, oracle_mod):
    d = synth.nms_stress_boxes(7000, 67).numpy()                          # > shared-memory capacity
    seg = torch.tensor([0, 7000], dtype=torch.int32).cuda()
    rows, src, cnt = ops.soft_nms_batched(dev(d.copy()), seg, sigma=0.5, Nt=0.7, threshold=0.1, method=2)
    ref = oracle_mod.soft_nms(d, sigma=0.5, Nt=0.7, threshold=0.1, method=2)
    n = int(cnt.item())
    assert n == ref.shape[0]
    np.testing.assert_array_equal(npy(rows[:n, :4]), ref[:, :4])
    assert rel_err(npy(rows[:n, 4]), ref[:, 4]) < TOL


# ===================================================================== target render (a11)
def test_render_demo_known_answer(ops):
    g = load_golden("render")
    a = g["demo_annos"]
    annos = torch.zeros(1, a.shape[0] + 3, 8)
    annos[0, : a.shape[0]] = torch.from_numpy(a)
    n_obj = torch.tensor([a.shape[0]], dtype=torch.int32)
    hm, wh, ind, off, msk = ops.render_targets(annos.cuda(), n_obj.cuda(), 540, 960)
    hm = npy(hm)[0]
    assert hm.shape == (10, 135, 240) and int((hm == 1).sum()) == 81
    assert abs(float(hm.astype(np.float64).sum()) - 294.5373923947336) < 1e-3
    np.testing.assert_array_equal(hm > 0, g["demo_hm"] > 0)
    np.testing.assert_array_equal(hm == 1, g["demo_hm"] == 1)
    assert rel_err(hm, g["demo_hm"]) < TOL
    n = a.shape[0]
    np.testing.assert_array_equal(npy(wh)[0, :n], g["demo_wh"])
    np.testing.assert_array_equal(npy(ind)[0, :n], g["demo_ind"])
    np.testing.assert_array_equal(npy(off)[0, :n], g["demo_off"])
    np.testing.assert_array_equal(npy(msk)[0, :n], g["demo_mask"])
    assert (npy(wh)[0, n:] == 0).all() and (npy(msk)[0, n:] == 0).all()   # collate zero padding


def test_render_batch_vs_oracle_config3(ops, oracle_mod):
    """Config 3 targets: B=32 at 512x512; order independence makes the result bit-reproducible."""
    B = 32
    annos = synth.train_annos(B, 512, 512, synth.SEED_C3)
    padded, cnt = synth.pad_annos(annos)
    hm, wh, ind, off, msk = ops.render_targets(padded.cuda(), cnt.cuda(), 512, 512)
    hm2 = ops.render_targets(padded.cuda(), cnt.cuda(), 512, 512)[0]
    assert torch.equal(hm, hm2)
    for b in (0, 7, 31):
        r = oracle_mod.render(annos[b].numpy(), 512, 512)
        np.testing.assert_array_equal(npy(hm[b]) > 0, r["hm"] > 0)
        np.testing.assert_array_equal(npy(hm[b]) == 1, r["hm"] == 1)
        assert rel_err(npy(hm[b]), r["hm"]) < TOL
        n = annos[b].shape[0]
        np.testing.assert_array_equal(npy(wh[b, :n]), r["wh"])
        np.testing.assert_array_equal(npy(ind[b, :n]), r["ind"])
        np.testing.assert_array_equal(npy(off[b, :n]), r["offset"])
        np.testing.assert_array_equal(npy(msk[b, :n]), r["reg_mask"])


# ===================================================================== focal loss (a12)
def _check_focal(ops, z, gt, loss_ref, grad_ref):
    zd, gd = dev(z), dev(gt)
    stats = npy(ops.focal_forward(zd, gd))
    assert abs(stats[0] - loss_ref) / abs(loss_ref) < TOL
    assert stats[3] == float((gt == 1).sum())
    grad = npy(ops.focal_backward(zd, gd, dev(stats.copy()), 1.0))
    assert rel_err(grad, grad_ref, floor=1.0) < TOL
    stats2, grad2 = ops.focal_fwd_bwd(zd, gd, 0.5)
    assert abs(float(stats2[0]) - loss_ref) / abs(loss_ref) < TOL
    assert rel_err(npy(grad2), 0.5 * np.asarray(grad_ref), floor=1.0) < TOL


def test_focal_golden(ops):
    g = load_golden("focal")
    _check_focal(ops, g["logits"], g["gt"], float(g["loss"]), g["grad"])
    _check_focal(ops, g["logits"], g["gt_nopos"], float(g["loss_nopos"]), g["grad_nopos"])


def test_focal_config3_vs_oracle(ops, oracle_mod):
    B = 32
    annos = synth.train_annos(B, 512, 512, synth.SEED_C3)
    padded, cnt = synth.pad_annos(annos)
    gt = ops.render_targets(padded.cuda(), cnt.cuda(), 512, 512)[0]
    z = torch.randn(B, 10, 128, 128, generator=torch.Generator().manual_seed(synth.SEED_C3)) * 2.0 - 2.5
    loss, sums, grad = oracle_mod.focal(z.numpy(), npy(gt), want_grad=True)
    _check_focal(ops, z.numpy(), npy(gt), loss, grad)
    # deterministic reduction: identical bits on a second run
    a = ops.focal_forward(z.cuda(), gt)
    b = ops.focal_forward(z.cuda(), gt)
    assert torch.equal(a, b)


def test_focal_odd_length(ops, oracle_mod):
    g = torch.Generator().manual_seed(5)
    z = torch.randn(4 * 1001 + 3, generator=g) * 3
    gt = torch.rand(4 * 1001 + 3, generator=g)
    gt[::50] = 1.0
    # 16-byte alignment is required of the base pointers only
    loss, _, grad = oracle_mod.focal(z.numpy(), gt.numpy(), want_grad=True)
    _check_focal(ops, z.numpy(), gt.numpy(), loss, grad)


def test_focal_render_fused_vs_unfused_and_oracle(ops, oracle_mod):
    """Fused target render + focal loss (config 3 shape): same loss / gradient as render -> focal, the on-the-fly
    target equals the rendered one bit for bit, and the loss matches the CPU oracle."""
    B, C, img = 32, 10, 512
    annos_l = synth.train_annos(B, img, img, synth.SEED_C3)
    annos, n_obj = synth.pad_annos(annos_l)
    g = torch.Generator().manual_seed(synth.SEED_C3)
    z = torch.randn(B, C, img // 4, img // 4, generator=g) * 2.5 - 2.0
    zd, ad, nd = dev(z), dev(annos), dev(n_obj)
    gt = ops.render_targets(ad, nd, img, img)[0]
    stats_u, grad_u = ops.focal_fwd_bwd(zd, gt, 0.5)
    stats_f, gt_f = ops.focal_render_forward(zd, ad, nd, img, img, want_gt=True)
    assert torch.equal(gt_f, gt)
    grad_f = ops.focal_render_backward(zd, ad, nd, img, img, stats_f, 0.5)
    assert float(stats_f[3]) == float(stats_u[3])                                     # num_pos
    assert abs(float(stats_f[0]) - float(stats_u[0])) <= 1e-6 * abs(float(stats_u[0]))
    assert rel_err(npy(grad_f), npy(grad_u), floor=1e-3) < 1e-6
    gts = np.stack([oracle_mod.render(a.numpy(), img, img)["hm"] for a in annos_l])
    loss, _, gref = oracle_mod.focal(z.numpy(), gts, want_grad=True)
    assert abs(float(stats_f[0]) - loss) / abs(loss) < TOL
    assert rel_err(npy(grad_f) / 0.5, gref, floor=1e-3) < TOL
    # loss and gradient in ONE pass (num_pos counted from the annotations first): same numbers
    stats_1, grad_1 = ops.focal_render_fwd_bwd(zd, ad, nd, img, img, 0.5)
    assert float(stats_1[3]) == float(stats_u[3])
    assert abs(float(stats_1[0]) - float(stats_u[0])) <= 1e-6 * abs(float(stats_u[0]))
    assert torch.equal(grad_1, grad_f)


@pytest.mark.parametrize("img_h,img_w", [(512, 512), (96, 176), (40, 48)])
def test_focal_render_fwd_bwd_center_count(ops, img_h, img_w):
    """The single-pass form needs num_pos before it has seen the map: objects sharing a centre cell and class count once,
    objects of different classes on the same cell count twice, undrawn rows (class 0 = 'ignored', padding) not at all;
    tile rows that do not start on a row boundary (176/4 = 44 pixels per row, 4096 % 44 != 0)."""
    B, C = 3, 10
    g = torch.Generator().manual_seed(img_h + img_w)
    lists = synth.train_annos(B, img_h, img_w, 91, n_range=(8, 24))
    a0 = lists[0]
    dup = a0[:3].clone()                           # same boxes again: same class, same centre
    other = a0[:2].clone(); other[:, 5] = (other[:, 5] % 10) + 1     # same centre, another class
    ignored = a0[:2].clone(); ignored[:, 5] = 0    # class 0: python index -1 -> the last plane (functional.py:249)
    lists[0] = torch.cat([a0, dup, other, ignored])
    annos, n_obj = synth.pad_annos(lists)
    z = torch.randn(B, C, img_h // 4, img_w // 4, generator=g) * 2.5 - 2.0
    zd, ad, nd = dev(z), dev(annos), dev(n_obj)
    gt = ops.render_targets(ad, nd, img_h, img_w)[0]
    stats_u, grad_u = ops.focal_fwd_bwd(zd, gt, 1.0)
    assert float(stats_u[3]) == float((gt == 1).sum())
    stats_1, grad_1 = ops.focal_render_fwd_bwd(zd, ad, nd, img_h, img_w, 1.0)
    assert float(stats_1[3]) == float(stats_u[3])
    assert abs(float(stats_1[0]) - float(stats_u[0])) <= 1e-6 * abs(float(stats_u[0]))
    assert rel_err(npy(grad_1), npy(grad_u), floor=1e-3) < 1e-6
    _, gt_f = ops.focal_render_forward(zd, ad, nd, img_h, img_w, want_gt=True)
    assert torch.equal(gt_f, gt)


def test_focal_render_no_objects_and_odd_plane(ops):
    z = torch.randn(2, 3, 10, 12, generator=torch.Generator().manual_seed(5)).cuda()   # 120 pixels per plane, < one CTA
    annos = torch.zeros(2, 1, 8).cuda()
    n_obj = torch.zeros(2, dtype=torch.int32).cuda()
    stats = ops.focal_render_forward(z, annos, n_obj, 40, 48)
    ref = ops.focal_forward(z, torch.zeros_like(z))
    assert float(stats[3]) == 0.0 and abs(float(stats[0]) - float(ref[0])) <= 1e-6 * abs(float(ref[0]))
    stats_1, grad_1 = ops.focal_render_fwd_bwd(z, annos, n_obj, 40, 48)
    ref_s, ref_g = ops.focal_fwd_bwd(z, torch.zeros_like(z))
    assert float(stats_1[3]) == 0.0 and abs(float(stats_1[0]) - float(ref_s[0])) <= 1e-6 * abs(float(ref_s[0]))
    assert rel_err(npy(grad_1), npy(ref_g), floor=1e-3) < 1e-6
    with pytest.raises(Exception):
        ops.focal_render_forward(torch.zeros(1, 1, 4, 5).cuda(), annos[:1], n_obj[:1], 16, 20)   # row length 5: not a multiple of 4


# ===================================================================== edge cases across the path
def test_eval_path_tiny_and_degenerate_inputs(ops, oracle_mod):
    """B=1, K=1 on a 4x4 map; a map of identical logits (massive ties -> exact fallback); wh all negative
    (every box degenerate: zero area, NaN IoU never suppresses, RoIs of size 0 still sample one point)."""
    hp = synth.head_params(3)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    g = torch.Generator().manual_seed(8)
    # (1) tiny
    hm = torch.randn(1, 3, 4, 4, generator=g)
    wh = torch.rand(1, 2, 4, 4, generator=g) * 3
    off = torch.rand(1, 2, 4, 4, generator=g)
    feat = torch.randn(1, 256, 4, 4, generator=g)
    path = ops.EvalPath(1, 3, 4, 4, 1, folded, keep_roi_feat=True)
    path.forward(dev(hm), dev(wh), dev(off), dev(feat))
    r = path.results()
    dets, _, _ = oracle_mod.decode(hm.numpy(), wh.numpy(), off.numpy(), 1)
    assert r["n"] == 1 and np.array_equal(npy(r["bxyxy"])[0, 1:], dets[0, 0, :4])
    roi = oracle_mod.roi_align(feat.numpy(), npy(r["bxyxy"]), relu=True)
    assert rel_err(npy(path.roi_feat[:1]), roi, floor=1e-3) < TOL
    # (2) degenerate boxes: wh < 0 everywhere -> clamp(min=0) -> x1 == x2, y1 == y2
    B, C, H, W, K = 2, 4, 24, 40, 64
    x = synth.eval_inputs(B, H, W, K, 77, C=C)
    x["wh"] = -x["wh"].abs()
    path = ops.EvalPath(B, C, H, W, K, folded, keep_roi_feat=True)
    path.forward(*[dev(x[k]) for k in ("hm", "wh", "off", "feat")])
    r = path.results()
    dets, _, _ = oracle_mod.decode(x["hm"].numpy(), x["wh"].numpy(), x["off"].numpy(), K)
    rows = []
    for b in range(B):
        kept, _ = oracle_mod.stage1_nms(dets[b], C, 0.7)
        assert kept.shape[0] == K                       # zero-area boxes never suppress each other
        rows.append(np.concatenate([np.full((K, 1), b, np.float32), kept[:, :4]], 1))
    bxyxy = np.concatenate(rows)
    np.testing.assert_array_equal(npy(r["bxyxy"]), bxyxy)
    roi = oracle_mod.roi_align(x["feat"].numpy(), bxyxy, relu=True)
    assert rel_err(npy(path.roi_feat[: r["n"]]), roi, floor=1e-3) < TOL
    reg = oracle_mod.head(roi, {k: v.numpy() for k, v in hp.items()})
    assert rel_err(npy(r["reg"]), reg, floor=1.0) < TOL
    # (3) constant heat-map: every score ties; canonical order = flat index ascending, still K rows per image
    hm_c = torch.full((1, 2, 8, 8), -1.25)
    d, inds = ops.decode_topk(dev(hm_c), dev(torch.ones(1, 2, 8, 8)), dev(torch.zeros(1, 2, 8, 8)), 20)
    assert npy(inds)[0].tolist() == list(range(20)) and float(d[0, :, 5].max()) == 0.0


def test_zero_rois_are_a_no_op(ops):
    feat = torch.randn(1, 32, 8, 8).cuda()
    assert tuple(ops.roi_align(feat, torch.zeros(0, 5).cuda()).shape) == (0, 32, 3, 3)
    hp = synth.head_params(1)
    folded = ops.head_fold({k: v.cuda() for k, v in hp.items()})
    assert tuple(ops.head_forward(torch.zeros(0, 256, 3, 3).cuda(), folded).shape) == (0, 4)
    n_dev = torch.zeros(1, dtype=torch.int32).cuda()               # capacity 5, live count 0: outputs untouched
    out = ops.roi_align(feat, torch.rand(5, 5).cuda(), n_dev=n_dev)
    assert tuple(out.shape) == (5, 32, 3, 3)
    s1, s2 = ops.generate_bbox(torch.zeros(0, 5).cuda(), torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda())
    assert s1.shape[0] == 0 and s2.shape[0] == 0


def test_sm_reserve_and_two_batches_in_flight(ops):
    """rr_set_sm_reserve only changes the grids of the persistent kernels: same bits.  Two EvalPaths on two streams
    (the bench's batches-in-flight mode) give what one gives."""
    B, C, H, W, K = 2, 10, 152, 272, 300
    x = synth.eval_inputs(B, H, W, K, 77)
    xd = {k: dev(v) for k, v in x.items()}
    folded = ops.head_fold({k: v.cuda() for k, v in synth.head_params(77).items()})
    base = ops.EvalPath(B, C, H, W, K, folded)
    base.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
    r0 = base.results()
    L = ops._lib.lib()
    assert L.rr_set_sm_reserve(-1) != 0 and L.rr_set_sm_reserve(148) != 0
    try:
        ops.set_sm_reserve(100)
        streams = [torch.cuda.Stream() for _ in range(2)]
        paths = []
        torch.cuda.synchronize()
        for st in streams:
            with torch.cuda.stream(st):
                p = ops.EvalPath(B, C, H, W, K, folded)
                for _ in range(3):
                    p.forward(xd["hm"], xd["wh"], xd["off"], xd["feat"])
            paths.append(p)
        torch.cuda.synchronize()
    finally:
        ops.set_sm_reserve(0)
    for p in paths:
        r = p.results()
        assert r["n"] == r0["n"]
        np.testing.assert_array_equal(npy(r["s2"]), npy(r0["s2"]))
        np.testing.assert_array_equal(npy(r["reg"]), npy(r0["reg"]))


@pytest.mark.parametrize("shape", [(4, 2, 32, 48, 37), (2, 2, 128, 128, 150), (1, 3, 8, 8, 0)])
def test_regl1_loss_and_gradient(ops, shape):
    """rr_regl1_fwd_bwd against the reference's formula (modules/loss/regl1loss.py:9-17) in torch on the CPU: the
    permute + gather + l1_loss(sum) / (mask.sum() + 1e-4) and its autograd gradient; duplicate centre cells,
    masked-out rows and an empty annotation list included."""
    B, c, H, W, max_n = shape
    g = torch.Generator().manual_seed(sum(shape))
    out = torch.randn(B, c, H, W, generator=g)
    ind = torch.randint(0, H * W, (B, max_n, 1), generator=g).float()
    if max_n > 4:
        ind[:, 1] = ind[:, 0]                                    # two objects in one cell
    mask = (torch.rand(B, max_n, 1, generator=g) > 0.3)
    target = torch.rand(B, max_n, c, generator=g) * 30
    ref_out = out.clone().requires_grad_(True)
    pred = ref_out.permute(0, 2, 3, 1).contiguous().view(B, -1, c)
    pred = pred.gather(1, ind.long().expand(B, max_n, c))
    m = mask.expand_as(pred).float()
    ref = torch.nn.functional.l1_loss(pred * m, target * m, reduction="sum") / (m.sum() + 1e-4)
    ref.backward()
    loss, grad = ops.regl1_fwd_bwd(dev(out), dev(mask.float()), dev(ind), dev(target), grad_scale=1.0)
    assert abs(float(loss) - float(ref)) <= TOL * max(abs(float(ref)), 1e-6)
    assert rel_err(npy(grad), ref_out.grad.numpy(), floor=1e-6) < TOL
    loss2, none = ops.regl1_fwd_bwd(dev(out), dev(mask.float()), dev(ind), dev(target), want_grad=False)
    assert none is None and float(loss2) == float(loss)          # fixed-order sums: bit-reproducible


def test_stage2_loss_and_gradients(ops):
    """rr_stage2_loss against the reference's per-image loop (operators/rrnet_operator.py:64-102) in torch on the CPU,
    with autograd for d/d s2_reg and d/d bxyxy (the targets are not detached).  Image 2 has no positive (factor 0),
    image 3 no RoI-GT overlap at all; padded ground-truth rows are zero boxes."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(9)
    B, max_n, scale = 4, 12, 4.0
    gt = torch.zeros(B, max_n, 8)
    rows, regs = [], []
    for b in range(B):
        n_gt = [7, 12, 5, 3][b]
        xy = torch.rand(n_gt, 2, generator=g) * 300 + 20
        wh = torch.rand(n_gt, 2, generator=g) * 60 + 12
        gt[b, :n_gt, :2] = xy
        gt[b, :n_gt, 2:4] = xy + wh                                       # already xyxy (after :67)
        n = [40, 33, 25, 18][b]
        pick = torch.randint(0, n_gt, (n,), generator=g)
        jit = (torch.rand(n, 4, generator=g) - 0.5) * (8 if b < 2 else 400)      # images 2, 3: far off -> no positives
        box = (torch.cat([xy[pick], xy[pick] + wh[pick]], 1) + jit) / scale
        if b == 3:
            box = box + 2000.0
        box[:, 2:] = torch.maximum(box[:, 2:], box[:, :2])
        rows.append(torch.cat([torch.full((n, 1), float(b)), box], 1))
        regs.append(torch.randn(n, 4, generator=g) * 0.7)
    bxyxy = torch.cat(rows)
    s2_reg = torch.cat(regs)
    # reference loop
    rb = bxyxy.clone().requires_grad_(True)
    rr = s2_reg.clone().requires_grad_(True)
    loss = 0
    n_pos_seen = []
    for b in range(B):
        flag = rb[:, 0] == b
        bbox = rb[flag][:, 1:]
        a, bb = bbox * scale, gt[b, :, :4]
        area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
        area_b = (bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])
        whi = (torch.min(a[:, None, 2:], bb[None, :, 2:]) - torch.max(a[:, None, :2], bb[None, :, :2])).clamp(min=0)
        inter = whi[..., 0] * whi[..., 1]
        iou = inter / (area_a[:, None] + area_b[None, :] - inter)
        max_iou, max_idx = torch.max(iou, dim=1)
        pos = max_iou > 0.5
        n_pos_seen.append(int(pos.sum()))
        if pos.sum() == 0:
            pos = torch.zeros_like(pos)
            pos[0] = True
            factor = 0
        else:
            factor = 1
        ex, gr = bbox[pos] * scale, bb[max_idx[pos]]
        ew, eh = ex[:, 2] - ex[:, 0] + 1.0, ex[:, 3] - ex[:, 1] + 1.0
        ecx, ecy = ex[:, 0] + 0.5 * ew, ex[:, 1] + 0.5 * eh
        gw, gh = gr[:, 2] - gr[:, 0] + 1.0, gr[:, 3] - gr[:, 1] + 1.0
        gcx, gcy = gr[:, 0] + 0.5 * gw, gr[:, 1] + 0.5 * gh
        tgt = torch.stack(((gcx - ecx) / ew, (gcy - ecy) / eh, torch.log(gw / ew), torch.log(gh / eh)), 1)
        loss = loss + F.smooth_l1_loss(rr[flag][pos], tgt) * factor / B
    loss.backward()
    assert n_pos_seen[0] > 0 and n_pos_seen[1] > 0 and n_pos_seen[2] == 0 and n_pos_seen[3] == 0
    seg = torch.tensor([0, 40, 73, 98, 116], dtype=torch.int32)
    parts, g_reg, g_box = ops.stage2_loss(dev(bxyxy), dev(seg), dev(s2_reg), dev(gt), scale)
    assert abs(float(parts.sum()) - float(loss)) <= TOL * abs(float(loss))
    assert float(parts[2]) == 0.0 and float(parts[3]) == 0.0
    assert rel_err(npy(g_reg), rr.grad.numpy(), floor=1e-4) < TOL
    assert rel_err(npy(g_box), rb.grad.numpy()[:, 1:], floor=1e-4) < TOL


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("C", [64, 40])
def test_roi_align_backward_vs_torchvision(ops, algo, C):
    """rr_roi_align_backward against autograd through torchvision.ops.roi_align(relu(feat)) on the CPU: small and
    large RoIs (a window over 64 pixels goes through the direct path), RoIs over the map border, empty and
    degenerate boxes, overlapping RoIs in one tile, two images; C = 40 forces the direct path (C % 32 != 0).
    Tolerance 3e-5 relative: a gradient element is a sum over all samples of all RoIs that touch the pixel; the
    reference accumulates it sample by sample (in an arbitrary atomic order on the GPU), this kernel through the
    separable per-axis weights, so the fp32 sums are associated differently (measured: 1.0e-5)."""
    import torchvision
    tol = 3e-5
    g = torch.Generator().manual_seed(31 + C)
    B, H, W = 2, 70, 100
    feat = torch.randn(B, C, H, W, generator=g)
    n = 60
    cx, cy = torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H
    bw, bh = torch.rand(n, generator=g) * 30 + 1, torch.rand(n, generator=g) * 30 + 1
    rois = torch.stack([torch.randint(0, B, (n,), generator=g).float(), cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    rois[0, 1:] = torch.tensor([3.0, 2.0, 95.0, 66.0])          # window > 64: direct path
    rois[1, 1:] = torch.tensor([-20.0, -10.0, 12.0, 9.0])        # over the top-left border
    rois[2, 1:] = torch.tensor([90.0, 60.0, 140.0, 99.0])        # over the bottom-right border
    rois[3, 1:] = torch.tensor([50.0, 30.0, 50.0, 30.0])         # degenerate (size clamps to 1)
    rois[4, 1:] = torch.tensor([300.0, 300.0, 320.0, 330.0])     # entirely outside
    rois = rois[torch.argsort(rois[:, 0], stable=True)]
    gout = torch.randn(n, C, 3, 3, generator=g)
    f = feat.clone().requires_grad_(True)
    out = torchvision.ops.roi_align(torch.relu(f), rois, (3, 3))
    (ref,) = torch.autograd.grad(out, f, gout)
    got = npy(ops.roi_align_backward(dev(feat), dev(rois), dev(gout), relu=True, algo=algo))
    assert rel_err(got, ref.numpy(), floor=1e-3) < tol
    # without the ReLU
    f2 = feat.clone().requires_grad_(True)
    (ref2,) = torch.autograd.grad(torchvision.ops.roi_align(f2, rois, (3, 3)), f2, gout)
    got2 = npy(ops.roi_align_backward(dev(feat), dev(rois), dev(gout), relu=False, algo=algo))
    assert rel_err(got2, ref2.numpy(), floor=1e-3) < tol
    if algo == 0 and C % 32 == 0:              # tile path: pieces summed in RoI order, one atomicAdd per element for RoI 0
        for _ in range(3):
            again = npy(ops.roi_align_backward(dev(feat), dev(rois), dev(gout), relu=True, algo=0))
            np.testing.assert_array_equal(again, got)


def test_ap_match_vs_reference_goldens_and_oracle(ops):
    """rr_ap_match: all seeded images as ONE padded batch against the reference's own get_tp outputs
    (tests/golden/ap_match.npz) and against oracle/metrics_np on the same inputs - bit-exact flags, counts, order."""
    from oracle import metrics_np
    g = load_golden("ap_match")
    cases = [synth.ap_match_case(*c) for c in synth.AP_MATCH_CASES]
    B = len(cases)
    M = max(p.shape[0] for p, _ in cases) + 3
    N = max(t.shape[0] for _, t in cases) + 2
    pred = torch.zeros(B, M, 6)
    tgt = torch.zeros(B, N, 6)
    for b, (p, t) in enumerate(cases):
        pred[b, : p.shape[0]] = p
        tgt[b, : t.shape[0]] = t
    n_pred = torch.tensor([p.shape[0] for p, _ in cases], dtype=torch.int32)
    n_tgt = torch.tensor([t.shape[0] for _, t in cases], dtype=torch.int32)
    thr = torch.from_numpy(g["thresholds"])
    order, tp, cls, cnt, img = [npy(x) for x in ops.ap_match(dev(pred), dev(n_pred), dev(tgt), dev(n_tgt), dev(thr))]
    for b, (p, t) in enumerate(cases):
        m = p.shape[0]
        o = metrics_np.get_tp_image(p.numpy(), t.numpy(), g["thresholds"])
        np.testing.assert_array_equal(cnt[b], g["target_count_%d" % b])
        np.testing.assert_array_equal(img[b], g["in_img_%d" % b])
        assert (order[b, m:] == -1).all() and (cls[b, m:] == -1).all()
        assert sorted(order[b, :m].tolist()) == list(range(m))
        sc = p[:, 4].numpy()[order[b, :m]]
        assert (sc[:-1] >= sc[1:]).all()
        # the emitted detections, class by class in rank order = the reference's concatenated lists
        tps, confs, sizes = [], [], []
        for c in range(1, 11):
            sel = cls[b, :m] == c
            tps.append(tp[b, :m][sel]); confs.append(sc[sel]); sizes.append(int(sel.sum()))
        np.testing.assert_array_equal(np.array(sizes), g["sizes_%d" % b])
        np.testing.assert_array_equal(np.concatenate(tps), g["tp_%d" % b])
        np.testing.assert_array_equal(np.concatenate(confs), g["conf_%d" % b])
        # and the oracle's view of the same image (kept detections, their classes)
        kept = np.zeros(m, bool)
        kept[o["order"][o["emit"]]] = True
        np.testing.assert_array_equal(cls[b, :m] >= 0, kept[order[b, :m]])
