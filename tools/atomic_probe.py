"""RR_OPT_COMBINE_IN_TILE_KERNEL = 0 (partial slots to the head, default) / 1 (deterministic in-kernel combine) / 2 (float
reductions into the RoI's row): per-kernel device times, graph-replay step, two batches in flight, and `reg` against mode 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from rrnet_b200 import ops, synth

w = bench.WORKLOAD
B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
dev = torch.device("cuda", 0)
x = {k: v.to(dev) for k, v in synth.eval_inputs(B, H, W, K, synth.SEED_C2).items()}
folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
ref = None
for mode in (0, 1, 2):
    ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, mode)
    p = ops.EvalPath(B, C, H, W, K, folded, device=dev)
    p.ws.fill_(77)                                      # garbage workspace: nothing may depend on its contents
    for _ in range(3):
        p.forward(x["hm"], x["wh"], x["off"], x["feat"])
    torch.cuda.synchronize()
    acc = {}
    for _ in range(7):
        with ops.KernelTrace(capacity=64) as kt:
            p.forward(x["hm"], x["wh"], x["off"], x["feat"])
        for k, ms in kt.kernels:
            acc.setdefault(k, []).append(ms)
    med = {k: float(np.median(v)) * 1e3 for k, v in acc.items()}
    g = p.capture(x["hm"], x["wh"], x["off"], x["feat"])
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    ms1 = a.elapsed_time(b) / 50
    reg = p.results()["reg"].clone()
    n = p.results()["n"]
    runs = []
    for _ in range(3):                                   # run-to-run reproducibility
        g.replay(); torch.cuda.synchronize()
        runs.append(p.results()["reg"].clone())
    repro = all(torch.equal(runs[0], r) for r in runs[1:])
    if ref is None:
        ref = reg
        diff = "-"
    else:
        d = (reg[:n] - ref[:n]).abs()
        diff = "max abs %.3g, rel to max|reg| %.3g, rows differing %d of %d" % (float(d.max()), float(d.max() / ref[:n].abs().max()), int((d > 0).any(dim=1).sum()), n)
    print("mode %d: tile %6.1f  head %6.1f us (memset etc. in step) | step %.4f ms | run-to-run identical: %s | vs mode 0: %s" % (
        mode, med.get("roi_tile_tma_kernel", 0), med.get("head_tc_kernel", 0), ms1, repro, diff), flush=True)
ops.set_option(ops.OPT_COMBINE_IN_TILE_KERNEL, 0)
