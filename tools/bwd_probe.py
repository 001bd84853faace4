"""Device time of rr_roi_align_backward at the bench workload (config 2 features, the step's own RoIs), per kernel
(ops.KernelTrace), for the shipped library or a tools/librrnet_var_<name>.so build (tools/tile_variants.py --build)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(name):
    import numpy as np, torch
    from rrnet_b200 import _lib
    if name != "base":
        _lib.LIB_PATH = os.path.join(ROOT, "tools", "librrnet_var_%s.so" % name)
    from rrnet_b200 import ops, synth
    import bench
    w = bench.WORKLOAD
    B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
    dev = torch.device("cuda", 0)
    x = {k: v.to(dev) for k, v in synth.eval_inputs(B, H, W, K, synth.SEED_C2).items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
    p = ops.EvalPath(B, C, H, W, K, folded, device=dev)
    p.forward(x["hm"], x["wh"], x["off"], x["feat"])
    n = p.results()["n"]
    rois = p.bxyxy[:n].clone()
    g = torch.Generator().manual_seed(1)
    gout = torch.randn(n, 256, 3, 3, generator=g).to(dev)
    rws = torch.empty(ops._lib.lib().rr_roi_align_workspace_bytes(n, *x["feat"].shape), dtype=torch.uint8, device=dev)
    for _ in range(2):
        out = ops.roi_align_backward(x["feat"], rois, gout, ws=rws)
    torch.cuda.synchronize()
    acc = {}
    for _ in range(5):
        with ops.KernelTrace(capacity=32) as kt:
            out = ops.roi_align_backward(x["feat"], rois, gout, ws=rws)
        for k, ms in kt.kernels:
            acc.setdefault(k, []).append(ms)
    print("%-10s" % name, "  ".join("%s %.1f us" % (k.replace("_kernel", ""), float(np.median(v)) * 1e3) for k, v in acc.items()),
          "| checksum %.6e" % float(out.double().abs().sum()), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        run_one(sys.argv[2])
    else:
        for nm in sys.argv[1:]:
            subprocess.call([sys.executable, os.path.abspath(__file__), "--one", nm])
