"""A/B of roi_tile_tma_kernel build variants on the bench workload (development tool).

    python tools/tile_variants.py --build name1:-DX=1,-DY=2 name2:...     (here, no GPU: compiles tools/librrnet_var_<name>.so)
    python tools/tile_variants.py name1 name2 ...                        (on the GPU box: one subprocess per variant)

Per variant: median device time of roi_prep / roi_fill / roi_tile_tma / head over 7 traced steps (ops.KernelTrace),
single-batch graph-replay ms per step, and whether `reg` equals the shipped library's bit for bit ("base" = the
library in rrnet_b200/)."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lib_of(name):
    return os.path.join(ROOT, "rrnet_b200", "librrnet_b200.so") if name == "base" else os.path.join(ROOT, "tools", "librrnet_var_%s.so" % name)


def build(specs):
    from rrnet_b200 import build as B
    B.build()
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        units = {}
        for f in [x for x in flags.split(",") if x]:
            unit = "rr_head_tc.cu" if "HEAD" in f else "rr_roialign.cu"
            units.setdefault(unit, []).append(f)
        objs = []
        for u in B.UNITS:
            if u in units:
                obj = os.path.join(ROOT, "tools", "var_%s_%s.o" % (name, u.replace(".cu", "")))
                procs.append(subprocess.Popen([B._nvcc()] + B.BASE + B.UNITS[u] + units[u] + ["-c", os.path.join(B.CSRC, u), "-o", obj]))
                objs.append(obj)
            else:
                objs.append(os.path.join(B.OBJ, u.replace(".cu", ".o")))
        procs.append((name, objs))
    pending = [p for p in procs if not isinstance(p, tuple)]
    for p in pending:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    for name, objs in [p for p in procs if isinstance(p, tuple)]:
        subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_of(name)] + objs + ["-lcudart"])
        print("built", lib_of(name))


def run_one(name):
    import numpy as np
    import torch
    from rrnet_b200 import _lib
    _lib.LIB_PATH = lib_of(name)
    from rrnet_b200 import ops, synth
    import bench
    w = bench.WORKLOAD
    Bn, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
    dev = torch.device("cuda", 0)
    d = {k: v.to(dev) for k, v in synth.eval_inputs(Bn, H, W, K, synth.SEED_C2).items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
    path = ops.EvalPath(Bn, C, H, W, K, folded, device=dev)
    for _ in range(3):
        path.forward(d["hm"], d["wh"], d["off"], d["feat"])
    torch.cuda.synchronize()
    acc = {}
    for _ in range(7):
        with ops.KernelTrace(capacity=64) as kt:
            path.forward(d["hm"], d["wh"], d["off"], d["feat"])
        for k, ms in kt.kernels:
            acc.setdefault(k, []).append(ms)
    med = {k: float(np.median(v)) * 1e3 for k, v in acc.items()}
    g = path.capture(d["hm"], d["wh"], d["off"], d["feat"])
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    reg = path.results()["reg"].clone()
    ref_file = "/tmp/tile_variants_ref.pt"
    same = "-"
    if name == "base":
        torch.save(reg.cpu(), ref_file)
    elif os.path.exists(ref_file):
        ref = torch.load(ref_file)
        same = "bit-identical" if torch.equal(ref, reg.cpu()) else "max diff %.3g" % float((ref - reg.cpu()).abs().max())
    print("%-14s prep %5.1f  fill %5.1f  tile %6.1f  head %6.1f us | step %.4f ms | reg vs base: %s" % (
        name, med.get("roi_prep_kernel", 0), med.get("roi_fill_kernel", 0), med.get("roi_tile_tma_kernel", 0),
        med.get("head_tc_kernel", 0), ms, same), flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--build":
        build(args[1:])
    elif args and args[0] == "--one":
        run_one(args[1])
    else:
        for name in args:
            subprocess.call([sys.executable, os.path.abspath(__file__), "--one", name])
