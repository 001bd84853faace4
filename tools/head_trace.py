"""Per-CTA phase timeline of head_tc_kernel on the bench workload (development tool, not shipped).

Builds a second copy of the library with -DRR_HEAD_TC_TRACE (time stamps in a __device__ array), runs the
eval path once warm, and prints where a CTA spends its time.  Usage on the GPU box: python tools/head_trace.py
"""
import ctypes, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnet_b200 import build as B  # noqa: E402

TRACE_LIB = os.path.join(ROOT, "tools", "librrnet_trace.so")


VARIANTS = {"": [], "oneslot": ["-DRR_HEAD_PROBE_ONE_SLOT"], "noscr": ["-DRR_HEAD_PROBE_NO_SCRATCH"],
            "oneslot_noscr": ["-DRR_HEAD_PROBE_ONE_SLOT", "-DRR_HEAD_PROBE_NO_SCRATCH"]}      # timing-only: wrong results by design


def lib_path(variant):
    return TRACE_LIB.replace(".so", ("_" + variant if variant else "") + ".so")


def build_trace_lib():
    B.build()
    for variant, defs in VARIANTS.items():          # timing-only experiments: results are wrong by design
        obj = os.path.join(ROOT, "tools", "rr_head_tc_trace_%s.o" % variant)
        subprocess.check_call([B._nvcc()] + B.BASE + ["-DRR_HEAD_TC_TRACE"] + defs +
                              ["-c", os.path.join(B.CSRC, "rr_head_tc.cu"), "-o", obj])
        objs = [os.path.join(B.OBJ, u.replace(".cu", ".o")) for u in B.UNITS if u != "rr_head_tc.cu"] + [obj]
        subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_path(variant)]
                              + objs + ["-lcudart"])


def main():
    if "--build" in sys.argv or not os.path.exists(TRACE_LIB):
        build_trace_lib()
        if "--build" in sys.argv:
            return
    variant = ""
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            variant = a.split("=", 1)[1]
    if variant == "all":
        for v in VARIANTS:
            print("================ variant '%s'" % v, flush=True)
            subprocess.call([sys.executable, os.path.abspath(__file__), "--variant=" + v])
        return
    import torch
    from rrnet_b200 import _lib
    _lib.LIB_PATH = lib_path(variant)
    from rrnet_b200 import ops, synth
    import bench
    w = bench.WORKLOAD
    Bn, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
    dev = torch.device("cuda", 0)
    x = synth.eval_inputs(Bn, H, W, K, synth.SEED_C2)
    d = {k: v.to(dev) for k, v in x.items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
    path = ops.EvalPath(Bn, C, H, W, K, folded, device=dev)
    for _ in range(3):
        path.forward(d["hm"], d["wh"], d["off"], d["feat"])
    torch.cuda.synchronize()
    L = _lib.lib()
    n = 2048 * 32
    buf = (ctypes.c_uint64 * n)()
    L.rr_debug_head_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = L.rr_debug_head_trace(buf, n)
    assert rc == 0, rc
    t = np.frombuffer(buf, dtype=np.uint64).reshape(2048, 32).astype(np.int64)
    n_live = int(path.counts[-1].item())
    n_cta = min(148, (n_live + 7) // 8)
    t = t[:n_cta]
    t0 = t[:, 0].min()
    n_my = t[:, 13]
    print("CTAs %d; kernel span %.1f us; tiles per CTA %d..%d" % (n_cta, (t[:, 12].max() - t0) / 1e3, n_my.min(), n_my.max()))
    print("setup %.2f us; first x tile staged after %.2f us; steady-state period %.2f us per tile (8 RoIs)" % (
        (t[:, 1] - t[:, 0]).mean() / 1e3, (t[:, 2] - t[:, 1]).mean() / 1e3, ((t[:, 12] - t[:, 2]) / n_my).mean() / 1e3))
    names = ["wait conv1", "E1", "-", "wait conv2", "E2", "E3 (incl. wait conv3)"]
    print("epilogue thread 0, mean us per tile: " + "  ".join("%s %.2f" % (nm, (t[:, 16 + i] / n_my).mean() / 1e3) for i, nm in enumerate(names)))
    print("issuer, mean us per tile blocked: on ring 2 %.2f  on t1/t2 %.2f  in blocking conv1 steps %.2f" % tuple(
        (t[:, i] / n_my).mean() / 1e3 for i in (22, 23, 24)))
    print("conv1 issuer, mean us per tile blocked: on x stages %.2f  on its weight slot %.2f | x loaders: %.2f us per tile, of which blocked on a free stage %.2f" % (
        (t[:, 25] / n_my).mean() / 1e3, (t[:, 26] / n_my).mean() / 1e3, (t[:, 28] / n_my).mean() / 1e3, (t[:, 27] / n_my).mean() / 1e3))
    # how many tile pieces the RoIs of this workload are cut into (approximation of roi_prep_kernel's count)
    bx = path.bxyxy[:n_live].float().cpu().numpy()
    ntx = np.floor((bx[:, 3] + 1) / 32) - np.floor(bx[:, 1] / 32) + 1
    nty = np.floor((bx[:, 4] + 1) / 24) - np.floor(bx[:, 2] / 24) + 1
    pcs = (ntx * nty).astype(int)
    print("RoIs %d; pieces histogram:" % n_live, np.bincount(pcs)[:16], " mean %.2f" % pcs.mean())
    grp = pcs[: (len(pcs) // 8) * 8].reshape(-1, 8)
    print("CTAs whose 8 RoIs include one with > 6 pieces: %.1f %%" % (100.0 * (grp.max(axis=1) > 6).mean()))


if __name__ == "__main__":
    main()
