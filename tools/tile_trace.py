"""Per-CTA phase times of roi_tile_kernel on the bench workload (development tool): staging (global loads -> smem),
wait at the barrier after staging, unit evaluation, wait at the barrier after the units; for warp 0 and the last warp."""
import ctypes, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnet_b200 import build as B  # noqa: E402

LIB = os.path.join(ROOT, "tools", "librrnet_tiletrace.so")


def build_lib():
    B.build()
    obj = os.path.join(ROOT, "tools", "rr_roialign_trace.o")
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    subprocess.check_call([B._nvcc()] + B.BASE + B.UNITS["rr_roialign.cu"] + ["-DRR_TILE_TRACE"] + extra + ["-c",
                          os.path.join(B.CSRC, "rr_roialign.cu"), "-o", obj])
    objs = [os.path.join(B.OBJ, u.replace(".cu", ".o")) for u in B.UNITS if u != "rr_roialign.cu"] + [obj]
    subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"])


def main():
    if "--build" in sys.argv or not os.path.exists(LIB):
        build_lib()
        if "--build" in sys.argv:
            return
    import torch
    from rrnet_b200 import _lib
    _lib.LIB_PATH = LIB
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
    L = _lib.lib()
    n = 512 * 16
    buf = (ctypes.c_uint64 * n)()
    L.rr_debug_tile_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert L.rr_debug_tile_trace(buf, n) == 0
    t = np.frombuffer(buf, dtype=np.uint64).reshape(512, 2, 8).astype(np.float64)[:296]
    for wi, name in ((0, "warp 0"), (1, "last warp")):
        x = t[:, wi]
        tick = x[:, 4]
        print("%-9s per ticket (us): stage %.2f  barrier %.2f  units %.2f  barrier %.2f | tickets/CTA %.1f  units/ticket %.2f | CTA %.1f us" % (
            name, *(float((x[:, i] / tick).mean()) / 1e3 for i in range(4)), tick.mean(), (x[:, 5] / tick).mean(), x[:, 6].mean() / 1e3))


if __name__ == "__main__":
    main()
