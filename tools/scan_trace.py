"""Where the greedy-scan resolver warp spends its cycles (debug build with -DRR_SCAN_TRACE, tools/librrnet_scantrace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rrnet_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librrnet_scantrace.so")
import torch
from rrnet_b200 import ops, synth
L = _lib.lib()
for n in (256, 1024, 1500, 20000):
    d = synth.nms_stress_boxes(n, synth.SEED_C5).cuda()
    seg = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    b, s = d[:, :4].contiguous(), d[:, 4].contiguous()
    for _ in range(3):
        ops.nms_batched(b, s, seg, 0.7)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    ctypes.CDLL(_lib.LIB_PATH).rr_debug_scan_trace(buf)
    t = list(buf)
    nt = max(t[5], 1)
    names = ["wait_ring", "wait_col", "resolve", "redux", "outputs", "-", "near_loads", "publish"]
    print(n, "tiles", nt, {names[i]: round(t[i] / nt) for i in range(8) if i != 5}, "cycles per tile")
