"""Two batches in flight: steps alternate between two streams (own buffers, own CUDA graph each), the persistent
kernels leave `reserve` SMs free so that the 8-CTA latency chains (decode, NMS, RoIAlign prep) of one batch run next
to the tile RoIAlign / head kernels of the other.  Prints images/s per (streams, reserve).  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rrnet_b200 import ops, synth, _lib


def run(n_streams, reserve, steps=100):
    w = bench.WORKLOAD
    B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    assert L.rr_set_sm_reserve(reserve) == 0
    x = synth.eval_inputs(B, H, W, K, synth.SEED_C2)
    d = {k: v.to(dev) for k, v in x.items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    paths, graphs = [], []
    for s in streams:
        with torch.cuda.stream(s):
            p = ops.EvalPath(B, C, H, W, K, folded, device=dev)
            g = p.capture(d["hm"], d["wh"], d["off"], d["feat"])
        paths.append(p); graphs.append(g)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    def loop(n):
        for s in streams:
            s.wait_stream(main)
        for i in range(n):
            with torch.cuda.stream(streams[i % n_streams]):
                graphs[i % n_streams].replay()
        for s in streams:
            main.wait_stream(s)
    loop(10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loop(steps)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    ref = paths[0].results()["reg"]
    same = all(torch.equal(p.results()["reg"], ref) for p in paths)
    print("streams %d reserve %2d: %.3f ms/step  %.0f images/s  (results identical across streams: %s)" % (
        n_streams, reserve, ms, B / ms * 1e3, same), flush=True)
    L.rr_set_sm_reserve(0)


if __name__ == "__main__":
    combos = [(1, 0)] + [(2, r) for r in (0, 4, 8, 12, 16, 24)] + [(3, 12)]
    if len(sys.argv) > 1:                              # e.g.  python tools/pipeline_probe.py 2:6 2:10 3:8
        combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]]
    for n_streams, reserve in combos:
        run(n_streams, reserve)
