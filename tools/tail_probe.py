"""Runs the fused stage-1 tail + eval path a few times at config 2 (for ncu captures of tail_conv_collect_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrnet_b200 import ops, synth
B, C, H, W, K = 8, 10, 272, 480, 1500
x = synth.eval_inputs(B, H, W, K, synth.SEED_C2)
dev = torch.device("cuda")
feat = x["feat"].to(dev)
t = torch.relu(feat)
g = torch.Generator().manual_seed(1)
w = (torch.randn(C, 256, 1, 1, generator=g) * 0.125).to(dev)
b = torch.full((C,), -2.19, device=dev)
folded = ops.head_fold({k: v.to(dev) for k, v in synth.head_params(synth.SEED_C2).items()})
path = ops.EvalPath(B, C, H, W, K, folded, feat_is_relu=True)
for _ in range(4):
    path.forward_from_tail(t, w, b, x["wh"].to(dev), x["off"].to(dev), t)
torch.cuda.synchronize()
print("rois", path.results()["n"])
