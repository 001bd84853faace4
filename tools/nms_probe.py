"""Per-kernel device times of the generic NMS at n boxes in one class (ops.KernelTrace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrnet_b200 import ops, synth
for n in (1500, 5000, 20000):
    d = synth.nms_stress_boxes(n, synth.SEED_C5).cuda()
    seg = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    b, s = d[:, :4].contiguous(), d[:, 4].contiguous()
    for _ in range(2):
        ops.nms_batched(b, s, seg, 0.7)
    acc = {}
    for _ in range(5):
        with ops.KernelTrace(16) as kt:
            ops.nms_batched(b, s, seg, 0.7)
        for k, ms in kt.kernels:
            acc[k] = acc.get(k, 0) + ms / 5
    print(n, {k: round(v * 1e3, 1) for k, v in acc.items()}, "us")
z = torch.randn(32, 10, 128, 128, device="cuda") * 2 - 2
annos, n_obj = synth.pad_annos(synth.train_annos(32, 512, 512, synth.SEED_C3))
gt = ops.render_targets(annos.cuda(), n_obj.cuda(), 512, 512)[0]
for _ in range(3):
    ops.focal_fwd_bwd(z, gt); ops.focal_forward(z, gt)
with ops.KernelTrace(16) as kt:
    ops.focal_fwd_bwd(z, gt); ops.focal_forward(z, gt); ops.render_targets(annos.cuda(), n_obj.cuda(), 512, 512)
print({k: round(v * 1e3, 1) for k, v in kt.kernels}, "us (warm L2)")
