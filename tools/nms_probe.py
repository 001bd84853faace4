import sys, torch
sys.path.insert(0, '/root/repo')
from rrnet_b200 import ops, synth
dev = torch.device('cuda')
d = synth.nms_stress_boxes(20000, synth.SEED_C5).to(dev)
seg = torch.tensor([0, 20000], dtype=torch.int32, device=dev)
boxes, scores = d[:, :4].contiguous(), d[:, 4].contiguous()
d5 = synth.nms_stress_boxes(5000, synth.SEED_C5 + 1).to(dev)
seg5 = torch.tensor([0, 5000], dtype=torch.int32, device=dev)
for _ in range(2):
    k, c = ops.nms_batched(boxes, scores, seg, 0.7, 0, False)
    r, s, c5 = ops.soft_nms_batched(d5, seg5, 0.5, 0.7, 0.1, 2)
torch.cuda.synchronize()
print(int(c[0]), int(c5[0]))
