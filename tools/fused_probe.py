import torch, sys
sys.path.insert(0, '/root/repo')
from rrnet_b200 import ops, synth
dev = torch.device('cuda')
B, C = 32, 10
g = torch.Generator().manual_seed(synth.SEED_C3)
z = (torch.randn(B, C, 128, 128, generator=g) * 2 - 2).to(dev)
annos, n_obj = synth.pad_annos(synth.train_annos(B, 512, 512, synth.SEED_C3))
annos, n_obj = annos.to(dev), n_obj.to(dev)
for _ in range(3):
    st = ops.focal_render_forward(z, annos, n_obj, 512, 512)
    gr = ops.focal_render_backward(z, annos, n_obj, 512, 512, st)
    gt = ops.render_targets(annos, n_obj, 512, 512)[0]
    ops.focal_fwd_bwd(z, gt)
torch.cuda.synchronize()
print(annos.shape, n_obj.tolist()[:8])
