"""Prints the actual errors of RRNetOperator.criterion (losses, gradients) against the reference golden, per term."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.conftest import load_golden, rel_err
from tests import test_host_api as T
from rrnet_b200 import synth
host = T.NS()
from rrnet_b200.host.models.rrnet import RRNet
from rrnet_b200.host.operators.rrnet_operator import RRNetOperator
host.RRNet, host.RRNetOperator = RRNet, RRNetOperator
g = load_golden("criterion")
B, C, H, W, K = g["shape"].tolist(); seed = int(g["seed"])
x = {k: v.cuda() for k, v in synth.eval_inputs(B, H, W, K, seed).items()}
hm = x["hm"].clone().requires_grad_(True); wh = x["wh"].clone().requires_grad_(True); off = x["off"].clone().requires_grad_(True)
net = T.make_net(host, hm, wh, off, seed)
op = RRNetOperator(T.CFG, model=net)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
outs = net([x["feat"], x["feat"]], k=K)
annos = torch.from_numpy(g["annos"]).cuda()
from rrnet_b200.host.datasets.transforms.functional import to_heatmap_batch
n_obj = torch.from_numpy(g["n_obj"]).int().cuda()
t = to_heatmap_batch(annos, n_obj, H * 4, W * 4)
losses = op.criterion(outs, (*t, annos.clone()))
got = np.array([float(l) for l in losses])
print("losses got", got, "ref", g["losses"], "rel", np.abs(got - g["losses"]) / np.abs(g["losses"]))
for name, ls, w in (("hm", losses[0], 1.0), ("wh", losses[1], 0.1), ("off", losses[2], 1.0), ("s2", losses[3], 1.0)):
    for p in (hm, wh, off):
        p.grad = None
    (w * ls).backward(retain_graph=True)
    print(name, "grad norms", [None if p.grad is None else float(p.grad.abs().max()) for p in (hm, wh, off)])
for p in (hm, wh, off):
    p.grad = None
(losses[0] + 0.1 * losses[1] + losses[2] + losses[3]).backward()
for nm, p, ref in (("hm", hm, g["grad_hm"]), ("wh", wh, g["grad_wh"]), ("off", off, g["grad_off"])):
    a = p.grad.cpu().numpy().astype(np.float64); b = ref.astype(np.float64)
    d = np.abs(a - b)
    i = np.unravel_index(d.argmax(), d.shape)
    print(nm, "max|ref|", np.abs(b).max(), "max abs err", d.max(), "at", i, "ref there", b[i], "got", a[i],
          "rel(max-normalised)", d.max() / np.abs(b).max(), "rel_err floor1e-3", rel_err(a, b, floor=1e-3), "nonzero", (b != 0).sum())
    nz = b != 0
    r = d[nz] / np.abs(b[nz])
    print("   per-element rel: max", r.max(), "99.9pct", np.quantile(r, 0.999), "median", np.median(r))
