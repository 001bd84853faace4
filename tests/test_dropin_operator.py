"""`RRNetOperator(cfg)` the way the reference constructs and drives it (operators/distributed_wrapper.py:28-45 ->
operators/rrnet_operator.py:23-40 -> training_process / evaluation_process), after `dropin.install()`.

The GPU box has no reference checkout, so the reference modules that are OUTSIDE the path (backbone, stage-1 conv heads,
dataset / loaders) are stood in for by small stubs registered under the reference's module names; everything on the
path is the mirror + librrnet_b200.so.  Runs in a subprocess: it registers top-level module names (`utils`, `datasets`,
`detectors`) and a process group."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

SCRIPT = textwrap.dedent(r'''
    import os, sys, types, glob
    import torch, torch.nn as nn, torch.distributed as dist
    sys.path.insert(0, %(repo)r)
    NS = types.SimpleNamespace
    workdir = %(workdir)r
    os.chdir(workdir)

    # ---- stubs for the reference modules that are outside the path ------------------------------------------
    def mod(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
        parent, _, leaf = name.rpartition('.')
        if parent:
            setattr(sys.modules[parent], leaf, m)
        return m

    class Backbone(nn.Module):                       # stands in for backbones/hourglass.py: list of num_stacks maps
        def __init__(self, num_stacks):
            super().__init__()
            self.num_stacks = num_stacks
            self.stem = nn.Sequential(nn.Conv2d(3, 256, 7, 4, 3), nn.BatchNorm2d(256))
        def forward(self, x):
            f = self.stem(x)
            return [f * (0.5 + 0.5 * i) for i in range(self.num_stacks)]

    class Heads(nn.Module):                          # stands in for detectors/centernet_detector.py
        def __init__(self, planes, num_stacks, hm=False):
            super().__init__()
            self.detect_layer = nn.ModuleList([nn.Conv2d(256, planes, 1) for _ in range(num_stacks)])
            if hm:
                for l in self.detect_layer:
                    l.bias.data.fill_(-2.19)
        def forward(self, feat, i):
            return self.detect_layer[i](feat)

    class WHHeads(Heads):
        def __init__(self, planes, num_stacks):
            super().__init__(2 * planes, num_stacks)

    for pkg in ('models', 'operators', 'ext', 'ext.nms', 'ext.nms.nms'):
        mod(pkg)
    mod('utils'); mod('utils.model_tools', get_backbone=lambda name, num_stacks=2: Backbone(num_stacks))
    mod('detectors'); mod('detectors.centernet_detector', CenterNetDetector=Heads, CenterNetWHDetector=WHHeads)

    class TrainLoader(object):                       # datasets/dataloader.py:4-40 interface: get_batch()
        def __init__(self, bs, deferred):
            self.bs, self.deferred, self.g = bs, deferred, torch.Generator().manual_seed(5)
        def __len__(self):
            return 4
        def get_batch(self, device='cuda'):
            from rrnet_b200 import synth
            imgs = torch.randn(self.bs, 3, 256, 256, generator=self.g)
            annos, n_obj = synth.pad_annos(synth.train_annos(self.bs, 256, 256, seed=int(torch.randint(0, 1000, (1,), generator=self.g)), n_range=(5, 30)))
            max_n = annos.size(1)
            if self.deferred:                        # what DeferredToHeatmap + collate_fn_ctnet produce
                hms = torch.zeros(self.bs, 0)
                whs, inds, offs = torch.zeros(self.bs, max_n, 2), torch.zeros(self.bs, max_n, 1), torch.zeros(self.bs, max_n, 2)
                masks = (torch.arange(max_n)[None, :, None] < n_obj[:, None, None]).float()
                out = [imgs, annos, hms, whs, inds, offs, masks]
            else:
                from rrnet_b200.host.datasets.transforms.functional import to_heatmap_batch
                out = [imgs, annos] + [t.cpu() for t in to_heatmap_batch(annos.cuda(), n_obj.int().cuda(), 256, 256)]
            return [t.to(device) for t in out] + [['img%%d' %% i for i in range(self.bs)]]

    class ValLoader(object):
        def __len__(self):
            return 2
        def __iter__(self):
            g = torch.Generator().manual_seed(9)
            for i in range(2):
                yield torch.randn(1, 3, 192, 256, generator=g), torch.zeros(1, 3, 8), ['val%%d' %% i]

    made = []
    def make_dataloader(cfg, collate_fn=None):
        made.append(collate_fn)
        return TrainLoader(cfg.Train.batch_size, cfg.deferred), ValLoader()
    mod('datasets', make_dataloader=make_dataloader)

    class Logger(object):                            # utils/vis/logger.py interface: log(dict, step), log_dir
        def __init__(self, cfg):
            self.log_dir = os.path.join('./log', cfg.log_prefix); os.makedirs(self.log_dir, exist_ok=True); self.rows = []
        def log(self, data, step):
            self.rows.append((step, data))
    mod('utils.vis'); mod('utils.vis.logger', Logger=Logger)

    # ---- what eval.py / train.py do -----------------------------------------------------------------------
    import rrnet_b200.host.dropin as dropin
    dropin.WHOLE = tuple(n for n in dropin.WHOLE)    # unchanged; SYMBOLS patch reference modules that do not exist here
    dropin.SYMBOLS = ()
    dropin.install()
    from operators.rrnet_operator import RRNetOperator                     # the reference's import line
    assert RRNetOperator.__module__.startswith('rrnet_b200.host')

    cfg = NS(seed=219, dataset='drones_det', log_prefix='dropin_test', num_classes=10, deferred=%(deferred)r,
             Train=NS(batch_size=2, lr=2.5e-4, lr_milestones=[2, 3], iter_num=3, scale_factor=4, print_interval=2,
                      checkpoint_interval=2, log_images=False),
             Val=NS(model_path=None, auto_test=False, scales=[1, 1.25], result_dir='./results/', batch_size=1),
             Model=NS(backbone='hourglass', num_stacks=2, nms_type_for_stage1='nms', nms_per_class_for_stage1=True),
             Distributed=NS(world_size=1, gpu_id=-1, rank=0, ngpus_per_node=1, dist_url='tcp://127.0.0.1:%(port)d'))

    # DistributedWrapper.init_operator (operators/distributed_wrapper.py:28-45)
    gpu = 0
    cfg.Distributed.gpu_id = gpu
    dist.init_process_group(backend='nccl', init_method=cfg.Distributed.dist_url, world_size=1, rank=0)
    torch.cuda.set_device(gpu)
    op = RRNetOperator(cfg)

    from torch.nn.parallel import DistributedDataParallel
    assert isinstance(op.model, DistributedDataParallel)
    assert made == ['rrnet']
    assert any(isinstance(m, nn.SyncBatchNorm) for m in op.model.modules())
    assert isinstance(op.optimizer, torch.optim.Adam) and op.optimizer.param_groups[0]['lr'] == 2.5e-4
    assert isinstance(op.lr_sch, torch.optim.lr_scheduler.MultiStepLR)
    assert next(op.model.parameters()).is_cuda and op.main_proc_flag

    before = [p.detach().clone() for p in op.model.parameters()]
    op.training_process()
    changed = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, op.model.parameters()))
    assert changed > 0, 'no parameter moved'
    ckps = sorted(os.path.basename(p) for p in glob.glob('./log/dropin_test/ckp-*.pth'))
    assert ckps == ['ckp-1.pth', 'ckp-2.pth'], ckps            # checkpoint_interval = 2 and the last step
    sd = torch.load('./log/dropin_test/ckp-2.pth', map_location='cpu')
    assert 'head_detector.regressor.weight' in sd and not any(k.startswith('module.') for k in sd)

    cfg.Val.model_path = './log/dropin_test/ckp-2.pth'
    op.evaluation_process()
    for name in ('val0', 'val1'):
        rows = open('./results/%%s.txt' %% name).read().strip().splitlines()
        for r in rows[:5]:
            f = r.split(',')
            assert len(f) == 8 and f[6] == '-1' and f[7] == '-1' and 1 <= int(f[5]) <= 10
    dist.destroy_process_group()
    print('dropin operator ok', changed, len(rows))
''')


@pytest.mark.parametrize("deferred", [False, True])
def test_operator_from_cfg_trains_and_evaluates(tmp_path, deferred):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    code = SCRIPT % dict(repo=REPO, workdir=str(tmp_path), deferred=deferred, port=34564 + int(deferred))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "dropin operator ok" in out.stdout, out.stdout[-3000:] + out.stderr[-6000:]
