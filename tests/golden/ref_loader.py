"""Bring up the UNMODIFIED reference (/root/reference) on CPU for golden generation.

Used only by tests/golden/make_golden.py, in the authoring container.  /root/reference does
not exist on the GPU box, so nothing that runs there imports this file.

Shims:
  1. ext.nms.nms.cpu_nms  <- oracle/_ref build of the reference's own cpu_nms.pyx
     (oracle/build_ref.py), ext.nms.nms.gpu_nms <- stub (needs a GPU);
  2. matplotlib.cm stub (utils/vis/annotations.py:2; matplotlib is not installed);
  3. models.rrnet.get_backbone -> identity module (the real hourglass loads ./hourglass.pth,
     backbones/hourglass.py:209); the "image" argument becomes the list of backbone features;
  4. configs are SimpleNamespace objects (easydict is not installed).
"""
import os
import sys
import types
import warnings

REF = "/root/reference"
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def load():
    """-> namespace with the reference modules and helpers."""
    if not available():
        raise RuntimeError("reference not mounted at " + REF)
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    from oracle import build_ref
    if not build_ref.build():
        raise RuntimeError("could not build oracle/_ref/cpu_nms")
    cpu_nms_mod = build_ref.load()

    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.modules["ext.nms.nms.cpu_nms"] = cpu_nms_mod
    g = types.ModuleType("ext.nms.nms.gpu_nms")
    g.gpu_nms = None
    sys.modules["ext.nms.nms.gpu_nms"] = g
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        m.cm = types.ModuleType("matplotlib.cm")
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.cm"] = m.cm

    import torch.nn as nn
    import models.rrnet as R
    import operators.rrnet_operator as O
    import modules.loss.functional as LF
    import datasets.transforms.functional as TF
    import ext.nms.nms_wrapper as NW
    from ext.nms.nms.py_cpu_nms import py_cpu_nms

    class _Identity(nn.Module):
        def forward(self, feats):
            return feats

    R.get_backbone = lambda name, num_stacks=2: _Identity()

    NS = types.SimpleNamespace
    cfg = NS(num_classes=10, Train=NS(scale_factor=4),
             Model=NS(num_stacks=2, backbone="hourglass", nms_type_for_stage1="nms",
                      nms_per_class_for_stage1=True))

    def make_net(stage1_maps=None):
        """RRNet(cfg).eval(); if stage1_maps=(hm,wh,off) is given, forward_stage1 returns those
        maps for every stack so the post-backbone path runs on chosen inputs."""
        net = R.RRNet(cfg).eval()
        if stage1_maps is not None:
            hm, wh, off = stage1_maps
            net.forward_stage1 = lambda feats: ([hm] * cfg.Model.num_stacks,
                                                [wh] * cfg.Model.num_stacks,
                                                [off] * cfg.Model.num_stacks)
        return net

    fake_op = NS(cfg=cfg, hm_focal_loss=O.FocalLossHM(), l1_loss=O.RegL1Loss(),
                 generate_bbox_target=O.RRNetOperator.generate_bbox_target)

    return NS(R=R, O=O, LF=LF, TF=TF, NW=NW, cpu_nms=cpu_nms_mod, py_cpu_nms=py_cpu_nms,
              cfg=cfg, make_net=make_net, fake_op=fake_op)
