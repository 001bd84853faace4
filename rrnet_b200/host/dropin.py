"""Install the mirror in front of an unmodified RRNet checkout.

The reference must be importable (its root on sys.path).  Modules that consist of the hot path only are replaced
wholesale; modules that also define things outside the path get single symbols patched, so the reference's other
operators and losses keep importing what they import today."""
import importlib
import sys

WHOLE = ("models.rrnet", "operators.rrnet_operator", "detectors.fasterrcnn_detector",
         "ext.nms.nms_wrapper", "ext.nms.nms.gpu_nms", "ext.nms.nms.cpu_nms", "ext.nms.nms.py_cpu_nms")

SYMBOLS = (("modules.loss.focalloss", "FocalLossHM", "rrnet_b200.host.modules.loss.focalloss"),
           ("modules.loss.functional", "focal_loss_for_hm", "rrnet_b200.host.modules.loss.functional"),
           ("modules.loss.regl1loss", "RegL1Loss", "rrnet_b200.host.modules.loss.regl1loss"),
           ("datasets.transforms.functional", "to_heatmap", "rrnet_b200.host.datasets.transforms.functional"),
           ("datasets.transforms.transforms", "ToHeatmap", "rrnet_b200.host.datasets.transforms.transforms"),
           ("utils.metrics.metrics", "get_tp", "rrnet_b200.host.utils.metrics.metrics"))


def install():
    """-> list of (reference name, what was done).  Idempotent."""
    done = []
    for name in WHOLE:
        sys.modules[name] = importlib.import_module("rrnet_b200.host." + name)
        done.append((name, "module replaced"))
    for name in WHOLE:                                   # `import models.rrnet` also looks the child up on its parent
        parent, _, child = name.rpartition(".")
        setattr(importlib.import_module(parent), child, sys.modules[name])
    for ref_mod, symbol, ours in SYMBOLS:
        target = importlib.import_module(ref_mod)                 # the reference's own module
        setattr(target, symbol, getattr(importlib.import_module(ours), symbol))
        done.append((ref_mod + "." + symbol, "symbol patched"))
    return done
