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
           ("utils.metrics.metrics", "get_tp", "rrnet_b200.host.utils.metrics.metrics"))

# The reference's `ToHeatmap` / `to_heatmap` run per sample inside forked DataLoader workers
# (configs/rrnet_config.py:48, datasets/__init__.py:23-28: num_workers=4, pin_memory=True), where CUDA cannot be
# initialised -- so they are NOT replaced by default: the CPU render stays in the workers.  With
# install(gpu_targets=True) `ToHeatmap` becomes `DeferredToHeatmap` (the worker only forwards the annotations) and
# RRNetOperator.training_process renders the whole batch on the GPU after collation.  Must be installed before
# configs.rrnet_config is imported (the config instantiates the transform).
GPU_TARGET_SYMBOLS = (("datasets.transforms.transforms", "ToHeatmap", "rrnet_b200.host.datasets.transforms.transforms",
                       "DeferredToHeatmap"),
                      ("datasets.transforms", "ToHeatmap", "rrnet_b200.host.datasets.transforms.transforms",
                       "DeferredToHeatmap"))


def install(gpu_targets=False):
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
    if gpu_targets:
        for ref_mod, symbol, ours, ours_symbol in GPU_TARGET_SYMBOLS:
            target = importlib.import_module(ref_mod)
            setattr(target, symbol, getattr(importlib.import_module(ours), ours_symbol))
            done.append((ref_mod + "." + symbol, "symbol patched (deferred GPU render)"))
    return done
