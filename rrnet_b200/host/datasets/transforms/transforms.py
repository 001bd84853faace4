"""datasets/transforms/transforms.py:154-161 -- ToHeatmap, and its deferred form for DataLoader workers."""
import torch

from . import functional as F


class ToHeatmap(object):
    """Per-sample target render on the GPU (main process only: it touches CUDA, so it must not run inside the
    reference's forked DataLoader workers -- use DeferredToHeatmap there)."""

    def __init__(self, scale_factor=4, cls_num=10):
        self.scale_factor = scale_factor
        self.cls_num = cls_num

    def __call__(self, data):
        img, annos, hm, wh, ind, offset, reg_mask = F.to_heatmap(data, self.scale_factor, self.cls_num)
        return img, annos, hm, wh, ind, offset, reg_mask


class DeferredToHeatmap(object):
    """Stands in for ToHeatmap at the end of cfg.Train.transforms (configs/rrnet_config.py:48) when the targets are to
    be rendered on the GPU.  The reference's transform pipeline runs in forked DataLoader workers (num_workers=4,
    pin_memory=True, datasets/__init__.py:23-28) where CUDA cannot be used, so the worker only passes the annotations
    on: the heat-map is an EMPTY tensor, wh / ind / offset are zero rows and reg_mask is one row of ones per object.
    `collate_fn_ctnet` (datasets/drones_det.py:70-94) pads these like the real ones; the training loop
    (RRNetOperator._targets_on_device) sees the empty heat-map, counts the objects from the collated reg_mask and
    renders all targets of the batch with one `rr_render_targets` launch after the H2D copy of the annotations."""

    def __init__(self, scale_factor=4, cls_num=10):
        self.scale_factor = scale_factor
        self.cls_num = cls_num

    def __call__(self, data):
        img, annos = data[0], data[1]
        n = annos.size(0)
        return (img, annos, torch.zeros(0), torch.zeros(n, 2), torch.zeros(n, 1), torch.zeros(n, 2), torch.ones(n, 1))
