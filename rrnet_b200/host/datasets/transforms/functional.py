"""datasets/transforms/functional.py:177-262 -- to_heatmap on the GPU (rr_render_targets).

The reference renders one image at a time on DataLoader workers (CPU, a Python loop per object) and
ships the [10,h/4,w/4] map over PCIe; here the annotations go to the device (a few KB) and the maps are
rendered there.  `to_heatmap` keeps the per-sample signature; `to_heatmap_batch` is the collated form
(datasets/drones_det.py:70-94 layout) the training loop should use."""
import torch

from rrnet_b200 import ops


def to_heatmap_batch(annos, n_obj, img_h, img_w, scale_factor=4, cls_num=10):
    """annos [B,max_n,8] (x,y,w,h,score,cls 1-based,..) zero padded, n_obj [B] int32, CUDA tensors ->
    hm [B,cls,h/sf,w/sf], wh [B,max_n,2], ind [B,max_n,1] (float), offset [B,max_n,2], reg_mask [B,max_n,1]."""
    return ops.render_targets(annos, n_obj, img_h, img_w, scale_factor, cls_num)


def to_heatmap(data, scale_factor=4, cls_num=10):
    """data = (img [3,h,w], annos [n,8]) -> (img, annos, hm [cls,h/sf,w/sf], wh [n,2], ind [n,1],
    offset [n,2], reg_mask [n,1] bool), tensors on the device of `annos` moved to CUDA."""
    img, annos = data
    h, w = img.size(1), img.size(2)
    a = annos.detach().float().cuda()
    n = a.size(0)
    if n == 0:
        dev = a.device
        z = torch.zeros
        return img, annos, z(cls_num, h // scale_factor, w // scale_factor, device=dev), z(0, 2, device=dev), \
            z(0, 1, device=dev), z(0, 2, device=dev), z(0, 1, dtype=torch.bool, device=dev)
    if a.size(1) < 8:
        a = torch.cat((a, a.new_zeros(n, 8 - a.size(1))), dim=1)
    n_obj = torch.tensor([n], dtype=torch.int32, device=a.device)
    hm, wh, ind, off, msk = ops.render_targets(a[None, :, :8].contiguous(), n_obj, h, w, scale_factor, cls_num)
    return img, annos, hm[0], wh[0], ind[0], off[0], msk[0] > 0
