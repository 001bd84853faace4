"""modules/loss/functional.py:25-51 -- focal_loss_for_hm, fused with the caller's
clamp(sigmoid(logits), 1e-4, 1-1e-4) (operators/rrnet_operator.py:55)."""
import torch

from rrnet_b200 import ops


class _FocalHMFromLogits(torch.autograd.Function):
    """loss(logits, gt): one fused forward+backward launch (rr_focal_fwd_bwd); backward scales the
    stored gradient by the upstream scalar.  gt receives no gradient (it is a target)."""

    @staticmethod
    def forward(ctx, logits, gt):
        need_grad = logits.requires_grad
        if need_grad:
            stats, grad = ops.focal_fwd_bwd(logits.detach(), gt.detach(), 1.0)
            ctx.save_for_backward(grad)
        else:
            stats = ops.focal_forward(logits.detach(), gt.detach())
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def focal_loss_for_hm_logits(logits, gt):
    """sigmoid + clamp + focal_loss_for_hm in one kernel; `logits` are the raw heat-map outputs."""
    return _FocalHMFromLogits.apply(logits, gt)


def focal_loss_for_hm(pred, gt):
    """Reference signature: pred = clamp(sigmoid(logits), 1e-4, 1-1e-4) (batch x c x h x w), gt same shape.
    The kernel works on logits; a probability input is mapped back with logit(p) (exact only inside the
    clamp, which is where the reference itself has non-zero gradient).  Prefer focal_loss_for_hm_logits."""
    p = pred.clamp(1e-4, 1 - 1e-4)
    return focal_loss_for_hm_logits(torch.log(p) - torch.log1p(-p), gt)


class _FocalHMFromAnnos(torch.autograd.Function):
    """Heat-map focal loss from the padded annotations: target render fused into the loss kernels, the
    [B,cls,h,w] target map is never materialised (rr_focal_render_fwd_bwd, or _forward when no gradient is needed)."""

    @staticmethod
    def forward(ctx, logits, annos, n_obj, img_h, img_w, scale_factor):
        if logits.requires_grad:       # loss and gradient in one pass over the logits (rr_focal_render_fwd_bwd)
            stats, grad = ops.focal_render_fwd_bwd(logits.detach(), annos, n_obj, img_h, img_w, 1.0, scale_factor)
            ctx.save_for_backward(grad)
        else:
            stats = ops.focal_render_forward(logits.detach(), annos, n_obj, img_h, img_w, scale_factor)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def focal_loss_for_hm_from_annos(logits, annos, n_obj, img_h, img_w, scale_factor=4):
    """criterion's heat-map term without a rendered target: logits [B,cls,h,w], annos [B,max_n,8] (the
    collate layout, datasets/drones_det.py:70-94), n_obj [B] int32."""
    return _FocalHMFromAnnos.apply(logits, annos, n_obj, img_h, img_w, scale_factor)
