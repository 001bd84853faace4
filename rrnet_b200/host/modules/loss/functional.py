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
