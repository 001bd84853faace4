"""modules/loss/regl1loss.py:5-17 -- RegL1Loss (SURVEY 8f "next" row).  Forward and backward are one CUDA launch
(`rr_regl1_fwd_bwd`): the <= B*max_n*c predictions are read straight from the NCHW map, no permuted copy of the map,
no gather / expand / multiply temporaries, and the gradient is written as a zero-filled map with the few entries
scattered in."""
import torch
import torch.nn as nn

from .... import ops


class _RegL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, mask, ind, target):
        loss, grad = ops.regl1_fwd_bwd(output.detach(), mask, ind, target.detach(), 1.0, want_grad=output.requires_grad)
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None


class RegL1Loss(nn.Module):
    def __init__(self):
        super(RegL1Loss, self).__init__()

    def forward(self, output, mask, ind, target):
        return _RegL1Fn.apply(output, mask, ind, target)
