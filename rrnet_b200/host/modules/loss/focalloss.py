"""modules/loss/focalloss.py:15-20 -- FocalLossHM."""
import torch.nn as nn

from .functional import focal_loss_for_hm, focal_loss_for_hm_logits


class FocalLossHM(nn.Module):
    def __init__(self):
        super(FocalLossHM, self).__init__()

    def forward(self, out, target):
        return focal_loss_for_hm(out, target)

    @staticmethod
    def from_logits(logits, target):
        """Fused path used by RRNetOperator.criterion: takes the raw heat-map logits."""
        return focal_loss_for_hm_logits(logits, target)
