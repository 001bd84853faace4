"""modules/loss/regl1loss.py:5-17 -- RegL1Loss (SURVEY 8f "next" row; plain torch, but gathers the K
points straight from the NCHW map instead of permuting the whole map first)."""
import torch.nn as nn
import torch.nn.functional as F


class RegL1Loss(nn.Module):
    def __init__(self):
        super(RegL1Loss, self).__init__()

    def forward(self, output, mask, ind, target):
        b, c, h, w = output.shape
        idx = ind.long().view(b, 1, -1).expand(b, c, -1)           # ind arrives as float [B,max_n,1]
        pred = output.reshape(b, c, h * w).gather(2, idx).permute(0, 2, 1)    # [B,max_n,c]
        mask = mask.float().view(b, -1, 1).expand_as(pred)
        loss = F.l1_loss(pred * mask, target * mask, reduction="sum")
        loss = loss / (mask.sum() + 1e-4)
        return loss
