"""operators/rrnet_operator.py -- RRNetOperator with the hot-path pieces on the sm_100a kernels.

Hot-path members: criterion's heat-map loss (:55-57, fused sigmoid+clamp+focal fwd/bwd),
generate_bbox (:188-209), _ext_nms (:211-232, Gaussian soft-NMS on the GPU), plus save_result and the
static generate_bbox_target.  The training / evaluation loops are the callers of the path: they follow
the reference step for step but take the model, loaders and optimiser as constructor arguments (the
dataset, backbone and logging code are outside this repo)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from rrnet_b200 import ops
from ..modules.loss.focalloss import FocalLossHM
from ..modules.loss.regl1loss import RegL1Loss


def _box_iou(a, b):
    """torchvision.ops.box_iou (rrnet_operator.py:72) in plain torch: [n,4] x [m,4] -> [n,m]."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


class _Stage2LossFn(torch.autograd.Function):
    """rr_stage2_loss with the reference's gradients: to s2_reg, and to the predicted boxes through the regression
    targets (the reference does not detach them, Appendix A.5)."""

    @staticmethod
    def forward(ctx, s2_reg, bxyxy, gt_annos, bs, scale):
        img = bxyxy[:, 0].detach().contiguous()
        seg = torch.searchsorted(img, torch.arange(bs + 1, dtype=img.dtype, device=img.device)).int()
        parts, g_reg, g_box = ops.stage2_loss(bxyxy.detach(), seg, s2_reg.detach(), gt_annos.detach().float(), scale)
        ctx.g_reg, ctx.g_box = g_reg, g_box
        return parts.sum()

    @staticmethod
    def backward(ctx, g):
        g_bxyxy = torch.cat([torch.zeros_like(ctx.g_box[:, :1]), ctx.g_box * g], dim=1)
        return ctx.g_reg * g, g_bxyxy, None, None, None


class RRNetOperator(object):
    def __init__(self, cfg, model=None, optimizer=None, lr_sch=None, training_loader=None,
                 validation_loader=None, logger=None):
        self.cfg = cfg
        self.model = model
        self.optimizer = optimizer
        self.lr_sch = lr_sch
        self.training_loader = training_loader
        self.validation_loader = validation_loader
        self.logger = logger
        self.hm_focal_loss = FocalLossHM()
        self.l1_loss = RegL1Loss()
        dist_cfg = getattr(cfg, "Distributed", None)
        self.main_proc_flag = getattr(dist_cfg, "gpu_id", 0) == 0

    # ------------------------------------------------------------------ rrnet_operator.py:42-84
    def criterion(self, outs, targets):
        s1_hms, s1_whs, s1_offsets, s2_reg, bxyxy, scores, _ = outs
        gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks, gt_annos = targets
        bs = s1_hms[0].size(0)
        hm_loss = 0
        wh_loss = 0
        off_loss = 0
        for s in range(self.cfg.Model.num_stacks):
            # :55-57 clamp(sigmoid(hm)) + FocalLossHM in one fused forward+backward kernel
            hm_loss += self.hm_focal_loss.from_logits(s1_hms[s], gt_hms) / self.cfg.Model.num_stacks
            wh_loss += self.l1_loss(s1_whs[s], gt_reg_masks, gt_inds, gt_whs) / self.cfg.Model.num_stacks
            off_loss += self.l1_loss(s1_offsets[s], gt_reg_masks, gt_inds, gt_offsets) / self.cfg.Model.num_stacks

        gt_annos[:, :, 2:4] += gt_annos[:, :, 0:2]                       # in place, as the reference (:67)
        # :68-84 for every image in one launch (IoU vs the padded ground truth, max, > 0.5, targets, smooth-L1 / bs)
        s2_reg_loss = _Stage2LossFn.apply(s2_reg, bxyxy, gt_annos, bs, float(self.cfg.Train.scale_factor))
        return hm_loss, wh_loss, off_loss, s2_reg_loss

    @staticmethod
    def generate_bbox_target(ex_rois, gt_rois):
        """:86-102, "+1" width convention."""
        ex_widths = ex_rois[:, 2] - ex_rois[:, 0] + 1.0
        ex_heights = ex_rois[:, 3] - ex_rois[:, 1] + 1.0
        ex_ctr_x = ex_rois[:, 0] + 0.5 * ex_widths
        ex_ctr_y = ex_rois[:, 1] + 0.5 * ex_heights
        gt_widths = gt_rois[:, 2] - gt_rois[:, 0] + 1.0
        gt_heights = gt_rois[:, 3] - gt_rois[:, 1] + 1.0
        gt_ctr_x = gt_rois[:, 0] + 0.5 * gt_widths
        gt_ctr_y = gt_rois[:, 1] + 0.5 * gt_heights
        targets_dx = (gt_ctr_x - ex_ctr_x) / ex_widths
        targets_dy = (gt_ctr_y - ex_ctr_y) / ex_heights
        targets_dw = torch.log(gt_widths / ex_widths)
        targets_dh = torch.log(gt_heights / ex_heights)
        return torch.stack((targets_dx, targets_dy, targets_dw, targets_dh), dim=1)

    # ------------------------------------------------------------------ rrnet_operator.py:188-209
    def generate_bbox(self, outs, batch_idx=0):
        """-> (s1_bboxes [n,6] = x,y,w,h,score,0 ; s2_bboxes [n,6] = x,y,w,h,score,cls+1) of one image,
        input-pixel units.  Does not modify `outs` (the reference aliases bxyxy*4 but that is a temporary)."""
        s1_hms, s1_whs, s1_offsets, s2_reg, bxyxy, scores, clses = outs
        batch_flag = bxyxy[:, 0] == batch_idx
        s1, s2 = ops.generate_bbox(bxyxy[batch_flag].contiguous(), s2_reg[batch_flag].contiguous(),
                                   scores[batch_flag].contiguous(), clses[batch_flag].float().contiguous(),
                                   scale=float(self.cfg.Train.scale_factor))
        return s1, s2

    # ------------------------------------------------------------------ rrnet_operator.py:211-232
    @staticmethod
    def _ext_nms(pred_bbox, per_cls=True):
        """pred_bbox [n,6] xywh,score,cls (any device) -> CPU tensor [n',6] xywh: Gaussian soft-NMS
        (Nt=0.7, threshold=0.1, method=2, sigma=0.5) per class, classes ascending, selection order inside."""
        if pred_bbox.size(0) == 0:
            return pred_bbox
        d = pred_bbox.detach().float().cuda()
        if per_cls:
            order = torch.sort(d[:, 5], stable=True).indices         # class-ascending, original order inside
            d = d[order]
            cls_u, counts = torch.unique_consecutive(d[:, 5], return_counts=True)
            seg = torch.zeros(cls_u.numel() + 1, dtype=torch.int32, device=d.device)
            seg[1:] = torch.cumsum(counts, 0).int()
        else:
            seg = torch.tensor([0, d.size(0)], dtype=torch.int32, device=d.device)
        boxes5 = d[:, :5].clone()
        boxes5[:, 2:4] += boxes5[:, 0:2]                             # xywh -> xyxy (:222-223)
        rows, _, cnt = ops.soft_nms_batched(boxes5, seg, 0.5, 0.7, 0.1, 2)
        seg_h, cnt_h = seg.tolist(), cnt.tolist()
        out = []
        for s, c in enumerate(cnt_h):
            lo = seg_h[s]
            # columns 0..4 come from the reordered rows; column 5 is NOT moved by the reference
            # (cpu_nms.pyx swaps 5 columns only) -- inside one class it is constant anyway
            out.append(torch.cat((rows[lo:lo + c], d[lo:lo + c, 5:6]), dim=1))
        keep = torch.cat(out, dim=0) if out else d[:0]
        keep[:, 2:4] -= keep[:, 0:2]
        return keep.cpu()

    # ------------------------------------------------------------------ rrnet_operator.py:234-244
    @staticmethod
    def save_result(file_path, pred_bbox):
        pred_bbox = torch.clamp(pred_bbox, min=0.)
        with open(file_path, 'w') as f:
            for i in range(pred_bbox.size()[0]):
                bbox = pred_bbox[i]
                line = '%f,%f,%f,%f,%.4f,%d,-1,-1\n' % (
                    float(bbox[0]), float(bbox[1]), float(bbox[2]), float(bbox[3]),
                    float(bbox[4]), int(bbox[5])
                )
                f.write(line)

    # ------------------------------------------------------------------ rrnet_operator.py:104-186
    def training_process(self):
        """The reference's loop: forward, criterion, loss = hm + 0.1*wh + off + s2 (after 2000 steps),
        backward, step.  Logging / checkpointing go through the optional `logger` / cfg.Train fields."""
        self.model.train()
        total_loss = 0
        for step in range(self.cfg.Train.iter_num):
            self.lr_sch.step()
            self.optimizer.zero_grad()
            try:
                imgs, annos, gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks, names = self.training_loader.get_batch()
                targets = gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks, annos
            except RuntimeError as e:
                if 'out of memory' in str(e):
                    print('WARNING: ran out of memory with exception at step {}.'.format(step))
                continue
            outs = self.model(imgs)
            hm_loss, wh_loss, offset_loss, s2_reg_loss = self.criterion(outs, targets)
            s2_factor = 0 if step < 2000 else 1
            loss = hm_loss + (0.1 * wh_loss) + offset_loss + s2_factor * s2_reg_loss
            loss.backward()
            self.optimizer.step()
            total_loss += float(loss)
            if self.main_proc_flag and self.logger is not None and step % self.cfg.Train.print_interval == \
                    self.cfg.Train.print_interval - 1:
                self.logger.log({'scalar': {'train/total_loss': total_loss / self.cfg.Train.print_interval,
                                            'train/hm_loss': float(hm_loss), 'train/wh_loss': float(wh_loss),
                                            'train/off_loss': float(offset_loss),
                                            'train/s2_reg_loss': float(s2_reg_loss)}}, step)
                total_loss = 0
        return total_loss

    # ------------------------------------------------------------------ rrnet_operator.py:246-284
    def evaluation_process(self):
        """Multi-scale test of every validation image, soft-NMS unless cfg.Val.auto_test, result files
        '<result_dir>/<name>.txt'.  Batch-1 like the reference (generate_bbox(outs) defaults to image 0)."""
        self.model.eval()
        model_path = getattr(self.cfg.Val, "model_path", None)
        if model_path:
            state_dict = torch.load(model_path, map_location='cpu')
            getattr(self.model, "module", self.model).load_state_dict(state_dict)
        step = 0
        with torch.no_grad():
            for data in self.validation_loader:
                multi_scale_bboxes = []
                step += 1
                imgs, annos, names = data
                imgs = imgs.cuda()
                for scale in self.cfg.Val.scales:
                    img = F.interpolate(imgs, scale_factor=scale, mode='bilinear', align_corners=True)
                    outs = self.model(img)
                    _, pred_bbox = self.generate_bbox(outs)
                    if not self.cfg.Val.auto_test:
                        pred_bbox = pred_bbox[pred_bbox[:, 4] > 0.01]
                    pred_bbox = pred_bbox.cpu()
                    pred_bbox[:, :4] = pred_bbox[:, :4] / scale
                    multi_scale_bboxes.append(pred_bbox)
                pred_bbox = torch.cat(multi_scale_bboxes, dim=0)
                _, idx = torch.sort(pred_bbox[:, 4], descending=True)
                pred_bbox = pred_bbox[idx]
                if not self.cfg.Val.auto_test:
                    pred_bbox = self._ext_nms(pred_bbox)
                _, idx = torch.sort(pred_bbox[:, 4], descending=True)
                pred_bbox = pred_bbox[idx]
                file_path = os.path.join(self.cfg.Val.result_dir, names[0] + '.txt')
                self.save_result(file_path, pred_bbox)
        print('=> Evaluation Done!')
