"""operators/rrnet_operator.py -- RRNetOperator with the hot-path pieces on the sm_100a kernels.

Hot-path members: criterion's heat-map loss (:55-57, fused sigmoid+clamp+focal fwd/bwd),
generate_bbox (:188-209), _ext_nms (:211-232, Gaussian soft-NMS on the GPU), plus save_result and the
static generate_bbox_target.  The training / evaluation loops are the callers of the path: they follow
the reference step for step; `RRNetOperator(cfg)` builds model / optimiser / scheduler / loaders / DDP from the
configuration like the reference's constructor (the dataset, backbone and logging code are the reference's own
modules, imported from sys.path), or takes them as constructor arguments."""
import os
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.nn.parallel import DistributedDataParallel

from rrnet_b200 import ops
from ..modules.loss.focalloss import FocalLossHM
from ..modules.loss.regl1loss import RegL1Loss


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
    """`RRNetOperator(cfg)` builds everything from the configuration exactly like the reference
    (operators/rrnet_operator.py:23-40 + operators/base_operator.py:11-26): seeds, RRNet(cfg) on
    cfg.Distributed.gpu_id, SyncBatchNorm conversion, Adam, MultiStepLR, make_dataloader(cfg, collate_fn='rrnet'),
    DistributedDataParallel(find_unused_parameters=True).  That is what DistributedWrapper.init_operator
    (operators/distributed_wrapper.py:28-45) calls, so train.py / eval.py run unchanged after dropin.install().
    Every piece can also be passed in (tests, other harnesses); a piece that is passed in is used as it is."""

    def __init__(self, cfg, model=None, optimizer=None, lr_sch=None, training_loader=None,
                 validation_loader=None, logger=None):
        self.cfg = cfg
        dist_cfg = getattr(cfg, "Distributed", None)
        gpu_id = getattr(dist_cfg, "gpu_id", 0)
        built = model is None
        if built:
            from ..models.rrnet import RRNet
            model = RRNet(cfg).cuda(gpu_id)                                         # :26
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model)                  # :27
        if optimizer is None and built:
            optimizer = optim.Adam(model.parameters(), lr=cfg.Train.lr)             # :29
        if lr_sch is None and optimizer is not None and built:
            lr_sch = optim.lr_scheduler.MultiStepLR(optimizer, milestones=cfg.Train.lr_milestones, gamma=0.1)   # :31
        if built and training_loader is None and validation_loader is None:
            from datasets import make_dataloader                                    # the reference's package (sys.path)
            training_loader, validation_loader = make_dataloader(cfg, collate_fn='rrnet')   # :32
        if built:
            # base_operator.py:19-24
            seed = getattr(cfg, "seed", None)
            if seed is not None:
                random.seed(seed)
                torch.manual_seed(seed)
                torch.cuda.manual_seed(seed)
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                model = DistributedDataParallel(model, find_unused_parameters=True, device_ids=[gpu_id])
            else:
                # the reference always runs under DistributedWrapper (a process group exists); without one the
                # model is used bare -- `self.model.module` below goes through _unwrap() for that reason
                pass
        self.model = model
        self.optimizer = optimizer
        self.lr_sch = lr_sch
        self.training_loader = training_loader
        self.validation_loader = validation_loader
        self.logger = logger
        self.hm_focal_loss = FocalLossHM()
        self.l1_loss = RegL1Loss()
        self.main_proc_flag = gpu_id == 0

    def _unwrap(self):
        return getattr(self.model, 'module', self.model)

    @staticmethod
    def save_ckp(models, step, path):
        """base_operator.py:44-51."""
        torch.save(models.state_dict(), os.path.join(path, 'ckp-{}.pth'.format(step)))

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
        """One VisDrone result file: 'x,y,w,h,score,cls,-1,-1' per row (%f x4, %.4f, %d), negative values clamped to 0
        first like the reference (:236).  Written in one call instead of a Python loop over the rows."""
        rows = torch.clamp(pred_bbox.detach().float().cpu(), min=0.).numpy().astype(np.float64)
        table = np.concatenate([rows[:, :5], np.trunc(rows[:, 5:6]), np.full((rows.shape[0], 2), -1.0)], axis=1)
        np.savetxt(file_path, table, fmt=['%f', '%f', '%f', '%f', '%.4f', '%d', '%d', '%d'], delimiter=',')

    # ------------------------------------------------------------------ rrnet_operator.py:104-186
    def _make_logger(self):
        """The reference builds `Logger(cfg)` on the main process (:105-106); its module is used when importable
        (tensorboard + ./log/<prefix>/log.txt).  A logger passed to the constructor wins."""
        if self.logger is not None or not self.main_proc_flag:
            return self.logger
        try:
            from utils.vis.logger import Logger                        # reference module (sys.path)
        except ImportError:
            return None
        return Logger(self.cfg)

    def _targets_on_device(self, imgs, annos, gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks):
        """A batch from the reference's CPU `ToHeatmap` is used as it is.  A batch from `DeferredToHeatmap`
        (dropin.install(gpu_targets=True): the worker only counted the objects) carries an empty heat-map: the
        targets are rendered here, on the device, from the collated annotations (one `rr_render_targets` launch)."""
        if gt_hms.numel() != 0:
            return gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks
        from ..datasets.transforms.functional import to_heatmap_batch
        n_obj = gt_reg_masks.reshape(gt_reg_masks.size(0), -1).sum(1).int().contiguous()
        return to_heatmap_batch(annos.float().contiguous(), n_obj, imgs.size(2), imgs.size(3),
                                self.cfg.Train.scale_factor, self.cfg.num_classes)

    def _log_images(self, outs, imgs, annos):
        """:159-175 -- three pictures of image 0 (stage-1 boxes, stage-2 boxes after soft-NMS, ground truth) through the
        reference's own drawing code; skipped when that code cannot be imported (cv2 / matplotlib missing)."""
        try:
            from datasets.transforms.functional import denormalize
            from utils.vis.annotations import visualize
        except ImportError:
            return None
        s1_pred_bbox, s2_pred_bbox = self.generate_bbox(outs, batch_idx=0)
        img = (denormalize(imgs[0].cpu()).permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
        s2_pred_bbox = self._ext_nms(s2_pred_bbox)
        pics = [visualize(img.copy(), s1_pred_bbox.detach().cpu(), xywh=True, with_score=True),
                visualize(img.copy(), s2_pred_bbox, xywh=True, with_score=True),
                visualize(img.copy(), annos[0, :, :6].cpu(), xywh=False)]
        return {'Train': [torch.from_numpy(p).permute(2, 0, 1).unsqueeze(0).float() / 255. for p in pics]}

    def training_process(self):
        """:104-186 step for step: scheduler first (:117), zero_grad, batch (an out-of-memory batch is skipped, :120-126),
        forward, criterion, loss = hm + 0.1 wh + off + s2 * (step >= 2000) (:132-136), backward, optimiser step; on the main
        process running means + learning rate (+ pictures) go to the logger every cfg.Train.print_interval steps and a
        checkpoint of model.module is written every cfg.Train.checkpoint_interval steps and after the last one (:184-186)."""
        logger = self._make_logger()
        self.model.train()
        names = ('total_loss', 'hm_loss', 'wh_loss', 'off_loss', 's2_reg_loss')
        running = dict.fromkeys(names, 0.0)
        every = getattr(self.cfg.Train, 'print_interval', 0)
        ckp_every = getattr(self.cfg.Train, 'checkpoint_interval', 0)
        iter_num = self.cfg.Train.iter_num
        for step in range(iter_num):
            self.lr_sch.step()
            self.optimizer.zero_grad()
            try:
                batch = self.training_loader.get_batch()
            except RuntimeError as e:
                if 'out of memory' in str(e):
                    print('WARNING: ran out of memory with exception at step {}.'.format(step))
                continue
            imgs, annos = batch[0], batch[1]
            gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks = self._targets_on_device(imgs, annos, *batch[2:7])
            outs = self.model(imgs)
            losses = self.criterion(outs, (gt_hms, gt_whs, gt_inds, gt_offsets, gt_reg_masks, annos))
            hm_loss, wh_loss, off_loss, s2_loss = losses
            s2_factor = 0 if step < 2000 else 1
            loss = hm_loss + (0.1 * wh_loss) + off_loss + s2_loss * s2_factor
            loss.backward()
            self.optimizer.step()
            for key, value in zip(names, (loss, hm_loss, wh_loss, off_loss, s2_loss)):
                running[key] += float(value)
            if not self.main_proc_flag:
                continue
            if every and step % every == every - 1:
                if logger is not None:
                    log_data = {'scalar': {'train/' + k: v / every for k, v in running.items()}}
                    log_data['scalar']['train/lr'] = self.optimizer.param_groups[-1]['lr']
                    if getattr(self.cfg.Train, 'log_images', True):
                        pics = self._log_images(outs, imgs, annos)
                        if pics is not None:
                            log_data['imgs'] = pics
                    logger.log(log_data, step)
                running = dict.fromkeys(names, 0.0)
            if (ckp_every and step % ckp_every == ckp_every - 1) or step == iter_num - 1:
                log_dir = getattr(logger, 'log_dir', None) or os.path.join('./log', str(getattr(self.cfg, 'log_prefix', 'rrnet')))
                os.makedirs(log_dir, exist_ok=True)
                self.save_ckp(self._unwrap(), step, log_dir)
        return running['total_loss']

    # ------------------------------------------------------------------ rrnet_operator.py:246-284, batched
    def detect_multi_scale(self, imgs, scales=(1,), score_thr=0.01, final_nms=True):
        """The body of the reference's evaluation loop for a whole batch (the reference handles image 0 only):
        every scale through the model, stage-2 boxes of ALL images from one generate_bbox launch, boxes back to the
        input resolution, then per image: score-descending order -> Gaussian soft-NMS per class (one launch for all
        (image, class) segments) -> score-descending order.  One device->host copy.
        -> list of CPU tensors [n_i,6] = x,y,w,h,score,cls."""
        B = imgs.size(0) if torch.is_tensor(imgs) else len(imgs)
        rows, owner = [], []
        for scale in scales:
            x = F.interpolate(imgs, scale_factor=scale, mode='bilinear', align_corners=True) if torch.is_tensor(imgs) else imgs
            outs = self.model(x)
            s2_reg, bxyxy, scores, clses = outs[3], outs[4], outs[5], outs[6]
            _, s2 = ops.generate_bbox(bxyxy.contiguous(), s2_reg.contiguous(), scores.contiguous(),
                                      clses.float().contiguous(), scale=float(self.cfg.Train.scale_factor))
            img_idx = bxyxy[:, 0].long()
            if final_nms:                                   # :266-267
                keep = s2[:, 4] > score_thr
                s2, img_idx = s2[keep], img_idx[keep]
            s2 = s2.clone()
            s2[:, :4] = s2[:, :4] / scale                   # :269
            rows.append(s2)
            owner.append(img_idx)
        rows, owner = torch.cat(rows), torch.cat(owner)
        by_score = torch.sort(rows[:, 4], descending=True, stable=True).indices          # :273-274
        rows, owner = rows[by_score], owner[by_score]
        if final_nms and rows.size(0):
            key = owner * 4096 + rows[:, 5].long()          # (image, class) segments, score order kept inside
            grouped = torch.sort(key, stable=True).indices
            rows, owner, key = rows[grouped], owner[grouped], key[grouped]
            _, counts = torch.unique_consecutive(key, return_counts=True)
            seg = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=rows.device)
            seg[1:] = torch.cumsum(counts, 0).int()
            boxes5 = rows[:, :5].clone()
            boxes5[:, 2:4] += boxes5[:, 0:2]                # xywh -> xyxy (:222-223)
            kept, _, cnt = ops.soft_nms_batched(boxes5.contiguous(), seg, 0.5, 0.7, 0.1, 2)
            kept = kept.clone()
            kept[:, 2:4] -= kept[:, 0:2]
            pos = torch.arange(rows.size(0), device=rows.device) - seg[:-1].long().repeat_interleave(counts)
            alive = pos < cnt.long().repeat_interleave(counts)           # the first cnt rows of a segment survive
            rows = torch.cat((kept, rows[:, 5:6]), dim=1)[alive]          # column 5 is constant inside a segment
            owner = owner[alive]
        host_rows, host_owner = rows.cpu(), owner.cpu()
        out = []
        for b in range(B):
            mine = host_rows[host_owner == b]
            out.append(mine[torch.sort(mine[:, 4], descending=True, stable=True).indices])   # :278-279
        return out

    def evaluation_process(self):
        """Multi-scale test of the validation set: `detect_multi_scale` per batch, one result file per image
        ('<cfg.Val.result_dir>/<name>.txt').  Any batch size; with cfg.Val.auto_test the score filter and the final
        soft-NMS are left to the metric code, as in the reference."""
        self.model.eval()
        model_path = getattr(self.cfg.Val, 'model_path', None)
        if model_path:
            self._unwrap().load_state_dict(torch.load(model_path, map_location='cpu'))
        os.makedirs(self.cfg.Val.result_dir, exist_ok=True)
        step, all_step = 0, len(self.validation_loader)
        with torch.no_grad():
            for imgs, annos, names in self.validation_loader:
                step += 1
                dets = self.detect_multi_scale(imgs.cuda(), self.cfg.Val.scales, final_nms=not self.cfg.Val.auto_test)
                for name, det in zip(names, dets):
                    self.save_result(os.path.join(self.cfg.Val.result_dir, name + '.txt'), det)
                print("\r[{}/{}]".format(step, all_step), end='', flush=True)
        print('=> Evaluation Done!')
