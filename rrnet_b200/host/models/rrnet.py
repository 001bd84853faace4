"""models/rrnet.py -- RRNet with the post-backbone path on the sm_100a kernels.

Same attributes (num_stacks, num_classes, nms_type, nms_per_class, backbone, hm, wh, offset_reg,
head_detector), same methods and return types as the reference class (models/rrnet.py:11-157).
The backbone and the three stage-1 convolution heads are outside this path (cuDNN): they are taken
from the reference when it is importable (utils.model_tools.get_backbone, detectors.centernet_detector)
or passed in as modules."""
import torch
import torch.nn as nn

from rrnet_b200 import ops
from ..detectors.fasterrcnn_detector import FasterRCNNDetector
from ..ext.nms.nms_wrapper import soft_nms


def _reference_stage1(cfg):
    try:
        from utils.model_tools import get_backbone                       # reference modules (sys.path)
        from detectors.centernet_detector import CenterNetDetector, CenterNetWHDetector
    except ImportError as e:
        raise ImportError("RRNet needs backbone/hm/wh/offset_reg modules: pass them in, or put the reference "
                          "repo on sys.path so its backbones/ and detectors/ can be imported (%s)" % e)
    ns = cfg.Model.num_stacks
    return (get_backbone(cfg.Model.backbone, num_stacks=ns),
            CenterNetDetector(planes=cfg.num_classes, num_stacks=ns, hm=True),
            CenterNetWHDetector(planes=1, num_stacks=ns),
            CenterNetDetector(planes=2, num_stacks=ns))


class RRNet(nn.Module):
    def __init__(self, cfg, backbone=None, hm=None, wh=None, offset_reg=None):
        super(RRNet, self).__init__()
        self.num_stacks = cfg.Model.num_stacks
        self.num_classes = cfg.num_classes
        self.nms_type = cfg.Model.nms_type_for_stage1
        self.nms_per_class = cfg.Model.nms_per_class_for_stage1
        if backbone is None or hm is None or wh is None or offset_reg is None:
            backbone, hm, wh, offset_reg = _reference_stage1(cfg)
        self.backbone = backbone
        self.hm = hm
        self.wh = wh
        self.offset_reg = offset_reg
        self.head_detector = FasterRCNNDetector()

    # ------------------------------------------------------------------ models/rrnet.py:25-54
    def forward(self, x, k=1500):
        pre_feat = self.backbone(x)
        feat = pre_feat[-1]
        fused = (not self.training) and not torch.is_grad_enabled() and self.nms_type != 'soft_nms' \
            and self.nms_per_class and feat.size(1) == 256
        tail = self._hm_tail() if (fused and self.fuse_hm_tail) else None
        hms, whs, offsets = self.forward_stage1(pre_feat, skip_last_hm=tail is not None)
        if fused:
            # decode -> per-class NMS -> RoIAlign -> head in one C-ABI call, one host sync for N.  RoIAlign reads
            # relu(pre_feat[-1]); forward_stage1 has just built that very tensor for the stage-1 heads (:144), so it is
            # passed on and the kernel skips its own ReLU
            path = self._eval_path((feat.size(0), self.num_classes, feat.size(2), feat.size(3)), k, feat.device)
            feat = self._relu_last
            if tail is not None:
                # the heat-map head's last 1x1 conv is fused with the decode's pass over the logits: the
                # map is written once for the caller (`hms`) and never read back
                body, last = tail
                t = body(self._relu_last)
                hms.append(path.forward_from_tail(t, last.weight, last.bias, whs[-1], offsets[-1], feat).clone())
            else:
                path.forward(hms[-1], whs[-1], offsets[-1], feat)
            self.__dict__.pop('_relu_last', None)               # do not keep the 1 GB activation alive between calls
            r = path.results()
            # the path's buffers are reused by the next forward of the same shape: hand out copies (a few hundred KB)
            return hms, whs, offsets, r["reg"].clone(), r["bxyxy"].clone(), r["scores"].clone(), r["clses"].clone()
        bboxs = self.transform_bbox(hms[-1], whs[-1], offsets[-1], k)     # (bs, k, 6)
        if self.nms_type != 'soft_nms' and self.nms_per_class:
            bxyxys, scores, clses = self._nms_batch(bboxs)                # all images at once, one host sync
        else:
            bxyxys, scores, clses = [], [], []
            for b_idx in range(bboxs.size(0)):
                bbox = self.nms(bboxs[b_idx])
                xyxy = bbox[:, :4]
                scores.append(bbox[:, 4])
                clses.append(bbox[:, 5])
                batch_idx = torch.ones((xyxy.size(0), 1), device=xyxy.device) * b_idx
                bxyxys.append(torch.cat((batch_idx, xyxy), dim=1))
            bxyxys = torch.cat(bxyxys, dim=0)
            scores = torch.cat(scores, dim=0)
            clses = torch.cat(clses, dim=0)
        roi_feat = _RoIAlignReLU.apply(feat, bxyxys)
        stage2_reg = self.forward_stage2(roi_feat)
        return hms, whs, offsets, stage2_reg, bxyxys, scores, clses

    fuse_hm_tail = True           # eval: fuse the heat-map head's final 1x1 conv into the decode (when the head has that shape)

    def _hm_tail(self):
        """(3x3 conv + ReLU module, final nn.Conv2d 1x1) of the last stack's heat-map head when it is the reference's
        CenterNetDetector layout (detectors/centernet_detector.py:11-15) and fits rr_hm_tail_collect, else None."""
        layers = getattr(self.hm, 'detect_layer', None)
        if layers is None or len(layers) < self.num_stacks:
            return None
        seq = layers[self.num_stacks - 1]
        if not isinstance(seq, nn.Sequential) or len(seq) != 2:
            return None
        body, last = seq[0], seq[1]
        ok = isinstance(last, nn.Conv2d) and tuple(last.kernel_size) == (1, 1) and tuple(last.stride) == (1, 1) \
            and last.groups == 1 and last.bias is not None and last.out_channels == self.num_classes \
            and last.out_channels <= 16 and last.in_channels <= 1024 and last.weight.is_cuda
        return (body, last) if ok else None

    _EVAL_PATH_SLOTS = 8          # multi-scale test: one pre-allocated path per scale (cfg.Val.scales has 6)

    def _eval_path(self, hm_shape, k, device):
        """Pre-allocated outputs + workspace per (B,C,H,W,k,device), reused across forwards (least recently used one
        dropped beyond _EVAL_PATH_SLOTS); only the folded head weights are refreshed (they follow the parameters'
        versions, FasterRCNNDetector.folded)."""
        cache = self.__dict__.setdefault('_eval_paths', {})
        key = (tuple(hm_shape), int(k), str(device))
        path = cache.pop(key, None)
        folded = self.head_detector.folded()
        if path is None:
            B, C, H, W = hm_shape
            path = ops.EvalPath(B, C, H, W, k, folded, device=device, feat_is_relu=True)
            while len(cache) >= self._EVAL_PATH_SLOTS:
                cache.pop(next(iter(cache)))
        path.folded = folded
        cache[key] = path                      # re-inserted last = most recently used
        return path

    def _nms_batch(self, bboxs):
        """The per-image loop of the reference's forward (models/rrnet.py:37-49) + its per-class hard NMS (:56-80) for the
        whole batch: bboxs [B,K,6] -> bxyxys [N,5] (image index, box), scores [N], clses [N]; images ascending, classes
        ascending inside an image, scores descending inside a class - the reference's concatenation order.  One launch and
        one host sync (for N) instead of a Python loop with two syncs per image.  When bboxs carries gradients (training:
        criterion back-propagates through the kept boxes, rrnet_operator.py:82-83) the kept rows are gathered from bboxs by
        index, so autograd keeps working."""
        B, K = bboxs.size(0), bboxs.size(1)
        device = bboxs.device
        if not (bboxs.requires_grad and torch.is_grad_enabled()):
            bx, sc, cl, counts = ops.stage1_nms(bboxs.detach().contiguous(), self.num_classes, 0.7)
            n = int(counts[B].item())
            return bx[:n], sc[:n], cl[:n]
        d = bboxs.detach().reshape(B * K, 6)
        img = torch.arange(B, device=device).repeat_interleave(K)
        key = img * self.num_classes + d[:, 5].long()                     # one segment per (image, class)
        order = torch.sort(key, stable=True).indices                      # stable: score order inside a segment survives
        _, counts = torch.unique_consecutive(key[order], return_counts=True)
        seg = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=device)
        seg[1:] = torch.cumsum(counts, 0).int()
        ds = d[order]
        keep_idx, keep_cnt = ops.nms_batched(ds[:, :4].contiguous(), ds[:, 4].contiguous(), seg, 0.7)
        # a segment's kept rows sit at the start of its own range of keep_idx
        start = seg[:-1].long().repeat_interleave(counts, output_size=B * K)
        kept = torch.arange(B * K, device=device) - start < keep_cnt.long().repeat_interleave(counts, output_size=B * K)
        rows = order[keep_idx[:B * K][kept].long()]                        # the one host sync: N is data dependent
        sel = bboxs.reshape(B * K, 6)[rows]
        bxyxys = torch.cat(((rows // K).to(sel.dtype)[:, None], sel[:, :4]), dim=1)
        return bxyxys, sel[:, 4], sel[:, 5]

    # ------------------------------------------------------------------ models/rrnet.py:56-80
    def nms(self, bbox):
        """bbox [K,6] (x1,y1,x2,y2,score,cls) -> kept rows: classes ascending, score descending inside.
        When bbox carries gradients (training: criterion back-propagates through the kept boxes,
        rrnet_operator.py:82-83) the rows are gathered from bbox by index so autograd keeps working."""
        device = bbox.device
        if bbox.requires_grad and torch.is_grad_enabled() and self.nms_type != 'soft_nms':
            d = bbox.detach()
            if self.nms_per_class:
                order = torch.sort(d[:, 5], stable=True).indices
                cls_u, counts = torch.unique_consecutive(d[order, 5], return_counts=True)
                seg = torch.zeros(cls_u.numel() + 1, dtype=torch.int32, device=device)
                seg[1:] = torch.cumsum(counts, 0).int()
            else:
                order = torch.arange(d.size(0), device=device)
                seg = torch.tensor([0, d.size(0)], dtype=torch.int32, device=device)
            ds = d[order]
            keep_idx, keep_cnt = ops.nms_batched(ds[:, :4].contiguous(), ds[:, 4].contiguous(), seg, 0.7)
            seg_h, cnt_h = seg.tolist(), keep_cnt.tolist()
            rows = torch.cat([keep_idx[seg_h[s]: seg_h[s] + c] for s, c in enumerate(cnt_h)]).long()
            return bbox[order[rows]]
        if self.nms_type == 'soft_nms':
            if self.nms_per_class:
                keep = [torch.from_numpy(soft_nms(bbox[bbox[:, 5] == c].detach().cpu().numpy(), Nt=0.7,
                                                  threshold=0.1, method=2)).to(device)
                        for c in bbox[:, 5].unique()]
                return torch.cat(keep)
            return torch.from_numpy(soft_nms(bbox.detach().cpu().numpy(), Nt=0.7, threshold=0.1, method=2)).to(device)
        if self.nms_per_class:
            bx, sc, cl, counts = ops.stage1_nms(bbox.detach().unsqueeze(0).contiguous(), self.num_classes, 0.7)
            n = int(counts[0].item())
            return torch.cat((bx[:n, 1:], sc[:n, None], cl[:n, None]), dim=1)
        keep_idx = ops.nms(bbox[:, :4].contiguous(), bbox[:, 4].contiguous(), 0.7)
        return bbox[keep_idx]

    # ------------------------------------------------------------------ models/rrnet.py:83-115
    @staticmethod
    def _gather_feat(feat, ind, mask=None):
        dim = feat.size(2)
        ind = ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim)
        feat = feat.gather(1, ind)
        if mask is not None:
            mask = mask.unsqueeze(2).expand_as(feat)
            feat = feat[mask]
            feat = feat.view(-1, dim)
        return feat

    def _topk(self, scores, k=1500):
        """scores [B,C,H,W] (already sigmoid-ed) -> topk_score [B,k], topk_inds [B,k] int64 (y*W+x),
        topk_clses [B,k] int32, topk_ys, topk_xs [B,k] float: one global top-k per image (== the
        reference's two-stage top-k; ties ordered by flat index)."""
        dets, inds = ops.decode_topk(scores, None, None, k, raw_scores=True)
        return dets[..., 4], inds, dets[..., 5].int(), dets[..., 1], dets[..., 0]

    def _transpose_and_gather_feat(self, feat, ind):
        """feat [B,ch,H,W], ind [B,K] -> [B,K,ch]; gathers from the NCHW map directly (the reference
        permutes the whole map to NHWC first)."""
        b, ch = feat.size(0), feat.size(1)
        idx = ind.view(b, 1, -1).expand(b, ch, -1)
        return feat.reshape(b, ch, -1).gather(2, idx).permute(0, 2, 1).contiguous()

    # ------------------------------------------------------------------ models/rrnet.py:117-138
    def transform_bbox(self, hm, wh, offset, k=250):
        """hm LOGITS [B,C,H,W], wh, offset [B,2,H,W] -> [B,k,6] = x1,y1,x2,y2,score,cls (stride-4 units)."""
        if torch.is_grad_enabled() and (wh.requires_grad or offset.requires_grad or hm.requires_grad):
            return _DecodeFn.apply(hm, wh, offset, k)
        dets, _ = ops.decode_topk(hm, wh, offset, k, want_inds=False)
        return dets

    # ------------------------------------------------------------------ models/rrnet.py:140-157
    def forward_stage1(self, feats, skip_last_hm=False):
        """skip_last_hm: the caller produces the last stack's heat map itself (fused tail, `forward`)."""
        hms, whs, offsets = [], [], []
        for i in range(self.num_stacks):
            feat = torch.relu(feats[i])
            if i == self.num_stacks - 1:
                self.__dict__['_relu_last'] = feat              # reused by the fused eval path (plain attribute, not a buffer)
            if not (skip_last_hm and i == self.num_stacks - 1):
                hms.append(self.hm(feat, i))
            whs.append(self.wh(feat, i))
            offsets.append(self.offset_reg(feat, i))
        return hms, whs, offsets

    def forward_stage2(self, feats,):
        return self.head_detector(feats)


class _RoIAlignReLU(torch.autograd.Function):
    """roi_align(relu(feat), rois, (3,3)) with the fused kernel; the backward (training only) is the tile-centric
    gather kernel `rr_roi_align_backward`: ReLU mask included, no atomics on the tile path."""

    @staticmethod
    def forward(ctx, feat, rois):
        ctx.save_for_backward(feat, rois)
        return ops.roi_align(feat.detach(), rois.detach(), relu=True)

    @staticmethod
    def backward(ctx, grad_out):
        feat, rois = ctx.saved_tensors
        return ops.roi_align_backward(feat.detach(), rois.detach(), grad_out.contiguous(), relu=True), None


class _DecodeFn(torch.autograd.Function):
    """transform_bbox with gradients to wh / offset / hm (the reference's top-k + gathers are
    differentiable and criterion back-propagates through the stage-1 boxes).  Forward is the decode
    kernel; backward is a K-point scatter in torch."""

    @staticmethod
    def forward(ctx, hm, wh, off, k):
        dets, inds = ops.decode_topk(hm.detach(), wh.detach(), off.detach(), k)
        ctx.save_for_backward(dets, inds, wh)
        ctx.hm_shape = hm.shape
        ctx.mark_non_differentiable(inds)
        return dets

    @staticmethod
    def backward(ctx, g):
        dets, inds, wh = ctx.saved_tensors
        B, C, H, W = ctx.hm_shape
        g = g.contiguous()
        wh_flat = wh.detach().reshape(B, 2, H * W)
        w_raw = wh_flat.gather(2, inds[:, None, :].expand(B, 2, -1))             # [B,2,K]
        passes = (w_raw >= 0).float()                                             # clamp(min=0) backward
        g_off = torch.stack((g[..., 0] + g[..., 2], g[..., 1] + g[..., 3]), dim=1)            # [B,2,K]
        g_wh = torch.stack((0.5 * (g[..., 2] - g[..., 0]), 0.5 * (g[..., 3] - g[..., 1])), dim=1) * passes
        idx = inds[:, None, :].expand(B, 2, -1)
        d_off = torch.zeros(B, 2, H * W, device=g.device).scatter_add_(2, idx, g_off).view(B, 2, H, W)
        d_wh = torch.zeros(B, 2, H * W, device=g.device).scatter_add_(2, idx, g_wh).view(B, 2, H, W)
        s = dets[..., 4]
        flat = dets[..., 5].long() * (H * W) + inds
        d_hm = torch.zeros(B, C * H * W, device=g.device).scatter_add_(1, flat, g[..., 4] * s * (1 - s)).view(B, C, H, W)
        return d_hm, d_wh, d_off, None
