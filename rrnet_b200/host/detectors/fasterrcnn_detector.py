"""detectors/fasterrcnn_detector.py:6-18 + backbones/resnet.py:17-53 -- the re-regression head.

Same parameter names as the reference (state_dict keys head_detector.top_layer.{conv1,bn1,conv2,bn2,
conv3,bn3}.*, head_detector.regressor.*).  In eval mode without autograd the whole head is one fused
kernel (rr_head_forward, BatchNorms folded once per weight version); in training mode it is the plain
PyTorch graph, because SyncBatchNorm batch statistics couple all RoIs of all ranks
(operators/rrnet_operator.py:27) and stay with PyTorch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from rrnet_b200 import ops


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(Bottleneck, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out += residual
        return self.relu(out)


class FasterRCNNDetector(nn.Module):
    def __init__(self):
        super(FasterRCNNDetector, self).__init__()
        self.top_layer = Bottleneck(inplanes=256, planes=64)
        self.regressor = nn.Conv2d(256, 4, kernel_size=1)
        self._folded = None
        self._folded_key = None

    def folded(self):
        """BN-folded parameter block for the fused kernel, rebuilt when any parameter/buffer changes."""
        ts = [t for _, t in sorted(self.state_dict().items()) if t.dtype.is_floating_point]
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if self._folded is None or key != self._folded_key:
            self._folded = ops.head_fold(ops.head_params_from_module(self))
            self._folded_key = key
        return self._folded

    def forward(self, feat):
        fused = (not self.training) and feat.is_cuda and not (torch.is_grad_enabled() and (
            feat.requires_grad or any(p.requires_grad for p in self.parameters())))
        if fused:
            return ops.head_forward(feat, self.folded())
        if not self.training and not feat.is_cuda:
            raise ops.RRNetB200Error("FasterRCNNDetector: eval-mode forward needs CUDA tensors (no CPU path)")
        feat = self.top_layer(feat)
        feat = F.adaptive_avg_pool2d(feat, 1)
        reg = self.regressor(feat)
        reg = reg.view(reg.size(0), reg.size(1))
        return reg
