"""Host-side mirror of the reference's ResNet18 patch encoder (models.py:13-77: torchvision ``BasicBlock`` stages,
``class_classifier`` head; dispatch ``build_model`` models.py:201-204 replaces the head by ``nn.Identity`` and reads the
512-d average-pooled feature).  Same constructor, attribute / parameter names and initialisation order as the
reference, so its checkpoints (and torchvision's ``resnet18-5c106cde.pth``) load with ``load_state_dict`` and a seeded
construction gives identical weights.

The forward runs in libacmil_b200.so: every convolution is ``acmil_im2col`` + ``acmil_gemm_nt`` (3xTF32 on tcgen05,
fp32-faithful) on NHWC activations, with the eval-mode BatchNorm folded into the GEMM (scale into the weights, shift
as the epilogue bias) and the residual add + ReLU of ``BasicBlock.forward`` in the same epilogue.  Inference only
(``model.eval()``, like Step2_feature_extract.py), CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn
from torchvision.models.resnet import BasicBlock

from . import _lib as L
from .transmil import SplitImage, _need_cuda, _no_grad_path, _ptr, _stream, gemm_mode, gemm_nt


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """[Cout, Cin, kh, kw] conv + eval BatchNorm -> ([Cout, k_pad] weights in (ky, kx, c) column order, [Cout] bias)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(conv.out_channels, -1)
    bias = bn.bias - bn.running_mean * scale
    if conv.bias is not None:
        bias = bias + conv.bias * scale
    k = w.shape[1]
    k_pad = (k + 3) // 4 * 4
    if k_pad != k:
        w = torch.nn.functional.pad(w, (0, k_pad - k))
    return w.contiguous(), bias.contiguous()


class ResNet(nn.Module):
    """models.py:13-77 -- ``forward(x[B, 3, H, W]) -> class_classifier(avgpool features [B, 512 * expansion])``."""

    CHUNK = 64      # images per pass: keeps the im2col matrices of the wide early layers inside a few hundred MB

    def __init__(self, block, layers, classes=100):
        self.inplanes = 64
        super().__init__()
        if block is not BasicBlock:
            raise NotImplementedError("acmil_b200.resnet: only the BasicBlock networks (ResNet18/34) are built")
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.class_classifier = nn.Linear(512 * block.expansion, classes)
        self.pecent = 1 / 3
        for m in self.modules():      # models.py:30-35
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._packed = None
        self._packed_key = None

    def _make_layer(self, block, planes, blocks, stride=1):      # models.py:37-52
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ folded weights, rebuilt when a tensor changes
    def _weights(self):
        tensors = list(self.parameters()) + list(self.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors) + (gemm_mode(),)
        if self._packed is None or key != self._packed_key:
            with torch.no_grad():
                pk = {"conv1": _fold(self.conv1, self.bn1)}
                for li in range(1, 5):
                    for bi, blk in enumerate(getattr(self, f"layer{li}")):
                        pk[f"l{li}.{bi}.1"] = _fold(blk.conv1, blk.bn1)
                        pk[f"l{li}.{bi}.2"] = _fold(blk.conv2, blk.bn2)
                        if blk.downsample is not None:
                            pk[f"l{li}.{bi}.d"] = _fold(blk.downsample[0], blk.downsample[1])
                # the folded weights are the B operands of every product: fp16 hi / lo images for the fp16-split kernel
                split = gemm_mode() == 2 and pk["conv1"][0].is_cuda
                pk = {name: (w, b, SplitImage(w) if split else None) for name, (w, b) in pk.items()}
            self._packed, self._packed_key = pk, key
        return self._packed

    @staticmethod
    def _conv(x, dims, conv: nn.Conv2d, wb, *, relu, addend=None, nchw=False):
        """x: NHWC rows [B*H*W, C] (or the NCHW image batch) -> NHWC rows of the convolution output."""
        B, H, W, Cin = dims
        kh, kw = conv.kernel_size
        s, p = conv.stride[0], conv.padding[0]
        Ho, Wo = (H + 2 * p - kh) // s + 1, (W + 2 * p - kw) // s + 1
        w, bias, img = wb
        if kh == 1 and kw == 1 and s == 1 and p == 0 and not nchw:
            col = x
        else:
            col = torch.empty(B * Ho * Wo, w.shape[1], device=x.device, dtype=torch.float32)
            L.check(L.load().acmil_im2col(_ptr(x), _ptr(col), B, H, W, Cin, kh, kw, s, p, w.shape[1], int(nchw),
                                          _stream(x.device)))
        out = gemm_nt(col, w, bias=bias, addend=addend, relu=relu, b_split=img)
        return out, (B, Ho, Wo, conv.out_channels)

    def _features(self, x):
        pk = self._weights()
        B, _, H, W = x.shape
        dev = x.device
        lib = L.load()
        h, d = self._conv(x, (B, H, W, 3), self.conv1, pk["conv1"], relu=True, nchw=True)      # conv1 + bn1 + relu
        mp = self.maxpool
        k, s, p = mp.kernel_size, mp.stride, mp.padding
        Ho, Wo = (d[1] + 2 * p - k) // s + 1, (d[2] + 2 * p - k) // s + 1
        y = torch.empty(B * Ho * Wo, d[3], device=dev, dtype=torch.float32)
        L.check(lib.acmil_maxpool_nhwc(_ptr(h), _ptr(y), B, d[1], d[2], d[3], k, s, p, _stream(dev)))
        h, d = y, (B, Ho, Wo, d[3])
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):      # BasicBlock.forward (torchvision resnet.py)
                identity = h
                if blk.downsample is not None:
                    identity, _ = self._conv(h, d, blk.downsample[0], pk[f"l{li}.{bi}.d"], relu=False)
                o, d1 = self._conv(h, d, blk.conv1, pk[f"l{li}.{bi}.1"], relu=True)
                h, d = self._conv(o, d1, blk.conv2, pk[f"l{li}.{bi}.2"], relu=True, addend=identity)
        feat = torch.empty(B, d[3], device=dev, dtype=torch.float32)
        L.check(lib.acmil_avgpool_nhwc(_ptr(h), _ptr(feat), B, d[1] * d[2], d[3], _stream(dev)))
        return feat

    def forward(self, x):
        _need_cuda(x, "ResNet")
        _no_grad_path("ResNet", *self.parameters())
        if self.training:
            raise NotImplementedError("acmil_b200.resnet: inference only (BatchNorm uses its running statistics); call "
                                      "model.eval() like Step2_feature_extract.py does")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"ResNet: expected images [B, 3, H, W], got {tuple(x.shape)}")
        x = x.to(torch.float32).contiguous()
        feats = torch.cat([self._features(x[i:i + self.CHUNK]) for i in range(0, x.shape[0], self.CHUNK)], dim=0)
        cc = self.class_classifier
        if isinstance(cc, nn.Linear):
            return gemm_nt(feats, cc.weight.detach(), bias=cc.bias.detach() if cc.bias is not None else None)
        return cc(feats)


def resnet18(pretrained=True, **kwargs):
    """models.py:80-88.  ``pretrained`` downloads torchvision's ImageNet weights exactly like the reference does."""
    model = ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)
    if pretrained:
        from torch.utils import model_zoo
        model.load_state_dict(model_zoo.load_url('https://download.pytorch.org/models/resnet18-5c106cde.pth'), strict=False)
    return model
