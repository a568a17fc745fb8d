"""Host-side mirror of the reference's patch encoder (models.py): ``vit_small`` (models.py:138-149, a timm 0.9.2
``VisionTransformer(img_size=224, patch_size=16, embed_dim=384, num_heads=6, num_classes=0)``), ``CustomModel``
(models.py:166-179) and ``build_model`` (models.py:191-215) for the ViT-S/16 backbones.  Parameter names follow timm's
(``cls_token, pos_embed, patch_embed.proj.*, blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*, norm.*``), so
the Lunit DINO checkpoint the reference downloads (models.py:113-123) loads with ``load_state_dict``.  The forward runs
in libacmil_b200.so (``acmil_vit_fwd``, include/acmil_transmil.h); inference only, CUDA tensors only.

timm is not installed in the build container, so parity with timm itself is UNPINNED; the same math is cross-checked
against torchvision's independent ``VisionTransformer`` (tests/golden/make_golden_vit.py).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .transmil import SplitCache, _need_cuda, _no_grad_path, _ptr, _stream, gemm_mode


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class VisionTransformer(nn.Module):
    """timm-0.9.2-shaped ViT: ``forward(images[B, 3, img, img]) -> features [B, embed_dim]`` (``num_classes=0``,
    ``global_pool='token'``: final LayerNorm, class token)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=384, depth=12, num_heads=6,
                 mlp_ratio=4., qkv_bias=True):
        super().__init__()
        if not qkv_bias:
            raise NotImplementedError("acmil_b200 ViT: qkv_bias=False is not supported")
        self.num_classes, self.embed_dim = num_classes, embed_dim
        self.num_features = embed_dim
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * .02)
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.precise = True
        self._ws = None
        self._split = SplitCache()
        self._init_weights()

    def _init_weights(self):      # timm: trunc_normal_(std=.02) for pos_embed / Linear weights, zero biases, cls ~ N(0, 1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _forward(self, x, head_w=None, head_b=None):
        _need_cuda(x, "VisionTransformer")
        _no_grad_path("VisionTransformer", x, *self.parameters())
        pe = self.patch_embed
        if x.dim() != 4 or x.shape[1] != pe.proj.in_channels or x.shape[2] != pe.img_size or x.shape[3] != pe.img_size:
            raise AssertionError(f"Input image size {tuple(x.shape)} doesn't match model ({pe.proj.in_channels}, {pe.img_size}, {pe.img_size})")
        x = x.contiguous()
        B = x.shape[0]
        blk0 = self.blocks[0]
        n_class = head_w.shape[0] if head_w is not None else 0
        shape = L.VitShape(B, pe.img_size, pe.patch_size, pe.proj.in_channels, self.embed_dim, len(self.blocks),
                           blk0.attn.num_heads, blk0.mlp.fc1.out_features, n_class, gemm_mode() if self.precise else 0,
                           float(self.norm.eps))
        lib = L.load()
        nbytes = C.c_size_t(0)
        L.check(lib.acmil_vit_workspace_bytes(C.byref(shape), C.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != x.device:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        keep = []

        def p(t):
            t = t.detach().contiguous()
            keep.append(t)
            return _ptr(t)

        split = self._split.get if self.precise else (lambda *a, **k: None)      # weight images for the fp16-split products

        def sp(name, t, view=None):
            img = split(name, t, view)
            keep.append(img)
            return img.ptr if img is not None else None

        blocks = (L.VitBlockWeights * len(self.blocks))()
        for i, b in enumerate(self.blocks):
            blocks[i] = L.VitBlockWeights(p(b.norm1.weight), p(b.norm1.bias), p(b.attn.qkv.weight), p(b.attn.qkv.bias),
                                          p(b.attn.proj.weight), p(b.attn.proj.bias), p(b.norm2.weight), p(b.norm2.bias),
                                          p(b.mlp.fc1.weight), p(b.mlp.fc1.bias), p(b.mlp.fc2.weight), p(b.mlp.fc2.bias),
                                          sp(f"{i}.qkv", b.attn.qkv.weight), sp(f"{i}.proj", b.attn.proj.weight),
                                          sp(f"{i}.fc1", b.mlp.fc1.weight), sp(f"{i}.fc2", b.mlp.fc2.weight))
        w = L.VitWeights(p(self.cls_token), p(self.pos_embed), p(pe.proj.weight), p(pe.proj.bias), p(self.norm.weight),
                         p(self.norm.bias), p(head_w) if head_w is not None else None,
                         p(head_b) if head_b is not None else None, blocks,
                         sp("patch", pe.proj.weight, lambda t: t.reshape(t.shape[0], -1)))
        feats = torch.empty(B, self.embed_dim, device=x.device, dtype=torch.float32)
        logits = torch.empty(B, n_class, device=x.device, dtype=torch.float32) if n_class else None
        L.check(lib.acmil_vit_fwd(C.byref(shape), C.byref(w), _ptr(x), _ptr(feats), _ptr(logits), _ptr(self._ws),
                                  self._ws.numel(), _stream(x.device)))
        return feats, logits

    def forward(self, x):
        if isinstance(self.head, nn.Linear):
            return self._forward(x, self.head.weight, self.head.bias)[1]
        return self._forward(x)[0]


def vit_small(pretrained, progress, key, **kwargs):
    """models.py:138-149."""
    patch_size = kwargs.get("patch_size", 16)
    model = VisionTransformer(img_size=224, patch_size=patch_size, embed_dim=384, num_heads=6, num_classes=0)
    if pretrained:
        from .vit_urls import get_pretrained_url
        verbose = model.load_state_dict(torch.hub.load_state_dict_from_url(get_pretrained_url(key), progress=progress))
        print(verbose)
    return model


class CustomModel(nn.Module):
    """models.py:166-179 -- ``forward(image, return_feature=False)``."""

    def __init__(self, cfg, encoder):
        super().__init__()
        self.encoder = encoder
        self.head = nn.Linear(encoder.embed_dim, cfg.n_class)

    def forward(self, image, return_feature=False):
        if isinstance(self.encoder, VisionTransformer) and isinstance(self.encoder.head, nn.Identity):
            _no_grad_path("CustomModel", self.head.weight)
            image_features, logits = self.encoder._forward(image, self.head.weight, self.head.bias)
        else:
            image_features = self.encoder(image)
            logits = self.head(image_features)
        if return_feature:
            return logits, image_features
        return logits


def build_model(cfg):
    """models.py:191-215, ViT-S/16 backbones (the Lunit DINO weights are downloaded exactly like the reference does; without
    network access construct ``CustomModel(cfg, vit_small(False, False, None))`` and load a local checkpoint)."""
    if cfg.backbone == 'ViT-S/16' and cfg.pretrain in ('medical_ssl',):
        encoder = vit_small(pretrained=True, progress=False, key="DINO_p16", patch_size=16)
    elif cfg.pretrain == 'tailored_sl':
        encoder = vit_small(pretrained=True, progress=False, key="DINO_p16", patch_size=16)
    elif cfg.pretrain == 'natural_supervised' and cfg.backbone == 'Resnet18':      # models.py:201-204
        from .resnet import resnet18
        encoder = resnet18()
        encoder.class_classifier = nn.Identity()
        encoder.embed_dim = encoder.inplanes
    else:
        raise NotImplementedError(f"acmil_b200.build_model: backbone {cfg.backbone!r} / pretrain {cfg.pretrain!r} is not built "
                                  "(only the ViT-S/16 and ResNet18 encoders of SURVEY section 8a rows a12-a13)")
    return CustomModel(cfg, encoder)
