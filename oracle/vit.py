"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ViT-S/16 patch encoder:
models.py:138-149 ``vit_small`` = timm 0.9.2 ``VisionTransformer(img_size=224, patch_size=16, embed_dim=384, num_heads=6,
num_classes=0)`` with timm's defaults (depth 12, mlp_ratio 4, qkv_bias, LayerNorm eps 1e-6, exact-erf GELU, class token +
learned position embedding, global_pool='token', final norm, head = Identity), and CustomModel.forward (models.py:174-179).

Parity status: UNPINNED at the timm boundary -- timm==0.9.2 (requirements.txt:72) is not installed in the build container and
its source is not under /root/reference, so the reference encoder cannot be run here.  The restatement follows timm's
published forward and is cross-checked against torchvision's independent ``VisionTransformer`` (same math, different code)
by tests/golden/make_golden_vit.py / tests/test_vit_oracle.py.  Weights: dict keyed by timm's parameter names.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf


def _f(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _ln(x, w, b, eps=1e-6):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def _gelu(x):
    return x.dtype.type(0.5) * x * (1 + erf(x / np.sqrt(2.0)).astype(x.dtype))


def vit_forward(p, images, *, num_heads, patch, dtype=np.float32):
    """timm VisionTransformer.forward_features + forward_head(pre_logits) -> [B, embed_dim]."""
    x = _f(images, dtype)
    B, C, S, _ = x.shape
    g = S // patch
    w = _f(p["patch_embed.proj.weight"], dtype)                     # PatchEmbed: Conv2d(k = stride = patch), flatten(2).T
    D = w.shape[0]
    cols = x.reshape(B, C, g, patch, g, patch).transpose(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * patch * patch)
    x = cols @ w.reshape(D, -1).T + _f(p["patch_embed.proj.bias"], dtype)
    cls = np.broadcast_to(_f(p["cls_token"], dtype), (B, 1, D))      # _pos_embed: cat the class token, then add pos_embed
    x = np.concatenate([cls, x], axis=1) + _f(p["pos_embed"], dtype)
    depth = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("blocks."))
    dh = D // num_heads
    for i in range(depth):                                           # Block: x += attn(norm1(x)); x += mlp(norm2(x))
        pre = f"blocks.{i}."
        y = _ln(x, _f(p[pre + "norm1.weight"], dtype), _f(p[pre + "norm1.bias"], dtype))
        qkv = y @ _f(p[pre + "attn.qkv.weight"], dtype).T + _f(p[pre + "attn.qkv.bias"], dtype)
        qkv = qkv.reshape(B, -1, 3, num_heads, dh).transpose(2, 0, 3, 1, 4)      # 3, B, heads, N, dh
        q, k, v = qkv[0], qkv[1], qkv[2]
        att = (q * dtype(dh ** -0.5)) @ np.swapaxes(k, -1, -2)
        att = np.exp(att - att.max(-1, keepdims=True))
        att = att / att.sum(-1, keepdims=True)
        y = (att @ v).transpose(0, 2, 1, 3).reshape(B, -1, D)
        x = x + y @ _f(p[pre + "attn.proj.weight"], dtype).T + _f(p[pre + "attn.proj.bias"], dtype)
        y = _ln(x, _f(p[pre + "norm2.weight"], dtype), _f(p[pre + "norm2.bias"], dtype))
        y = _gelu(y @ _f(p[pre + "mlp.fc1.weight"], dtype).T + _f(p[pre + "mlp.fc1.bias"], dtype))
        x = x + y @ _f(p[pre + "mlp.fc2.weight"], dtype).T + _f(p[pre + "mlp.fc2.bias"], dtype)
    x = _ln(x, _f(p["norm.weight"], dtype), _f(p["norm.bias"], dtype))
    return x[:, 0]                                                   # global_pool == 'token'


def custom_model_forward(p, images, *, num_heads, patch, dtype=np.float32):
    """CustomModel.forward(image, return_feature=True): models.py:174-179 -> (logits, features); encoder keys under 'encoder.'."""
    enc = {k[len("encoder."):]: v for k, v in p.items() if k.startswith("encoder.")}
    feat = vit_forward(enc, images, num_heads=num_heads, patch=patch, dtype=dtype)
    return feat @ _f(p["head.weight"], dtype).T + _f(p["head.bias"], dtype), feat
