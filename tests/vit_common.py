"""Shared by the ViT tests: the seeded weights of tests/golden/make_golden_vit.py (regenerated, not stored)."""
import numpy as np


def timm_shapes(img, patch, dim, depth, mlp):
    t = (img // patch) ** 2 + 1
    s = {"cls_token": (1, 1, dim), "pos_embed": (1, t, dim), "patch_embed.proj.weight": (dim, 3, patch, patch),
         "patch_embed.proj.bias": (dim,)}
    for i in range(depth):
        b = f"blocks.{i}."
        s.update({b + "norm1.weight": (dim,), b + "norm1.bias": (dim,), b + "attn.qkv.weight": (3 * dim, dim),
                  b + "attn.qkv.bias": (3 * dim,), b + "attn.proj.weight": (dim, dim), b + "attn.proj.bias": (dim,),
                  b + "norm2.weight": (dim,), b + "norm2.bias": (dim,), b + "mlp.fc1.weight": (mlp, dim),
                  b + "mlp.fc1.bias": (mlp,), b + "mlp.fc2.weight": (dim, mlp), b + "mlp.fc2.bias": (dim,)})
    s.update({"norm.weight": (dim,), "norm.bias": (dim,)})
    return s


def seeded_weights(shapes, seed):
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in shapes.items():
        v = (rng.standard_normal(shp) * 0.05).astype(np.float32)
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            v = v + 1
        out[k] = v
    return out
