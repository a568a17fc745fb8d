"""Cross-check vectors for the ViT patch encoder.  timm 0.9.2 (what models.py:140 instantiates) is absent from the build
container, so these come from torchvision's independent VisionTransformer -- same architecture and math -- with its
parameters renamed to timm's layout.  Weights are drawn by numpy seed (tests regenerate them; only outputs are stored).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_vit.py
"""
import os
import sys

import numpy as np
import torch
from torchvision.models.vision_transformer import VisionTransformer as TVViT

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
from make_golden import make_x, save, sha  # noqa: E402

torch.set_num_threads(8)


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
    """Same generator as tests/conftest.np_seeded_state: N(0, 0.05) except LayerNorm weights ~ 1 + N(0, 0.05)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in shapes.items():
        v = (rng.standard_normal(shp) * 0.05).astype(np.float32)
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            v = v + 1
        out[k] = v
    return out


def to_torchvision(w, depth):
    m = {"class_token": w["cls_token"], "encoder.pos_embedding": w["pos_embed"], "conv_proj.weight": w["patch_embed.proj.weight"],
         "conv_proj.bias": w["patch_embed.proj.bias"], "encoder.ln.weight": w["norm.weight"], "encoder.ln.bias": w["norm.bias"]}
    for i in range(depth):
        a, b = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        m.update({b + "ln_1.weight": w[a + "norm1.weight"], b + "ln_1.bias": w[a + "norm1.bias"],
                  b + "self_attention.in_proj_weight": w[a + "attn.qkv.weight"], b + "self_attention.in_proj_bias": w[a + "attn.qkv.bias"],
                  b + "self_attention.out_proj.weight": w[a + "attn.proj.weight"], b + "self_attention.out_proj.bias": w[a + "attn.proj.bias"],
                  b + "ln_2.weight": w[a + "norm2.weight"], b + "ln_2.bias": w[a + "norm2.bias"],
                  b + "mlp.0.weight": w[a + "mlp.fc1.weight"], b + "mlp.0.bias": w[a + "mlp.fc1.bias"],
                  b + "mlp.3.weight": w[a + "mlp.fc2.weight"], b + "mlp.3.bias": w[a + "mlp.fc2.bias"]})
    return {k: torch.from_numpy(v) for k, v in m.items()}


def vit_case(name, w_seed, x_seed, batch, img, patch, dim, depth, heads, mlp):
    w = seeded_weights(timm_shapes(img, patch, dim, depth, mlp), w_seed)
    tv = TVViT(image_size=img, patch_size=patch, num_layers=depth, num_heads=heads, hidden_dim=dim, mlp_dim=mlp, num_classes=5)
    tv.heads = torch.nn.Identity()
    missing = tv.load_state_dict(to_torchvision(w, depth), strict=True)
    tv.eval()
    x = make_x(x_seed, (batch, 3, img, img))
    with torch.no_grad():
        y = tv(x)
    save(name, meta_x_seed=x_seed, meta_x_shape=np.array(x.shape), meta_x_sha=sha(x), meta_w_seed=w_seed,
         meta_cfg=np.array([img, patch, dim, depth, heads, mlp]), out=y.numpy())


if __name__ == "__main__":
    vit_case("vit_tiny_64", 61, 901, 3, 64, 16, 96, 2, 3, 384)
    vit_case("vit_small16_224", 62, 902, 2, 224, 16, 384, 12, 6, 1536)      # the ViT-S/16 of models.py:140
