"""Golden vectors for ACMIL_MHA / MHA, produced by running the REFERENCE itself (/root/reference,
architecture/transformer.py:50-182).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mha.py

The query tokens are initialised with std 1e-6 by the reference (transformer.py:59), which makes every attention
uniform; so that the vectors exercise the score path, q is re-drawn with std 1 after construction (stored in 'w::q' like
every other weight).  Train-mode cases: masking on (self.training), Dropout modules switched to eval (their noise is not
part of the path), and the uniform draws of transformer.py:168 are captured and stored ('rand_{i}').
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, OUT)

from architecture.transformer import ACMIL_MHA, MHA  # noqa: E402
from make_golden import Struct, make_x, save, sd_np, sha  # noqa: E402

torch.set_num_threads(8)


def meta(x, seed, scale):
    return dict(meta_x_seed=seed, meta_x_shape=np.array(x.shape), meta_x_sha=sha(x), meta_x_scale=scale)


def acmil_mha_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class, n_token, n_masked, drop, x_scale=1.0):
    torch.manual_seed(model_seed)
    m = ACMIL_MHA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), n_token=n_token, n_masked_patch=n_masked,
                  mask_drop=drop)
    with torch.no_grad():
        m.q.normal_(0, 1.0)
    x = make_x(x_seed, (1, n, d_feat), x_scale)
    out = dict(sd_np(m), **meta(x, x_seed, x_scale))
    m.eval()
    with torch.no_grad():
        sub, slide, attns = m(x)
    out.update(eval_sub=sub.numpy(), eval_slide=slide.numpy(), eval_attns=attns.numpy())
    if n_masked > 0:
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.eval()
        draws = []
        real_rand = torch.rand

        def spy(*a, **k):
            r = real_rand(*a, **k)
            draws.append(r.clone())
            return r

        torch.rand = spy
        try:
            torch.manual_seed(7)
            with torch.no_grad():
                sub, slide, attns = m(x)
        finally:
            torch.rand = real_rand
        assert len(draws) == n_token
        out.update(train_sub=sub.numpy(), train_slide=slide.numpy(), train_attns=attns.numpy())
        for i, r in enumerate(draws):
            out[f"rand_{i}"] = r.numpy()
    save(name, **out, meta_conf=np.array([d_feat, d_inner, n_class, n_token, n_masked]), meta_mask_drop=np.float64(drop))


def mha_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class):
    torch.manual_seed(model_seed)
    m = MHA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)).eval()
    with torch.no_grad():
        m.q.normal_(0, 1.0)
    x = make_x(x_seed, (1, n, d_feat))
    with torch.no_grad():
        y = m(x)
    save(name, **sd_np(m), **meta(x, x_seed, 1.0), out=y.numpy(), meta_conf=np.array([d_feat, d_inner, n_class]))


if __name__ == "__main__":
    acmil_mha_case("acmilmha_k5_n1500", 61, 901, 1500, 384, 128, 2, 5, 10, 0.6)
    acmil_mha_case("acmilmha_k3_d256_n700", 62, 902, 700, 512, 256, 3, 3, 10, 0.6, x_scale=0.5)
    acmil_mha_case("acmilmha_k1_n7", 63, 903, 7, 384, 128, 2, 1, 10, 0.6)        # bag smaller than n_masked_patch
    mha_case("mha_n900", 64, 904, 900, 384, 128, 2)
