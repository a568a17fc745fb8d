"""Golden vectors for the TransMIL / Nystrom path, produced by running the REFERENCE itself (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_transmil.py

Shims: `architecture/transMIL.py:5` imports the pip module nystrom-attention==0.0.12, which is
absent; the vendored `architecture/nystrom_attention.py` is the same algorithm and is registered under that name.
`transMIL.py:71` calls `.cuda()` on the class token; on this CPU-only container it is made a no-op.
Weights come from the reference's own initialisers under torch.manual_seed and are stored explicitly ('w::' keys).
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

import architecture.nystrom_attention as ref_nys  # noqa: E402

sys.modules["nystrom_attention"] = ref_nys
torch.Tensor.cuda = lambda self, *a, **k: self      # CPU container
import architecture.transMIL as ref_tm  # noqa: E402
from make_golden import Struct, make_x, save, sd_np, sha  # noqa: E402

torch.set_num_threads(8)


def meta(x, seed):
    return dict(meta_x_seed=seed, meta_x_shape=np.array(x.shape), meta_x_sha=sha(x))


def nystrom_case(name, model_seed, x_seed, b, n, dim, dim_head, heads, m, ks=33, residual=True, iters=6):
    torch.manual_seed(model_seed)
    mod = ref_nys.NystromAttention(dim=dim, dim_head=dim_head, heads=heads, num_landmarks=m, pinv_iterations=iters,
                                   residual=residual, residual_conv_kernel=ks, dropout=0.1).eval()
    x = make_x(x_seed, (b, n, dim))
    with torch.no_grad():
        y = mod(x)
    save(name, **sd_np(mod), **meta(x, x_seed), out=y.numpy(),
         meta_cfg=np.array([dim, dim_head, heads, m, ks, int(residual), iters]))


def translayer_case(name, model_seed, x_seed, n, dim):
    torch.manual_seed(model_seed)
    mod = ref_tm.TransLayer(dim=dim).eval()
    with torch.no_grad():      # a non-trivial LayerNorm affine
        mod.norm.weight.uniform_(0.5, 1.5)
        mod.norm.bias.uniform_(-0.3, 0.3)
    x = make_x(x_seed, (1, n, dim))
    with torch.no_grad():
        y = mod(x)
    save(name, **sd_np(mod), **meta(x, x_seed), out=y.numpy(), meta_cfg=np.array([dim]))


def ppeg_case(name, model_seed, x_seed, b, gh, gw, dim):
    torch.manual_seed(model_seed)
    mod = ref_tm.PPEG(dim=dim).eval()
    x = make_x(x_seed, (b, 1 + gh * gw, dim))
    with torch.no_grad():
        y = mod(x, gh, gw)
    save(name, **sd_np(mod), **meta(x, x_seed), out=y.numpy(), meta_cfg=np.array([dim, gh, gw]))


def transmil_case(name, model_seed, x_seed, b, n, d_feat, d_inner, n_class):
    torch.manual_seed(model_seed)
    mod = ref_tm.TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)).eval()
    x = make_x(x_seed, (b, n, d_feat))
    with torch.no_grad():
        y = mod(x)
    save(name, **sd_np(mod), **meta(x, x_seed), out=y.numpy(), meta_cfg=np.array([d_feat, d_inner, n_class]))


if __name__ == "__main__":
    nystrom_case("nystrom_pad_n77", 41, 801, 1, 77, 128, 32, 4, 16)               # front padding 3
    nystrom_case("nystrom_nopad_b2_n256", 42, 802, 2, 256, 64, 16, 8, 32)          # batch 2, n % m == 0, inner != dim
    nystrom_case("nystrom_noresid_n100", 43, 803, 1, 100, 96, 24, 4, 20, residual=False, iters=4)
    translayer_case("translayer_d128_n500", 44, 804, 500, 128)
    ppeg_case("ppeg_d64_9x7", 45, 805, 2, 9, 7, 64)
    transmil_case("transmil_d64_n300", 46, 806, 1, 300, 96, 64, 2)
    transmil_case("transmil_d128_n1000", 47, 807, 1, 1000, 384, 128, 3)
    transmil_case("transmil_d64_b2_n50", 48, 808, 2, 50, 32, 64, 2)


def transmil_seeded_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class):
    """Big cases: the weights are NOT stored (2.4 M parameters); the test rebuilds them with torch.manual_seed(model_seed)
    through acmil_b200's TransMIL, whose constructor draws in the reference's order (checked here through a digest of
    the reference module's state_dict).  Stored: logits, the seeds, sha256 of input and of the weights."""
    import hashlib
    torch.manual_seed(model_seed)
    mod = ref_tm.TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)).eval()
    h = hashlib.sha256()
    for k, v in mod.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    x = make_x(x_seed, (1, n, d_feat))
    with torch.no_grad():
        y = mod(x)
    save(name, **meta(x, x_seed), out=y.numpy(), meta_cfg=np.array([d_feat, d_inner, n_class]),
         meta_model_seed=model_seed, meta_w_sha=h.hexdigest())


if __name__ == "__main__":
    # TransMIL(512, 512, 2).eval() after torch.manual_seed(0), x = randn(1, 1000, 512), seed 1234
    transmil_seeded_case("seeded_transmil_kat4_n1000", 0, 1234, 1000, 512, 512, 2)
    # BASELINE.json configs[2] at its own size: N = 50 000, dim 512 -> 256 landmarks
    transmil_seeded_case("seeded_transmil_c3_n50000", 5, 6, 50000, 512, 512, 2)
