"""Golden vectors for CLAM_SB / CLAM_MB / IBMIL, produced by running the REFERENCE itself (/root/reference,
architecture/clam.py, architecture/ibmil.py).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_consumers.py

architecture/clam.py imports utils.utils, which needs h5py / wandb (absent here): both are stubbed in sys.modules for the
import only -- no function of theirs is on the path.  The reference initialises every Linear bias with zero
(initialize_weights); the biases are re-drawn N(0, 0.1) after construction so that the vectors exercise them (stored like
every other weight).  Instance-eval cases store the reference's instance loss for label 1 (dropout=False so that the
training-mode call is deterministic).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, OUT)
for missing in ("h5py", "wandb"):
    sys.modules.setdefault(missing, types.ModuleType(missing))

from architecture.clam import CLAM_MB, CLAM_SB  # noqa: E402
from architecture.ibmil import IBMIL  # noqa: E402
from make_golden import Struct, make_x, save, sd_np, sha  # noqa: E402

torch.set_num_threads(8)


def meta(x, seed, scale):
    return dict(meta_x_seed=seed, meta_x_shape=np.array(x.shape), meta_x_sha=sha(x), meta_x_scale=scale)


def redraw_biases(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in m.named_parameters():
            if k.endswith("bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)


def clam_case(name, cls, model_seed, x_seed, n, d_feat, d_inner, n_class, gate, dropout, label):
    torch.manual_seed(model_seed)
    m = cls(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), gate=gate, dropout=dropout)
    redraw_biases(m, model_seed + 1)
    x = make_x(x_seed, (1, n, d_feat))
    out = dict(sd_np(m), **meta(x, x_seed, 1.0))
    out["meta_conf"] = np.array([d_feat, d_inner, n_class, int(gate), int(dropout), int(cls is CLAM_MB), label])
    m.eval()
    with torch.no_grad():
        out["eval_logits"] = m(x).numpy()
        out["eval_A_raw"] = m(x, attention_only=True).numpy()
        lab = torch.tensor([label])
        logits, inst = m(x, lab, instance_eval=True)
        out["eval_inst_loss"] = np.array(float(inst))
        np.testing.assert_array_equal(logits.numpy(), out["eval_logits"])
    save(name, **out)


def ibmil_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class, n_conf=0, merge="cat"):
    c_path = None
    tmp = None
    if n_conf:
        tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
        np.save(tmp.name, np.random.default_rng(model_seed).standard_normal((n_conf, d_inner)).astype(np.float32))
        c_path = [tmp.name]
    torch.manual_seed(model_seed)
    m = IBMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, c_path=c_path, c_learn=False), confounder_merge=merge)
    x = make_x(x_seed, (1, n, d_feat))
    out = dict(sd_np(m), **meta(x, x_seed, 1.0))
    out["meta_conf"] = np.array([d_feat, d_inner, n_class, n_conf])
    out["meta_merge"] = np.array(merge)
    m.eval()
    with torch.no_grad():
        y, M, A = m(x)
    out.update(out_Y=y.numpy(), out_M=M.numpy(), out_A=A.numpy())
    save(name, **out)
    if tmp is not None:
        os.unlink(tmp.name)


if __name__ == "__main__":
    clam_case("clam_sb_n1200", CLAM_SB, 71, 72, 1200, 384, 128, 2, True, True, 1)
    clam_case("clam_sb_nogate_d512_n800", CLAM_SB, 73, 74, 800, 512, 256, 2, False, False, 0)
    clam_case("clam_mb_c3_n1000", CLAM_MB, 75, 76, 1000, 384, 128, 3, True, True, 2)
    ibmil_case("ibmil_n900", 77, 78, 900, 384, 128, 2)
    ibmil_case("ibmil_conf_n700", 79, 80, 700, 384, 128, 2, n_conf=6, merge="cat")
    ibmil_case("ibmil_sub_n500", 81, 82, 500, 384, 128, 2, n_conf=4, merge="sub")
