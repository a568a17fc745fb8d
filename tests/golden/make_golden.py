"""Generate golden vectors by running the REFERENCE itself (read-only, /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Inputs are stored by (seed, shape) + a sha256 of their
bytes; weights are stored explicitly when small (ACMIL heads: the reference's own
initialiser under torch.manual_seed) and by numpy seed when large (attmil 1024->512
front layers, loaded INTO the reference module with load_state_dict).

Nothing here is product code; nothing here is imported at test time.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from architecture.transformer import ACMIL_GA, ABMIL, Attention_Gated as TAttentionGated  # noqa: E402
from architecture import Attention as RefAttention  # noqa: E402
from architecture import attmil as ref_attmil  # noqa: E402


class Struct:  # utils/utils.py:246-248 (utils.utils itself needs h5py/wandb, absent here)
    def __init__(self, **entries):
        self.__dict__.update(entries)


torch.set_num_threads(8)


def make_x(seed, shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def sd_np(m):
    return {"w::" + k: v.detach().numpy().copy() for k, v in m.state_dict().items()}


def save(name, **kw):
    np.savez(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items() if not k.startswith("w::")})


def div_loss(attn):  # Step3_WSI_classification_ACMIL.py:208-214 executed verbatim by torch
    k = attn.shape[1]
    a = torch.softmax(attn, dim=-1)
    tot = torch.tensor(0.0)
    for i in range(k):
        for j in range(i + 1, k):
            tot = tot + torch.cosine_similarity(a[:, i], a[:, j], dim=-1).mean() / (k * (k - 1) / 2)
    return tot


def ent_loss(attn):  # Step3_WSI_classification_ACMIL.py:259
    import torch.nn.functional as F
    return torch.sum(F.softmax(attn, dim=-1) * F.log_softmax(attn, dim=-1)) / attn.shape[1]


def acmil_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class, n_token, n_masked, drop,
               train_seed=None, x_scale=1.0, fp16_round=False):
    conf = Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=n_token)
    torch.manual_seed(model_seed)
    m = ACMIL_GA(conf, n_token=n_token, n_masked_patch=n_masked, mask_drop=drop)
    x = make_x(x_seed, (1, n, d_feat), x_scale)
    if fp16_round:
        x = x.half().float()
    out = dict(sd_np(m))
    out.update(meta_model_seed=model_seed, meta_x_seed=x_seed, meta_x_shape=np.array([1, n, d_feat]),
               meta_x_scale=x_scale, meta_x_fp16=int(fp16_round), meta_x_sha=sha(x),
               meta_conf=np.array([d_feat, d_inner, n_class, n_token, n_masked]), meta_mask_drop=drop)
    m.eval()
    with torch.no_grad():
        sub, slide, a = m(x)
        feat = m.forward_feature(x)
    out.update(eval_sub=sub.numpy(), eval_slide=slide.numpy(), eval_A=a.numpy(), eval_feat=feat.numpy(),
               eval_div=div_loss(a).numpy(), eval_ent=ent_loss(a).numpy())
    if train_seed is not None and n_masked > 0:
        m.train()
        nm = min(n_masked, n)
        torch.manual_seed(train_seed)
        rand = torch.rand(n_token, nm)          # what transformer.py:316 will draw next
        torch.manual_seed(train_seed)
        with torch.no_grad():
            sub, slide, a = m(x)
        masked = (a[0] == -1e9).nonzero()
        keep = int(nm * drop)
        mi = masked[:, 1].reshape(n_token, keep) if keep > 0 else np.zeros((n_token, 0), np.int64)
        torch.manual_seed(train_seed)
        with torch.no_grad():
            feat_m = m.forward_feature(x, use_attention_mask=True)
        out.update(train_seed=train_seed, train_rand=rand.numpy(), train_sub=sub.numpy(),
                   train_slide=slide.numpy(), train_A=a.numpy(),
                   train_masked_sorted=np.asarray(mi), train_feat=feat_m.numpy(),
                   train_div=div_loss(a).numpy())
    save(name, **out)


def abmil_case(name, model_seed, x_seed, n, d_feat, d_inner, n_class):
    conf = Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=1)
    torch.manual_seed(model_seed)
    m = ABMIL(conf).eval()
    x = make_x(x_seed, (1, n, d_feat))
    with torch.no_grad():
        y = m(x)
    save(name, **sd_np(m), meta_model_seed=model_seed, meta_x_seed=x_seed,
         meta_x_shape=np.array([1, n, d_feat]), meta_x_sha=sha(x),
         meta_conf=np.array([d_feat, d_inner, n_class, 1, 0]), eval_out=y.numpy())


def attention_py_case(name, model_seed, x_seed, n, L, D, K, num_cls):
    torch.manual_seed(model_seed)
    g = RefAttention.Attention_Gated(L, D, K).eval()
    torch.manual_seed(model_seed + 1)
    awc = RefAttention.Attention_with_Classifier(L, D, K, num_cls).eval()
    torch.manual_seed(model_seed + 2)
    tg = TAttentionGated(L, D, K).eval()
    x = make_x(x_seed, (n, L))
    with torch.no_grad():
        a_norm = g(x)
        a_raw = g(x, isNorm=False)
        pred = awc(x)
        t_raw = tg(x)
    out = {"w::gate::" + k: v.numpy().copy() for k, v in g.state_dict().items()}
    out.update({"w::awc::" + k: v.numpy().copy() for k, v in awc.state_dict().items()})
    out.update({"w::tgate::" + k: v.numpy().copy() for k, v in tg.state_dict().items()})
    save(name, **out, meta_model_seed=model_seed, meta_x_seed=x_seed, meta_x_shape=np.array([n, L]),
         meta_x_sha=sha(x), meta_conf=np.array([L, D, K, num_cls]),
         gate_norm=a_norm.numpy(), gate_raw=a_raw.numpy(), awc_pred=pred.numpy(), tgate_raw=t_raw.numpy())


def np_state(module, seed, scale=0.05):
    """Seeded numpy weights for big modules (kept out of the fixture file)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in module.state_dict().items():
        sd[k] = torch.from_numpy((rng.standard_normal(tuple(v.shape)) * scale).astype(np.float32))
    module.load_state_dict(sd)
    return module


def attmil_case(name, w_seed, x_seed, n):
    out = dict(meta_w_seed=w_seed, meta_x_seed=x_seed, meta_x_shape=np.array([1, n, 1024]), meta_w_scale=0.05)
    x = make_x(x_seed, (1, n, 1024))
    out["meta_x_sha"] = sha(x)
    for act in ("relu", "gelu", "tanh"):
        for bias in (False, True):
            m = np_state(ref_attmil.AttentionGated(act=act, bias=bias), w_seed).eval()
            with torch.no_grad():
                out[f"ag_{act}_{int(bias)}"] = m(x).numpy()
    for act in ("relu", "gelu"):
        m = np_state(ref_attmil.DAttention(n_classes=3, dropout=False, act=act), w_seed + 1).eval()
        with torch.no_grad():
            y, a = m(x, return_attn=True)
            _, a_ori = m(x, return_attn=True, no_norm=True)
            y2 = m(x)
        assert torch.equal(y, y2)
        out[f"da_{act}_y"] = y.numpy()
        out[f"da_{act}_A"] = a.numpy()
        out[f"da_{act}_Aori"] = a_ori.numpy()
    save(name, **out)


if __name__ == "__main__":  # pragma: no cover
    # KAT1/2/3 of SURVEY.md section 4 (C1 shapes: N=1024, D_feat=384, D_inner=128)
    acmil_case("acmil_ga_k1_n1024", 1, 1234, 1024, 384, 128, 2, 1, 0, 0.0)
    acmil_case("acmil_ga_k5_n1024", 1, 1234, 1024, 384, 128, 2, 5, 10, 0.6, train_seed=7)
    # ragged / tiny bags: N < n_masked_patch (min() at transformer.py:314), N = 1
    acmil_case("acmil_ga_k5_n7", 3, 99, 7, 384, 128, 2, 5, 10, 0.6, train_seed=11)
    acmil_case("acmil_ga_k5_n1", 3, 98, 1, 384, 128, 2, 5, 10, 0.6, train_seed=12)
    acmil_case("acmil_ga_k5_n333", 4, 97, 333, 384, 128, 2, 5, 10, 0.6, train_seed=13)
    # natural_supervised table entry (512 -> 256), 3 classes, 3 branches, wide-range fp16-rounded input
    acmil_case("acmil_ga_k3_d512_n2000", 5, 96, 2000, 512, 256, 3, 3, 10, 0.6, train_seed=14,
               x_scale=3.0, fp16_round=True)
    # 8 branches, larger mask
    acmil_case("acmil_ga_k8_n4099", 6, 95, 4099, 384, 128, 2, 8, 20, 0.5, train_seed=15)
    abmil_case("abmil_n777", 2, 4321, 777, 384, 128, 2)
    abmil_case("abmil_d512_n1500", 8, 4322, 1500, 512, 256, 3)
    attention_py_case("attention_py_n900", 21, 555, 900, 512, 128, 1, 2)
    attention_py_case("attention_py_k4_n640", 22, 556, 640, 256, 128, 4, 3)
    attmil_case("attmil_n600", 31, 777, 600)
