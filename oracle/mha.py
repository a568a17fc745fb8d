"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ACMIL_MHA / MHA forward
(architecture/transformer.py:50-236), op by op as the reference executes it (NOT the collapsed single-query algebra of
acmil_b200/mha.py, so that the two check each other).  Parity status: PINNED by tests/golden/make_golden_mha.py, which
runs the reference modules themselves; tests/test_mha_oracle.py checks this file against those vectors.

Nothing under acmil_b200/ imports this file.
"""
from __future__ import annotations

import numpy as np


def _lin(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def _layer_norm(x, w, b, eps=1e-6):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def _softmax(a):
    a = a - a.max(-1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(-1, keepdims=True)


def _heads(x, nh):      # [n, c] -> [nh, n, c / nh]            (transformer.py:138-141)
    n, c = x.shape
    return x.reshape(n, nh, c // nh).transpose(1, 0, 2)


def multi_head_attention(p, pre, q, h, nh=8, n_masked=0, mask_drop=0.0, rand=None, dtype=np.float32):
    """MutiHeadAttention.forward (transformer.py:148-182), dropout off.  q [1, E], h [N, E] (k = v = h).
    rand [nh, min(n_masked, N)]: the uniform draws of :168 (masking is applied iff it is given).
    -> (out1[0] [1, E], attn_out[0] [nh, 1, N], masked indices [nh, keep] or None)."""
    f = lambda a: np.asarray(a, dtype)      # noqa: E731
    qp = _heads(_lin(f(q), f(p[pre + "q_proj.weight"]), f(p[pre + "q_proj.bias"])), nh)      # [nh, 1, d]
    kp = _heads(_lin(f(h), f(p[pre + "k_proj.weight"]), f(p[pre + "k_proj.bias"])), nh)      # [nh, N, d]
    vp = _heads(_lin(f(h), f(p[pre + "v_proj.weight"]), f(p[pre + "v_proj.bias"])), nh)
    d = qp.shape[-1]
    attn = qp @ kp.transpose(0, 2, 1) / dtype(np.sqrt(d))                                    # [nh, 1, N]   (:158-160)
    masked = None
    if rand is not None and n_masked > 0:                                                      # :162-174
        n = attn.shape[-1]
        nm = min(n_masked, n)
        a2 = attn.reshape(nh, n)
        top = np.argsort(-a2, axis=-1, kind="stable")[:, :nm]                                 # torch.topk order
        keep = int(nm * mask_drop)
        rsel = np.argsort(np.asarray(rand)[:, :nm], axis=-1, kind="stable")[:, :keep]
        masked = np.take_along_axis(top, rsel, axis=-1)
        a2 = a2.copy()
        np.put_along_axis(a2, masked, dtype(-1e9), axis=-1)
        attn = a2.reshape(nh, 1, n)
    attn_out = attn
    out = _softmax(attn) @ vp                                                                  # [nh, 1, d]   (:176-179)
    out = out.transpose(1, 0, 2).reshape(1, -1)                                                # recombine heads
    out = _lin(out, f(p[pre + "out_proj.weight"]), f(p[pre + "out_proj.bias"]))
    out = _layer_norm(out, f(p[pre + "layer_norm.weight"]), f(p[pre + "layer_norm.bias"]))
    return out, attn_out, masked


def acmil_mha_forward(p, x, n_token, nh=8, n_masked=0, mask_drop=0.0, rands=None, dtype=np.float32):
    """ACMIL_MHA.forward (transformer.py:69-84).  x [1, N, D_feat]; rands: per token [nh, nm] draws or None (eval).
    -> dict(sub [K, C], slide [1, C], attns [nh, K, N], masked [K][nh, keep])."""
    f = lambda a: np.asarray(a, dtype)      # noqa: E731
    h = np.maximum(f(x[0]) @ f(p["dimreduction.fc1.weight"]).T, 0)                             # network.py:49-57
    q = f(p["q"])
    subs, attns, masked = [], [], []
    for i in range(n_token):
        feat, a, m = multi_head_attention(p, f"sub_attention.{i}.", q[:, i], h, nh, n_masked, mask_drop,
                                          None if rands is None else rands[i], dtype)
        subs.append(_lin(feat, f(p[f"classifier.{i}.fc.weight"]), f(p[f"classifier.{i}.fc.bias"])))
        attns.append(a)
        masked.append(m)
    attns = np.concatenate(attns, 1)                                                           # [nh, K, N]
    pbar = _softmax(attns).mean(1, keepdims=True)                                              # [nh, 1, N]   (:82)
    vp = _heads(_lin(h, f(p["bag_attention.v_proj.weight"]), f(p["bag_attention.v_proj.bias"])), nh)
    out = (pbar @ vp).transpose(1, 0, 2).reshape(1, -1)                                        # :222-231
    out = _lin(out, f(p["bag_attention.out_proj.weight"]), f(p["bag_attention.out_proj.bias"]))
    out = _layer_norm(out, f(p["bag_attention.layer_norm.weight"]), f(p["bag_attention.layer_norm.bias"]))
    slide = _lin(out, f(p["Slide_classifier.fc.weight"]), f(p["Slide_classifier.fc.bias"]))
    return dict(sub=np.concatenate(subs, 0), slide=slide, attns=attns, masked=masked)


def mha_forward(p, x, nh=8, dtype=np.float32):
    """MHA.forward (transformer.py:97-103) -> [1, C]."""
    f = lambda a: np.asarray(a, dtype)      # noqa: E731
    h = np.maximum(f(x[0]) @ f(p["dimreduction.fc1.weight"]).T, 0)
    feat, _, _ = multi_head_attention(p, "attention.", f(p["q"])[:, 0], h, nh, dtype=dtype)
    return _lin(feat, f(p["classifier.fc.weight"]), f(p["classifier.fc.bias"]))
