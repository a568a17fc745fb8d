"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's CLAM_SB / CLAM_MB (architecture/clam.py:158-280,
bag path; Dropout off) and IBMIL (architecture/ibmil.py:69-117) forwards.  Parity status: PINNED by
tests/golden/make_golden_consumers.py, which runs the reference modules themselves; tests/test_consumers_oracle.py checks
this file against those vectors.

Nothing under acmil_b200/ imports this file.
"""
from __future__ import annotations

import numpy as np


def _lin(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def _softmax(a, axis=-1):
    e = np.exp(a - a.max(axis, keepdims=True))
    return e / e.sum(axis, keepdims=True)


def _f(p, dtype):
    return {k: np.asarray(v, dtype) for k, v in p.items()}


def clam_forward(p, x, multi_branch, gate=True, att_index=2, dtype=np.float32):
    """CLAM_SB.forward (clam.py:158-209) / CLAM_MB.forward (:241-280) without instance_eval.
    x [1, N, D_feat]; ``att_index`` = position of the attention net inside ``attention_net`` (2 without Dropout, 3 with).
    -> dict(A_raw [K, N], M [K, D_inner], logits [1, C])."""
    p = _f(p, dtype)
    pre = f"attention_net.{att_index}."
    h = np.maximum(_lin(np.asarray(x[0], dtype), p["attention_net.0.weight"], p["attention_net.0.bias"]), 0)
    if gate:                                                                    # Attn_Net_Gated.forward  clam.py:64-69
        a = np.tanh(_lin(h, p[pre + "attention_a.0.weight"], p[pre + "attention_a.0.bias"]))
        b = _sigmoid(_lin(h, p[pre + "attention_b.0.weight"], p[pre + "attention_b.0.bias"]))
        A = _lin(a * b, p[pre + "attention_c.weight"], p[pre + "attention_c.bias"])
    else:                                                                       # Attn_Net.forward        clam.py:33-34
        last = max(int(k.split(".")[3]) for k in p if k.startswith(pre + "module."))
        A = _lin(np.tanh(_lin(h, p[pre + "module.0.weight"], p[pre + "module.0.bias"])),
                 p[pre + f"module.{last}.weight"], p[pre + f"module.{last}.bias"])
    A = A.T                                                                     # [K, N]
    if multi_branch:
        e = np.exp(A)                                                           # softmax_one  utils/utils.py:54-64
        P = e / (e.sum(1, keepdims=True) + 1)
        M = P @ h
        logits = np.stack([_lin(M[c], p[f"classifiers.{c}.weight"], p[f"classifiers.{c}.bias"]) for c in range(A.shape[0])])
        logits = logits.reshape(1, -1)
    else:
        P = _softmax(A, -1)
        M = P @ h
        logits = _lin(M, p["classifiers.weight"], p["classifiers.bias"])
    return dict(A_raw=A, P=P, M=M, logits=logits)


def ibmil_forward(p, x, merge="cat", dtype=np.float32):
    """IBMIL.forward (ibmil.py:69-117).  -> dict(Y [1, C], M, A) (A = deconf_A on the deconfounded path)."""
    p = _f(p, dtype)
    h = np.maximum(np.asarray(x[0], dtype) @ p["dimreduction.fc1.weight"].T, 0)
    a = np.tanh(_lin(h, p["attention.attention_V.0.weight"], p["attention.attention_V.0.bias"]))
    b = _sigmoid(_lin(h, p["attention.attention_U.0.weight"], p["attention.attention_U.0.bias"]))
    A = _softmax(_lin(a * b, p["attention.attention_weights.weight"], p["attention.attention_weights.bias"]).T, 1)
    M = A @ h
    if "confounder_feat" in p:
        cf = p["confounder_feat"]
        q = _lin(M, p["W_q.weight"], p["W_q.bias"])
        k = _lin(cf, p["W_k.weight"], p["W_k.bias"])
        dA = _softmax((k @ q.T) / np.sqrt(dtype(k.shape[1])), 0)
        c = dA.T @ cf
        M = np.concatenate([M, c], 1) if merge == "cat" else (M + c if merge == "add" else M - c)
        return dict(Y=_lin(M, p["classifier.weight"], p["classifier.bias"]), M=M, A=dA)
    return dict(Y=_lin(M, p["classifier.fc.weight"], p["classifier.fc.bias"]), M=M, A=A)
