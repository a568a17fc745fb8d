"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's gated-attention
MIL pooling path.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's own
modules from /root/reference (read-only) in the build container, runs them on seeded
inputs and stores inputs-by-seed + weights + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Every function cites the reference file:line it restates.  ``dtype`` selects the
arithmetic type: float32 mimics the reference (IEEE fp32, ``allow_tf32`` off),
float64 gives a tighter "truth" used to judge which of two fp32 answers is closer.

The RNG of the stochastic top-k masking (``torch.rand`` at transformer.py:316) is
NOT restated: callers pass the uniform matrix ``rand`` they drew from torch, so the
generator stream stays the reference's.
"""
from __future__ import annotations

import numpy as np

MASK_FILL = -1e9  # transformer.py:320 `masked_fill(random_mask == 0, -1e9)`


# ----------------------------------------------------------------------------- helpers
def _f(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def softmax_rows(a: np.ndarray) -> np.ndarray:
    """softmax along the last axis (F.softmax(A, dim=1) on a [K, N] matrix)."""
    m = a.max(axis=-1, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(axis=-1, keepdims=True)


def log_softmax_rows(a: np.ndarray) -> np.ndarray:
    m = a.max(axis=-1, keepdims=True)
    z = a - m
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def sigmoid(z: np.ndarray) -> np.ndarray:
    return 1.0 / (1.0 + np.exp(-z))


def gelu_erf(z: np.ndarray) -> np.ndarray:
    from math import sqrt
    try:
        from scipy.special import erf
    except Exception:  # pragma: no cover
        erf = np.vectorize(__import__("math").erf)
    return 0.5 * z * (1.0 + erf(z / sqrt(2.0)))


# ----------------------------------------------------------------------------- a1
def dim_reduction(x, w1, dtype=np.float32):
    """network.py:49-57 `DimReduction.forward`: relu(x @ W1^T), no bias, numLayer_Res=0.

    x [N, D_feat], w1 [D_inner, D_feat] -> [N, D_inner]
    """
    x, w1 = _f(x, dtype), _f(w1, dtype)
    return np.maximum(x @ w1.T, 0)


# ----------------------------------------------------------------------------- a2
def attention_gated(h, wv, bv, wu, bu, ww, bw, dtype=np.float32, act_a="tanh", gated=True):
    """transformer.py:259-267 (same math Attention.py:49-54, attmil.py:87-93).

    A = (act_a(h Wv^T + bv) * sigmoid(h Wu^T + bu)) Ww^T + bw, transposed -> [K, N].
    `gated=False` gives the non-gated flavour (Attention.py:6-26, attmil.py:117-121).
    Biases may be None (attmil.AttentionGated(bias=False)).
    """
    h = _f(h, dtype)
    zv = h @ _f(wv, dtype).T
    if bv is not None:
        zv = zv + _f(bv, dtype)
    if act_a == "tanh":
        a = np.tanh(zv)
    elif act_a == "relu":
        a = np.maximum(zv, 0)
    elif act_a == "gelu":
        a = gelu_erf(zv).astype(dtype)
    else:
        raise ValueError(act_a)
    if gated:
        zu = h @ _f(wu, dtype).T
        if bu is not None:
            zu = zu + _f(bu, dtype)
        a = a * sigmoid(zu)
    s = a @ _f(ww, dtype).T
    if bw is not None:
        s = s + _f(bw, dtype)
    return np.ascontiguousarray(s.T)


# ----------------------------------------------------------------------------- a3
def topk_indices(a: np.ndarray, k: int) -> np.ndarray:
    """torch.topk(A, k, dim=-1).indices: k largest per row, sorted descending
    (transformer.py:315).  Equal scores: lower index first (stable), which is what
    torch's CPU topk produces for the sort-based path; exact ties do not occur in the
    fixtures."""
    idx = np.argsort(-a, axis=-1, kind="stable")[:, :k]
    return idx.astype(np.int64)


def stkim_select(rand: np.ndarray, n_masked: int, mask_drop: float) -> np.ndarray:
    """transformer.py:316: argsort(rand, dim=-1)[:, :int(n_masked * mask_drop)].

    `rand` is the [K, n_masked] uniform matrix the caller drew with torch.rand.
    """
    keep = int(n_masked * mask_drop)
    return np.argsort(rand, axis=-1, kind="stable")[:, :keep].astype(np.int64)


def stkim_mask(a: np.ndarray, n_masked_patch: int, mask_drop: float, rand: np.ndarray):
    """transformer.py:311-320: per branch, of the top-`n` scoring patches a random
    int(n*mask_drop) are set to -1e9.  Returns (masked A, masked_indices [K, keep])."""
    k, n = a.shape
    n_masked = min(n_masked_patch, n)
    top = topk_indices(a, n_masked)
    rsel = stkim_select(rand, n_masked, mask_drop)
    masked_indices = np.take_along_axis(top, rsel, axis=-1)
    out = a.copy()
    for b in range(k):
        out[b, masked_indices[b]] = MASK_FILL
    return out, masked_indices


# ----------------------------------------------------------------------------- a4/a5
def acmil_ga_forward(p: dict, x, training=False, n_masked_patch=0, mask_drop=0.0,
                     rand=None, dtype=np.float32):
    """transformer.py:305-330 `ACMIL_GA.forward`.

    p: state_dict-style mapping (names as SURVEY.md section 8b):
       dimreduction.fc1.weight, attention.attention_V.0.{weight,bias},
       attention.attention_U.0.{weight,bias}, attention.attention_weights.{weight,bias},
       classifier.{i}.fc.{weight,bias}, Slide_classifier.fc.{weight,bias}
    x: [1, N, D_feat].  Returns dict(sub [K,C], slide [1,C], A_out [1,K,N],
       afeat [K,D_inner], bag_feat [1,D_inner], masked_indices or None).
    """
    x0 = _f(x, dtype)[0]                                        # :306
    h = dim_reduction(x0, p["dimreduction.fc1.weight"], dtype)  # :307
    a = attention_gated(h, p["attention.attention_V.0.weight"], p["attention.attention_V.0.bias"],
                        p["attention.attention_U.0.weight"], p["attention.attention_U.0.bias"],
                        p["attention.attention_weights.weight"], p["attention.attention_weights.bias"],
                        dtype)                                  # :308
    masked_indices = None
    if n_masked_patch > 0 and training:                         # :311
        if rand is None:
            raise ValueError("training-mode masking needs the torch.rand matrix")
        a, masked_indices = stkim_mask(a, n_masked_patch, mask_drop, np.asarray(rand))
        a = a.astype(dtype)
    a_out = a                                                   # :322
    pr = softmax_rows(a)                                        # :323
    afeat = pr @ h                                              # :324
    k = a.shape[0]
    sub = np.stack([afeat[i] @ _f(p[f"classifier.{i}.fc.weight"], dtype).T
                    + _f(p[f"classifier.{i}.fc.bias"], dtype) for i in range(k)], 0)  # :325-327
    bag_a = softmax_rows(a_out).mean(0, keepdims=True)          # :328
    bag_feat = bag_a @ h                                        # :329
    slide = bag_feat @ _f(p["Slide_classifier.fc.weight"], dtype).T + _f(p["Slide_classifier.fc.bias"], dtype)
    return dict(sub=sub, slide=slide, A_out=a_out[None], afeat=afeat, bag_feat=bag_feat,
                masked_indices=masked_indices, h=h)


def acmil_ga_forward_feature(p, x, use_attention_mask=False, n_masked_patch=0, mask_drop=0.0,
                             rand=None, dtype=np.float32):
    """transformer.py:332-352 `ACMIL_GA.forward_feature` -> bag_feat [1, D_inner]."""
    r = acmil_ga_forward(p, x, training=use_attention_mask, n_masked_patch=n_masked_patch,
                         mask_drop=mask_drop, rand=rand, dtype=dtype)
    return r["bag_feat"]


# ----------------------------------------------------------------------------- a7
def abmil_forward(p: dict, x, dtype=np.float32):
    """transformer.py:277-286 `ABMIL.forward` -> [1, n_class] (single branch, eval)."""
    x0 = _f(x, dtype)[0]
    h = dim_reduction(x0, p["dimreduction.fc1.weight"], dtype)
    a = attention_gated(h, p["attention.attention_V.0.weight"], p["attention.attention_V.0.bias"],
                        p["attention.attention_U.0.weight"], p["attention.attention_U.0.bias"],
                        p["attention.attention_weights.weight"], p["attention.attention_weights.bias"],
                        dtype)
    afeat = softmax_rows(a) @ h
    out = afeat @ _f(p["classifier.fc.weight"], dtype).T + _f(p["classifier.fc.bias"], dtype)
    return dict(out=out, A=a, afeat=afeat)


# ----------------------------------------------------------------------------- a8
def branch_diversity_loss(a_out, dtype=np.float32):
    """Step3_WSI_classification_ACMIL.py:208-214: mean pairwise cosine similarity of
    the softmaxed branch attentions.  a_out [1, K, N] -> scalar.
    torch.cosine_similarity clamps each norm at eps=1e-8."""
    pr = softmax_rows(_f(a_out, dtype))[0]
    k = pr.shape[0]
    if k < 2:
        return dtype(0.0)
    tot = dtype(0.0)
    for i in range(k):
        for j in range(i + 1, k):
            ni = max(float(np.sqrt((pr[i] * pr[i]).sum())), 1e-8)
            nj = max(float(np.sqrt((pr[j] * pr[j]).sum())), 1e-8)
            tot = tot + dtype((pr[i] * pr[j]).sum() / (ni * nj)) / dtype(k * (k - 1) / 2)
    return tot


def attention_entropy_loss(a_out, dtype=np.float32):
    """Step3_WSI_classification_ACMIL.py:259: sum(softmax * log_softmax) / K."""
    a = _f(a_out, dtype)
    return (softmax_rows(a) * log_softmax_rows(a)).sum() / dtype(a.shape[1])


# ----------------------------------------------------------------------------- a9
def attention_with_classifier(p: dict, x, dtype=np.float32):
    """Attention.py:67-71 `Attention_with_Classifier.forward`: softmaxed gate (isNorm
    default True) -> AA @ x -> Classifier_1fc.  x [N, L] -> [K, num_cls]."""
    x = _f(x, dtype)
    a = attention_gated(x, p["attention.attention_V.0.weight"], p["attention.attention_V.0.bias"],
                        p["attention.attention_U.0.weight"], p["attention.attention_U.0.bias"],
                        p["attention.attention_weights.weight"], p["attention.attention_weights.bias"], dtype)
    afeat = softmax_rows(a) @ x
    return afeat @ _f(p["classifier.fc.weight"], dtype).T + _f(p["classifier.fc.bias"], dtype)


def attmil_attention_gated(p: dict, x, act="relu", dtype=np.float32):
    """attmil.py:84-98 `AttentionGated.forward` (eval: Dropout inactive).
    x [1, N, 1024] -> Y_prob [1, 2]."""
    x0 = _f(x, dtype)
    x0 = x0.reshape(-1, x0.shape[-1]) if x0.ndim == 3 else x0
    h = np.maximum(x0 @ _f(p["feature.0.weight"], dtype).T + _f(p["feature.0.bias"], dtype), 0)
    a = attention_gated(h, p["attention_a.0.weight"], p.get("attention_a.0.bias"),
                        p["attention_b.0.weight"], p.get("attention_b.0.bias"),
                        p["attention_c.weight"], p.get("attention_c.bias"), dtype, act_a=act)
    m = softmax_rows(a) @ h
    return dict(out=m @ _f(p["classifier.0.weight"], dtype).T + _f(p["classifier.0.bias"], dtype), A=a)


def attmil_dattention(p: dict, x, act="relu", dtype=np.float32):
    """attmil.py:128-146 `DAttention.forward` (eval).  x [1, N, 1024] ->
    (Y_prob [1, n_classes], A softmaxed [1, N], A_ori raw [N, 1])."""
    x0 = _f(x, dtype)[0]
    z = x0 @ _f(p["feature.0.weight"], dtype).T + _f(p["feature.0.bias"], dtype)
    h = gelu_erf(z).astype(dtype) if act.lower() == "gelu" else np.maximum(z, 0)
    a = attention_gated(h, p["attention.0.weight"], p["attention.0.bias"], None, None,
                        p["attention.2.weight"], p["attention.2.bias"], dtype, gated=False)
    pr = softmax_rows(a)
    m = pr @ h
    return dict(out=m @ _f(p["classifier.0.weight"], dtype).T + _f(p["classifier.0.bias"], dtype),
                A=pr, A_ori=np.ascontiguousarray(a.T))


# ----------------------------------------------------------------------------- sharding
def pool_partials(h: np.ndarray, a: np.ndarray):
    """Per-shard online-softmax partials used by the row-sharded head:
    m_k = max_n A[k,n]; l_k = sum_n exp(A-m); acc_k = sum_n exp(A-m) h_n."""
    m = a.max(axis=1)
    e = np.exp(a - m[:, None])
    return m, e.sum(axis=1), e @ h


def merge_partials(ms, ls, accs):
    """LSE merge of row-shard partials -> afeat [K, D_inner] (exactly softmax(A) @ h)."""
    ms, ls, accs = np.stack(ms), np.stack(ls), np.stack(accs)   # [P,K], [P,K], [P,K,D]
    m = ms.max(axis=0)
    w = np.exp(ms - m[None])
    l = (w * ls).sum(axis=0)
    acc = (w[:, :, None] * accs).sum(axis=0)
    return acc / l[:, None], m, l
