"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's TransMIL path
(SURVEY.md section 8a rows a10-a11): NystromAttention (architecture/nystrom_attention.py, the
vendored copy of pip nystrom-attention==0.0.12 that architecture/transMIL.py:5 imports), TransLayer,
PPEG and TransMIL (architecture/transMIL.py).

Parity status: PINNED.  ``tests/golden/make_golden_transmil.py`` runs the reference modules themselves
(/root/reference, build container only) on seeded inputs and stores weights + outputs under
``tests/golden/``; ``tests/test_transmil_oracle.py`` checks every function below against them.

Eval-mode forward, ``mask=None``, ``return_attn=False`` (the only way TransMIL calls the attention).
``dtype`` float32 mimics the reference, float64 is the tighter truth.  Weights are passed as a dict keyed
by the reference's parameter names (``state_dict`` keys) under ``prefix``.
"""
from __future__ import annotations

import math

import numpy as np


def _f(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _softmax(a):
    m = a.max(axis=-1, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(axis=-1, keepdims=True)


def layernorm(x, w, b, eps=1e-5):
    """nn.LayerNorm over the last axis (biased variance): transMIL.py:11,27,58,86."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def moore_penrose_iter_pinv(x, iters=6):
    """nystrom_attention.py:12-27 -- x: [b, h, m, m]; the start value is scaled by the maxima over the WHOLE tensor."""
    ax = np.abs(x)
    col = ax.sum(axis=-1)
    row = ax.sum(axis=-2)
    z = np.swapaxes(x, -1, -2) / (col.max() * row.max())
    eye = np.eye(x.shape[-1], dtype=x.dtype)
    for _ in range(iters):
        xz = x @ z
        z = x.dtype.type(0.25) * z @ (13 * eye - (xz @ (15 * eye - (xz @ (7 * eye - xz)))))
    return z


def nystrom_attention(p, x, *, heads, dim_head, num_landmarks, pinv_iterations=6, residual=True, prefix="",
                      dtype=np.float32):
    """NystromAttention.forward(x, mask=None, return_attn=False), eval mode: nystrom_attention.py:67-140, 149."""
    x = _f(x, dtype)
    b, n, _ = x.shape
    h, m, d = heads, num_landmarks, dim_head
    wqkv = _f(p[prefix + "to_qkv.weight"], dtype)
    wout, bout = _f(p[prefix + "to_out.0.weight"], dtype), _f(p[prefix + "to_out.0.bias"], dtype)
    rem = n % m                                                     # :72-76 pad at the FRONT
    if rem > 0:
        x = np.concatenate([np.zeros((b, m - rem, x.shape[2]), dtype), x], axis=1)
    n_pad = x.shape[1]
    qkv = x @ wqkv.T                                                # :83
    q, k, v = np.split(qkv, 3, axis=-1)
    to_heads = lambda t: t.reshape(b, n_pad, h, d).transpose(0, 2, 1, 3)      # :84  b n (h d) -> b h n d
    q, k, v = to_heads(q), to_heads(k), to_heads(v)
    q = q * dtype(d ** -0.5)                                        # :93
    l = math.ceil(n / m)                                            # :97 (n is the UNPADDED length)
    ql = q.reshape(b, h, m, l, d).sum(axis=3) / dtype(l)            # :99-114
    kl = k.reshape(b, h, m, l, d).sum(axis=3) / dtype(l)
    sim1 = q @ np.swapaxes(kl, -1, -2)                              # :119-121
    sim2 = ql @ np.swapaxes(kl, -1, -2)
    sim3 = ql @ np.swapaxes(k, -1, -2)
    a1, a2, a3 = _softmax(sim1), _softmax(sim2), _softmax(sim3)     # :133
    a2 = moore_penrose_iter_pinv(a2, pinv_iterations)               # :134
    out = (a1 @ a2) @ (a3 @ v)                                      # :135
    if residual:                                                    # :137-138 depth-wise conv along the sequence
        wc = _f(p[prefix + "res_conv.weight"], dtype)               # [h, 1, ks, 1]
        ks = wc.shape[2]
        half = ks // 2
        vp = np.pad(v, ((0, 0), (0, 0), (half, half), (0, 0)))
        for t in range(ks):
            out = out + wc[None, :, 0, t, 0, None, None] * vp[:, :, t:t + n_pad]
    out = out.transpose(0, 2, 1, 3).reshape(b, n_pad, h * d)        # :141
    out = out @ wout.T + bout                                       # :142 (Dropout is the identity in eval)
    return out[:, -n:]                                              # :143


def trans_layer(p, x, *, prefix, dtype=np.float32):
    """TransLayer.forward: transMIL.py:25-28 with the constructor arguments of :12-23."""
    x = _f(x, dtype)
    dim = x.shape[-1]
    xn = layernorm(x, _f(p[prefix + "norm.weight"], dtype), _f(p[prefix + "norm.bias"], dtype))
    return x + nystrom_attention(p, xn, heads=8, dim_head=dim // 8, num_landmarks=dim // 2, pinv_iterations=6,
                                 residual=True, prefix=prefix + "attn.", dtype=dtype)


def _dwconv2d(feat, w, bias):
    """nn.Conv2d(dim, dim, k, 1, k // 2, groups=dim) on feat [B, C, H, W]."""
    ksz = w.shape[-1]
    half = ksz // 2
    fp = np.pad(feat, ((0, 0), (0, 0), (half, half), (half, half)))
    hh, ww = feat.shape[2:]
    out = np.zeros_like(feat) + bias[None, :, None, None]
    for dy in range(ksz):
        for dx in range(ksz):
            out = out + w[None, :, 0, dy, dx, None, None] * fp[:, :, dy:dy + hh, dx:dx + ww]
    return out


def ppeg(p, x, gh, gw, *, prefix="", dtype=np.float32):
    """PPEG.forward: transMIL.py:38-45."""
    x = _f(x, dtype)
    b, _, c = x.shape
    cls, feat = x[:, :1], x[:, 1:]
    cnn = feat.transpose(0, 2, 1).reshape(b, c, gh, gw)
    y = cnn
    for name in ("proj", "proj1", "proj2"):
        y = y + _dwconv2d(cnn, _f(p[prefix + name + ".weight"], dtype), _f(p[prefix + name + ".bias"], dtype))
    y = y.reshape(b, c, gh * gw).transpose(0, 2, 1)
    return np.concatenate([cls, y], axis=1)


def transmil_forward(p, x, *, dtype=np.float32):
    """TransMIL.forward: transMIL.py:60-91 -> logits [B, n_class]."""
    x = _f(x, dtype)
    h = np.maximum(x @ _f(p["_fc1.0.weight"], dtype).T + _f(p["_fc1.0.bias"], dtype), 0)      # :61
    n = h.shape[1]
    side = int(np.ceil(np.sqrt(n)))                                                            # :64-67
    add = side * side - n
    h = np.concatenate([h, h[:, :add]], axis=1)
    cls = np.broadcast_to(_f(p["cls_token"], dtype), (h.shape[0], 1, h.shape[2]))              # :70-72
    h = np.concatenate([cls, h], axis=1)
    h = trans_layer(p, h, prefix="layer1.", dtype=dtype)                                       # :75
    h = ppeg(p, h, side, side, prefix="pos_layer.", dtype=dtype)                               # :78
    h = trans_layer(p, h, prefix="layer2.", dtype=dtype)                                       # :81
    h = layernorm(h, _f(p["norm.weight"], dtype), _f(p["norm.bias"], dtype))[:, 0]             # :84
    return h @ _f(p["_fc2.weight"], dtype).T + _f(p["_fc2.bias"], dtype)                       # :87
