"""Backward of the gated-attention pool on the library's own kernels.

``loss.backward()`` of Step3_WSI_classification_ACMIL.py:216-221 reaches the pool through heads._PoolFn; this module is
its kernel path.  It is recompute-based -- nothing but x, the raw scores and the softmax statistics (m, l) is kept from
the forward: h and the gate pre-activations are rebuilt with the tcgen05 GEMM engine (acmil_gemm_nt, 3xTF32 =
fp32-faithful), the row-local chain rule runs in csrc/gp_bwd.cu, and the contractions over the N rows (the weight
gradients) are K-split GEMMs over operands those kernels emit already transposed.  Formulas: csrc/gp_bwd.cu header.

Supported: a front layer with ReLU (DimReduction, CLAM's Linear+bias+ReLU), any gate flavour of the family, d_attn 128,
d_inner 128 / 256 / 512 (the reference's table, Step3_WSI_classification_ACMIL.py:69-87).  Anything else (GELU front layer of attmil.DAttention, no front layer) keeps the
torch-op recompute of heads._PoolFn.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .transmil import gemm_nt


def supported(spec) -> bool:
    return (spec.front and spec.front_act == "relu" and spec.d_attn == 128 and spec.d_inner in (128, 256, 512)
            and spec.d_in % 4 == 0)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def _k_split(m: int, n: int, k: int, sms: int = 148) -> int:
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    return max(1, min(sms // tiles, (k + 511) // 512))


@torch.no_grad()
def pool_backward(spec, x, w, scores, lse_m, lse_l, afeat, g_afeat, g_bag, g_scores, need_dx=False, _debug=None):
    """x [n, d_in]; w: dict of nn.Linear-layout weights (w1, b1?, wv, bv?, wu?, bu?, ww, bw?); scores [K, n] (raw, -1e9 at
    masked positions), lse_m / lse_l [K], afeat [K, d_inner]: the forward's outputs for this bag; g_*: upstream gradients
    (any may be None).  Returns a dict name -> gradient for every entry of ``w`` (+ 'x' when need_dx)."""
    lib = L.load()
    dev = x.device
    n, d_in = x.shape
    Li, K, D = spec.d_inner, spec.n_branch, spec.d_attn
    gated = bool(spec.gated)
    zc = 2 * D if gated else D
    f32 = dict(device=dev, dtype=torch.float32)
    x = x.detach().to(torch.float32).contiguous()
    w1 = w["w1"].detach().float().contiguous()
    b1 = w["b1"].detach().float().contiguous() if w.get("b1") is not None else None
    wv = w["wv"].detach().float()
    wcat = torch.cat([wv, w["wu"].detach().float()], 0).contiguous() if gated else wv.contiguous()       # [zc, Li]
    bcat = None
    if w.get("bv") is not None:
        bcat = torch.cat([w["bv"].detach().float(), w["bu"].detach().float()]).contiguous() if gated else w["bv"].detach().float().contiguous()
    ww = w["ww"].detach().float().contiguous()
    n4 = _pad4(n)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        # ---- recompute: h (and its transpose, the K-major operand of dWv / dWu) and the gate pre-activations
        h = torch.empty(n, Li, **f32)
        ht = torch.zeros(Li, n4, **f32) if n4 != n else torch.empty(Li, n4, **f32)
        gemm_nt(x, w1, bias=b1, relu=True, out=h, out_t=ht[:, :n])
        z = gemm_nt(h, wcat, bias=bcat)                                                                # [n, zc]
        # ---- row-local chain rule
        gate_f, relu_f = C.c_int64(0), C.c_int64(0)
        L.check(lib.acmil_gp_bwd_workspace_floats(Li, C.byref(gate_f), C.byref(relu_f)))
        partials = torch.empty(max(gate_f.value, relu_f.value), **f32)
        small = torch.empty(L.MAX_BRANCH * D + L.MAX_BRANCH + 2 * D, **f32)
        dz = torch.empty(n, zc, **f32)
        dzt = torch.zeros(zc, n4, **f32) if n4 != n else torch.empty(zc, n4, **f32)
        dhp = torch.empty(n, Li, **f32)
        sc = scores.detach()
        if sc.dtype != torch.float32 or sc.stride(-1) != 1:
            sc = sc.float().contiguous()
        gs = None
        if g_scores is not None:
            gs = g_scores.detach().reshape(K, n).float()
            if gs.stride(-1) != 1:
                gs = gs.contiguous()
        ga = None if g_afeat is None else g_afeat.detach().reshape(K, Li).float().contiguous()
        gb = None if g_bag is None else g_bag.detach().reshape(Li).float().contiguous()
        keep = (lse_m.detach().reshape(-1).float().contiguous(), lse_l.detach().reshape(-1).float().contiguous(),
                afeat.detach().reshape(K, Li).float().contiguous())
        args = L.GpBwdGateArgs(_ptr(h), _ptr(z), _ptr(sc), _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(ga), _ptr(gb),
                               _ptr(gs), _ptr(ww), n, sc.stride(0), 0 if gs is None else gs.stride(0), n4,
                               Li, D, K, L.ACT_IDS[spec.act_a], int(gated), 0,
                               _ptr(dz), _ptr(dzt), _ptr(dhp), _ptr(partials), _ptr(small))
        L.check(lib.acmil_gp_bwd_gate(C.byref(args), st))
        # ---- dh = dZ [Wv; Wu] + pool path, through the front layer's ReLU
        dh = gemm_nt(dz, wcat.t().contiguous(), addend=dhp)                                             # [n, Li]
        dz1 = torch.empty(n, Li, **f32) if need_dx else None
        dz1t = torch.zeros(Li, n4, **f32) if n4 != n else torch.empty(Li, n4, **f32)
        db1 = torch.empty(Li, **f32)
        L.check(lib.acmil_gp_bwd_relu_mask(_ptr(dh), _ptr(h), n, Li, _ptr(dz1), _ptr(dz1t), n4, _ptr(partials), _ptr(db1), st))
        # ---- weight gradients: contractions over the n rows (K-split GEMMs on K-major operands)
        xt = torch.zeros(d_in, n4, **f32) if n4 != n else torch.empty(d_in, n4, **f32)
        L.check(lib.acmil_transpose_f32(_ptr(x), d_in, n, d_in, _ptr(xt), n4, st))
        dw1 = gemm_nt(dz1t[:, :n], xt[:, :n], k_split=_k_split(Li, d_in, n))                           # [Li, d_in]
        dwcat = gemm_nt(dzt[:, :n], ht[:, :n], k_split=_k_split(zc, Li, n))                            # [zc, Li]
        out = {"w1": dw1, "wv": dwcat[:D], "ww": small[:K * D].reshape(K, D)}
        if gated:
            out["wu"] = dwcat[D:]
        if w.get("b1") is not None:
            out["b1"] = db1
        off = L.MAX_BRANCH * D
        if w.get("bw") is not None:
            out["bw"] = small[off:off + K]
        if w.get("bv") is not None:
            out["bv"] = small[off + L.MAX_BRANCH:off + L.MAX_BRANCH + D]
        if gated and w.get("bu") is not None:
            out["bu"] = small[off + L.MAX_BRANCH + D:off + L.MAX_BRANCH + 2 * D]
        if need_dx:
            out["x"] = gemm_nt(dz1, w1.t().contiguous())                                               # [n, d_in]
    if _debug is not None:      # intermediates, for tests
        _debug.update(h=h, ht=ht, z=z, dz=dz, dzt=dzt, dhp=dhp, dh=dh, dz1=dz1, dz1t=dz1t, xt=xt)
    return out
