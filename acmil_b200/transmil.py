"""Host-side mirror of the reference's TransMIL path: ``NystromAttention`` (architecture/nystrom_attention.py:30-149,
the vendored copy of pip ``nystrom-attention`` 0.0.12), ``TransLayer`` / ``PPEG`` / ``TransMIL``
(architecture/transMIL.py:8-91).  Same constructors, forward signatures, parameter names and creation order
(reference checkpoints load with ``load_state_dict``; same initial values under the same seed); the compute runs in
libacmil_b200.so through the C-ABI of include/acmil_transmil.h.  CUDA tensors only -- there is no CPU path.

Forward only in this round: the kernels are not differentiable yet, so a forward under autograd with parameters or
inputs that require grad raises instead of silently detaching.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gemm_mode():
    """``precise`` value of the products: 2 = weight products on the fp16-split kernel (pre-split weight images, half the
    tensor-pipe time), everything else 3xTF32; ``ACMIL_GEMM_SPLIT=tf32`` keeps 3xTF32 everywhere (1)."""
    return 1 if os.environ.get("ACMIL_GEMM_SPLIT", "f16").lower() == "tf32" else 2


class SplitImage:
    """fp16 hi / lo image of a 2-D fp32 weight (``acmil_gemm_split_b``): the B operand of ``gemm_nt(..., precise=2)``."""

    def __init__(self, w2d):
        _need_cuda(w2d, "SplitImage")
        w2d = w2d.detach()
        if w2d.dim() != 2:
            raise ValueError(f"SplitImage expects a 2-D weight, got {tuple(w2d.shape)}")
        w2d = w2d if w2d.stride(1) == 1 else w2d.contiguous()
        self.rows, self.k = w2d.shape
        lib = L.load()
        nbytes = C.c_size_t(0)
        L.check(lib.acmil_gemm_split_bytes(self.rows, self.k, C.byref(nbytes)))
        self.image = torch.empty(nbytes.value, dtype=torch.uint8, device=w2d.device)
        L.check(lib.acmil_gemm_split_b(_ptr(w2d), self.rows, self.k, w2d.stride(0), _ptr(self.image), nbytes.value,
                                       _stream(w2d.device)))

    @property
    def ptr(self):
        return C.c_void_p(self.image.data_ptr())


class SplitCache:
    """Split images of a module's weights, rebuilt when a weight's storage or version counter changes (``clear()`` after
    an update through ``.data``, which bumps neither).  A new image is complete before ``get`` returns (the build is
    followed by a stream synchronize unless the stream is being captured), so other host threads / streams may use it."""

    def __init__(self):
        self._d = {}
        self._lock = threading.Lock()

    def clear(self):
        self._d.clear()

    def __deepcopy__(self, memo):      # copies and pickles of the owning module start with an empty cache
        return SplitCache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    def get(self, name, w, view=None):
        """image of ``w`` (``view(w)`` if given: e.g. a conv weight flattened to 2-D); None unless gemm_mode() == 2"""
        if gemm_mode() != 2:
            return None
        key = (w.data_ptr(), w._version, tuple(w.shape), w.device)
        with self._lock:
            hit = self._d.get(name)
            if hit is None or hit[0] != key:
                with torch.no_grad(), torch.cuda.device(w.device):
                    hit = (key, SplitImage(view(w) if view is not None else w))
                    if not torch.cuda.is_current_stream_capturing():
                        torch.cuda.current_stream().synchronize()
                self._d[name] = hit
        return hit[1]


def _need_cuda(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{what}: acmil_b200 runs on CUDA tensors only (no CPU path)")
    if x.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {x.dtype}")


def _no_grad_path(what, *tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(f"{what}: the B200 TransMIL kernels are forward-only; call under torch.no_grad()")


def gemm_nt(a, b, *, bias=None, addend=None, alpha=1.0, beta=1.0, diag=0.0, relu=False, gelu=False, precise=True,
            out=None, out_t=None, k_split=1, b_split=None, b_split_row0=0, stats_out=None, stats_in=None, _prof=None):
    """out[..., m, n] = alpha * a[..., m, k] @ b[..., n, k]^T (+ diag I) (+ bias) (+ beta * addend) (relu).

    ``a`` / ``b`` are fp32 CUDA tensors, 2-D or 3-D (leading batch; a 2-D operand is shared by the batch).
    ``out_t`` optionally receives the transposed result.  ``b_split``: a ``SplitImage`` of the 2-D weight whose rows
    ``b_split_row0 .. + n`` are ``b`` -- the product then runs on the fp16-split kernel (precise = 2).
    ``stats_out`` (fp32 ``[batch, m, ceil(n / 32), 2]``): the result is stored as exp(c - chunk max) per 32-column chunk
    with the chunk's (max, sum) in ``stats_out``; a following product that passes the same tensor as ``stats_in`` with
    those exponentials as ``a`` computes ``softmax(c) @ b^T`` (acmil_gemm_desc.softmax_stats_*)."""
    _need_cuda(a, "gemm_nt")
    _need_cuda(b, "gemm_nt")
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    b3 = b if b.dim() == 3 else b.unsqueeze(0)
    batch = max(a3.shape[0], b3.shape[0])
    def rows_ok(t):      # a row-strided 2-D view (e.g. the first k columns of a padded K-major buffer) is used in place
        return t.shape[0] == 1 and t.stride(2) == 1 and t.stride(1) % 4 == 0 and t.stride(1) >= t.shape[2] and t.data_ptr() % 16 == 0
    a3 = a3 if rows_ok(a3) else a3.contiguous()
    b3 = b3 if rows_ok(b3) else b3.contiguous()
    m, k = a3.shape[1:]
    n = b3.shape[1]
    if b3.shape[2] != k:
        raise ValueError(f"gemm_nt: inner dimensions differ ({k} vs {b3.shape[2]})")
    squeeze = a.dim() == 2 and b.dim() == 2
    if out is None:
        out = torch.empty(batch, m, n, device=a.device, dtype=torch.float32)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    g = L.GemmDesc()
    g.a, g.b, g.c = _ptr(a3), _ptr(b3), _ptr(o3)
    g.m, g.n, g.k, g.batch = m, n, k, batch
    g.lda, g.ldb, g.ldc = a3.stride(1), b3.stride(1), o3.stride(1)
    g.a_batch_stride = m * k if a3.shape[0] > 1 else 0
    g.b_batch_stride = n * k if b3.shape[0] > 1 else 0
    g.c_batch_stride = o3.stride(0) if batch > 1 else 0
    if out_t is not None:
        t3 = out_t if out_t.dim() == 3 else out_t.unsqueeze(0)
        g.ct, g.ldct, g.ct_batch_stride = _ptr(t3), t3.stride(1), t3.stride(0) if batch > 1 else 0
    if bias is not None:
        g.bias = _ptr(bias.contiguous())
    if addend is not None:
        ad = (addend if addend.dim() == 3 else addend.unsqueeze(0)).contiguous()
        g.addend, g.ld_addend, g.addend_batch_stride = _ptr(ad), ad.stride(1), ad.stride(0) if ad.shape[0] > 1 else 0
    g.alpha, g.beta, g.diag, g.precise = float(alpha), float(beta), float(diag), int(precise)
    g.act = 2 if gelu else int(relu)
    if b_split is not None:
        if b.dim() != 2 or b_split.k != k or b_split_row0 + n > b_split.rows:
            raise ValueError("gemm_nt: b_split does not match b")
        g.precise, g.b_split, g.b_split_rows, g.b_split_row0 = 2, b_split.ptr, b_split.rows, int(b_split_row0)
    for name, t, cols in (("stats_out", stats_out, n), ("stats_in", stats_in, k)):
        if t is not None:
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != batch * m * ((cols + 31) // 32) * 2:
                raise ValueError(f"gemm_nt: {name} must be a contiguous fp32 CUDA tensor [batch, m, ceil(cols / 32), 2]")
            setattr(g, "softmax_" + name, _ptr(t))
    ws = None
    if _prof is not None:      # kernel-variant builds with -DTM_GEMM_PROF=1 only (tests/cuda/gemm_h_prof.py): 16 x int64 counters
        g.split_ws = _ptr(_prof)
    if k_split > 1:
        ws = torch.empty(k_split * batch * m * n, device=a.device, dtype=torch.float32)
        g.k_split, g.split_ws = int(k_split), _ptr(ws)
    L.check(L.load().acmil_gemm_nt(C.byref(g), _stream(a.device)))
    return out[0] if squeeze and out.dim() == 3 else out


def layernorm_rows(x2d, weight, bias, eps, out=None):
    _need_cuda(x2d, "layernorm_rows")
    x2d = x2d if x2d.stride(-1) == 1 else x2d.contiguous()
    if out is None:
        out = torch.empty(x2d.shape, device=x2d.device, dtype=torch.float32)
    L.check(L.load().acmil_layernorm_rows(_ptr(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], _ptr(weight), _ptr(bias),
                                          float(eps), _ptr(out), out.stride(0), _stream(x2d.device)))
    return out


class NystromAttention(nn.Module):
    """architecture/nystrom_attention.py:30-149 -- ``forward(x[b, n, dim], mask=None, return_attn=False)``."""

    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6, residual=True,
                 residual_conv_kernel=33, eps=1e-8, dropout=0., n_token=1):
        super().__init__()
        self.eps = eps
        inner_dim = heads * dim_head
        self.n_token = n_token
        self.num_landmarks = num_landmarks
        self.pinv_iterations = pinv_iterations
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
        self.residual = residual
        self.conv_kernel = residual_conv_kernel
        if residual:
            self.res_conv = nn.Conv2d(heads, heads, (residual_conv_kernel, 1), padding=(residual_conv_kernel // 2, 0),
                                      groups=heads, bias=False)
        self.precise = True      # fp32-faithful products (gemm_mode(): fp16-split weights + 3xTF32); False = plain TF32
        self._ws = None
        self._split = SplitCache()

    def _mode(self):
        return gemm_mode() if self.precise else 0

    def _split_ptrs(self, w):
        """fills the optional split-image pointers of an acmil_nystrom_weights; returns what must stay alive"""
        if self._mode() != 2:
            return ()
        qkv, out = self._split.get("qkv", self.to_qkv.weight), self._split.get("out", self.to_out[0].weight)
        w.d_split_qkv, w.d_split_out = qkv.ptr, out.ptr
        return qkv, out

    def _run(self, x, *, ln=None, residual=None, n_out=0, padded_out=False):
        _need_cuda(x, "NystromAttention")
        if x.dim() != 3:
            raise ValueError(f"NystromAttention expects [b, n, dim], got {tuple(x.shape)}")
        params = [self.to_qkv.weight, self.to_out[0].weight, self.to_out[0].bias] + ([self.res_conv.weight] if self.residual else [])
        _no_grad_path("NystromAttention", x, residual, *params, *(ln[:2] if ln else ()))
        x = x.contiguous()
        b, n, dim = x.shape
        shape = L.NystromShape(b, n, dim, self.heads, self.dim_head, self.num_landmarks, self.pinv_iterations,
                               int(self.residual), self.conv_kernel if self.residual else 1, int(n_out), int(padded_out),
                               self._mode())
        lib = L.load()
        nbytes = C.c_size_t(0)
        L.check(lib.acmil_nystrom_workspace_bytes(C.byref(shape), C.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != x.device:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        w = L.NystromWeights()
        if ln is not None:
            w.d_ln_w, w.d_ln_b, w.ln_eps = _ptr(ln[0].contiguous()), _ptr(ln[1].contiguous()), float(ln[2])
        keep = [self.to_qkv.weight.contiguous(), self.to_out[0].weight.contiguous(), self.to_out[0].bias.contiguous()]
        w.d_wqkv, w.d_wout, w.d_bout = _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2])
        keep.extend(self._split_ptrs(w))
        if self.residual:
            keep.append(self.res_conv.weight.contiguous())
            w.d_wconv = _ptr(keep[-1])
        m = self.num_landmarks
        n_pad = n if n % m == 0 else m * math.ceil(n / m)
        rows = n_pad if padded_out else (n_out if n_out > 0 else n)
        out = torch.empty(b, rows, dim, device=x.device, dtype=torch.float32)
        res = residual.contiguous() if residual is not None else None
        L.check(lib.acmil_nystrom_attn_fwd(C.byref(shape), C.byref(w), _ptr(x), _ptr(res), _ptr(out), _ptr(self._ws),
                                           self._ws.numel(), _stream(x.device)))
        return out

    def _fwd(self, x, *, ln=None, residual=None, n_out=0):
        """LN -> attention -> (dropout) -> + residual, rows [0, n_out) only when n_out > 0."""
        p = self.to_out[1].p
        if self.training and p > 0:
            # the dropout mask covers the padded sequence (nystrom_attention.py:142-143): same RNG consumption
            n = x.shape[1]
            full = F.dropout(self._run(x, ln=ln, padded_out=True), p, True)[:, -n:]
            full = full[:, :n_out] if n_out > 0 else full
            return full + (residual[:, :full.shape[1]] if residual is not None else 0)
        return self._run(x, ln=ln, residual=residual, n_out=n_out)

    def forward(self, x, mask=None, return_attn=False):
        if mask is not None or return_attn:
            raise NotImplementedError("NystromAttention on B200: mask / return_attn are not supported (TransMIL uses neither)")
        return self._fwd(x)


class TransLayer(nn.Module):
    """architecture/transMIL.py:8-28."""

    def __init__(self, norm_layer=nn.LayerNorm, dim=512):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                     residual=True, dropout=0.1)

    def forward(self, x, n_out=0):
        if isinstance(self.norm, nn.LayerNorm) and self.norm.elementwise_affine:
            return self.attn._fwd(x, ln=(self.norm.weight, self.norm.bias, self.norm.eps), residual=x, n_out=n_out)
        y = self.attn._fwd(self.norm(x), residual=x, n_out=n_out)
        return y


class PPEG(nn.Module):
    """architecture/transMIL.py:31-45."""

    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def forward(self, x, H, W):
        _need_cuda(x, "PPEG")
        _no_grad_path("PPEG", x, *self.parameters())
        B, T, Cc = x.shape
        if T != 1 + H * W:
            raise RuntimeError(f"PPEG: {T} tokens do not form 1 + {H}x{W}")
        x = x.contiguous()
        out = torch.empty_like(x)
        ps = [t.contiguous() for t in (self.proj.weight, self.proj.bias, self.proj1.weight, self.proj1.bias,
                                       self.proj2.weight, self.proj2.bias)]
        L.check(L.load().acmil_ppeg_fwd(_ptr(x), B, H, W, Cc, *[_ptr(t) for t in ps], _ptr(out), _stream(x.device)))
        return out


class TransMIL(nn.Module):
    """architecture/transMIL.py:48-91 -- ``forward(input[B, n, D_feat]) -> logits [B, n_class]``."""

    def __init__(self, conf):
        super().__init__()
        self.pos_layer = PPEG(dim=conf.D_inner)
        self._fc1 = nn.Sequential(nn.Linear(conf.D_feat, conf.D_inner), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, conf.D_inner))
        self.n_classes = conf.n_class
        self.layer1 = TransLayer(dim=conf.D_inner)
        self.layer2 = TransLayer(dim=conf.D_inner)
        self.norm = nn.LayerNorm(conf.D_inner)
        self._fc2 = nn.Linear(conf.D_inner, conf.n_class)
        self._split = SplitCache()

    def forward(self, input):
        _need_cuda(input, "TransMIL")
        if input.dim() != 3:
            raise ValueError(f"TransMIL expects [B, n, D_feat], got {tuple(input.shape)}")
        fc1 = self._fc1[0]
        _no_grad_path("TransMIL", input, fc1.weight, fc1.bias, self.cls_token, self._fc2.weight)
        x = input.contiguous()
        B, n, _ = x.shape
        D = fc1.out_features
        # square grid, wrap-around padding, class token in front (transMIL.py:63-72)
        _H = _W = int(np.ceil(np.sqrt(n)))
        add = _H * _W - n
        h = torch.empty(B, 1 + _H * _W, D, device=x.device, dtype=torch.float32)
        gemm_nt(x, fc1.weight, bias=fc1.bias, relu=True, out=h[:, 1:1 + n],      # _fc1 (:61)
                b_split=self._split.get("fc1", fc1.weight))
        if add:
            h[:, 1 + n:] = h[:, 1:1 + add]
        h[:, 0] = self.cls_token[0, 0]
        h = self.layer1(h)                                  # (:75)
        h = self.pos_layer(h, _H, _W)                       # (:78)
        h = self.layer2(h, n_out=1)                         # (:81) only the class token is read afterwards (:84)
        cls = layernorm_rows(h[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)
        return gemm_nt(cls, self._fc2.weight, bias=self._fc2.bias)      # (:87)
