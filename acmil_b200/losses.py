"""Losses of the reference's training loop that sit on the [K, N] score matrix, as one kernel each way.

``diversity_loss(attn)`` == the ``diff_loss`` of Step3_WSI_classification_ACMIL.py:208-214
(softmax over N, mean pairwise cosine similarity of the branches), differentiable w.r.t. ``attn``.  The reference spells
it with ~40 eager ops forward and ~400 tiny kernels through autograd per step; here it is one forward and one backward
launch (csrc/gp_bwd.cu), so that a whole training step stays a handful of launches (and CUDA-graph capturable).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class _DiversityLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):                    # a [K, n] fp32 CUDA, rows contiguous
        lib = L.load()
        K, n = a.shape
        dev = a.device
        ml = torch.empty(2 * L.MAX_BRANCH + 1, device=dev, dtype=torch.float32)
        gram = torch.empty(L.MAX_BRANCH * L.MAX_BRANCH, device=dev, dtype=torch.float32)
        div = torch.empty((), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.acmil_div_loss_fwd(_ptr(a), a.stride(0), K, n, _ptr(ml), _ptr(gram), _ptr(div), st))
        ctx.save_for_backward(a, ml, gram)
        return div

    @staticmethod
    def backward(ctx, g):
        a, ml, gram = ctx.saved_tensors
        lib = L.load()
        K, n = a.shape
        ds = torch.empty(K, n, device=a.device, dtype=torch.float32)
        g = g.to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            st = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
            L.check(lib.acmil_div_loss_bwd(_ptr(a), a.stride(0), K, n, _ptr(ml), _ptr(gram), _ptr(g), _ptr(ds), ds.stride(0), st))
        return ds


def diversity_loss(attn: torch.Tensor) -> torch.Tensor:
    """attn: the third output of ACMIL_GA.forward, [1, K, N] (or [K, N]) raw scores.  Returns the scalar diff_loss."""
    if not attn.is_cuda:
        raise RuntimeError("acmil_b200 losses run on CUDA only")
    a = attn.reshape(-1, attn.shape[-1]) if attn.dim() == 3 else attn
    if attn.dim() == 3 and attn.shape[0] != 1:
        raise ValueError("diversity_loss takes the scores of one bag ([1, K, N])")
    if a.shape[0] > L.MAX_BRANCH:
        raise ValueError(f"at most {L.MAX_BRANCH} branches")
    if a.shape[0] < 2:
        return torch.zeros((), device=attn.device, dtype=torch.float32)      # the reference's double loop is empty
    a = a.to(torch.float32)
    if a.stride(-1) != 1:
        a = a.contiguous()
    return _DiversityLoss.apply(a)
