"""ACMIL_MHA / MHA -- the other ``--arch`` of the reference's Step3 CLI (architecture/transformer.py:50-236) on the
gated-pool kernels.

Every sub-attention has ONE query (a learned token), so its 8-head attention over the N patches collapses algebraically:

    score_hd[n] = (q'_hd . k'_n,hd) / sqrt(d)          with k'_n = Wk h_n + bk          (transformer.py:152-160)
                = h_n . u_hd + c_hd,                    u_hd = Wk[hd]^T q'_hd / sqrt(d),  c_hd = bk[hd] . q'_hd / sqrt(d)
    out_hd      = sum_n softmax(score_hd)[n] v'_n,hd    with v'_n = Wv h_n + bv           (:176-179)
                = Wv[hd] (sum_n softmax(score_hd)[n] h_n) + bv[hd]

i.e. per token a softmax-pool of h = relu(x W1^T) under 8 LINEAR score vectors -- the pooling head of the gated-attention
path with n_branch = 8, no gate nonlinearity, and the same stochastic top-k masking (:162-174: top n_masked_patch scores
per head row, a random mask_drop share of them set to -1e9 before the softmax).  A linear score is exact on the
kernel's gate  A = relu(h Wv'^T) Ww'^T + bw'  with Wv' = [u_0, -u_0, u_1, -u_1, ...] and Ww' = [+1, -1] pairs, since
relu(z) - relu(-z) = z.  The bag feature (MutiHeadAttention_modify, :184-236) averages the tokens' softmaxes per head,
which is the mean of the pooled h the token calls already produced: no further pass over the rows.

forward(x [1, N, D_feat]) -> (sub [K, C], slide [1, C], attns [8, K, N]) like the reference (heads-first layout).
h is computed once (tcgen05 GEMM, relu epilogue) and pooled K times (exact-fp32 FFMA pool kernel, general shapes).
Under grad mode the projections are built with torch ops and the pool goes through the recompute backward of heads._PoolFn.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from .gated_pool import GatedPool, GatedPoolSpec
from .heads import Classifier_1fc, DimReduction, _PoolFn


class MutiHeadAttention(nn.Module):
    """transformer.py:105-182.  forward(q [1, 1, E], k [1, N, E], v = k) -> (out [1, E], attn_out [heads, 1, N]).
    (sic: the reference's spelling.)  k and v must be the same tensor, as at every call site of the reference."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1, dropout: float = 0.1,
                 n_masked_patch: int = 0, mask_drop: float = 0.0) -> None:
        super().__init__()
        self.n_masked_patch = n_masked_patch
        self.mask_drop = mask_drop
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)
        self.layer_norm = nn.LayerNorm(embedding_dim, eps=1e-6)
        self.dropout = nn.Dropout(dropout)
        if num_heads > L.MAX_BRANCH:
            raise ValueError(f"num_heads must be <= {L.MAX_BRANCH}")
        spec = GatedPoolSpec(d_in=embedding_dim, d_inner=embedding_dim, n_branch=num_heads, d_attn=128, front=False,
                             act_a="relu", gated=False, gate_bias=False, score_bias=True)
        object.__setattr__(self, "_op", GatedPool(spec))

    # ---- the single-query algebra of the module docstring ----
    def _score_vectors(self, q):
        """q [1, 1, E] -> (u [heads, E], c [heads]): score_hd[n] = h_n . u_hd + c_hd."""
        hd, d = self.num_heads, self.internal_dim // self.num_heads
        qp = self.q_proj(q).reshape(hd, d)                                          # q'_hd
        wk = self.k_proj.weight.reshape(hd, d, self.embedding_dim)                  # Wk[hd]
        u = torch.einsum("hd,hde->he", qp, wk) / math.sqrt(d)
        c = (self.k_proj.bias.reshape(hd, d) * qp).sum(-1) / math.sqrt(d)
        return u, c

    def _gate_weights(self, u, c):
        """kernel gate weights for linear scores: relu(h.u) - relu(-h.u) = h.u."""
        hd, e = u.shape
        wv = torch.zeros(128, e, device=u.device, dtype=u.dtype)
        wv[0:2 * hd:2] = u
        wv[1:2 * hd:2] = -u
        ww = torch.zeros(hd, 128, device=u.device, dtype=u.dtype)
        idx = torch.arange(hd, device=u.device)
        ww[idx, 2 * idx] = 1.0
        ww[idx, 2 * idx + 1] = -1.0
        return wv, ww, c

    def pool(self, h2d: torch.Tensor, q: torch.Tensor):
        """h2d [N, E] -> (pooled h per head [heads, E] under the (masked) softmax, attn_out [heads, N])."""
        u, c = self._score_vectors(q)
        wv, ww, bw = self._gate_weights(u, c)
        n = h2d.shape[0]
        n_masked = keep = 0
        rand = None
        if self.n_masked_patch > 0 and self.training:
            nm = min(self.n_masked_patch, n)
            # same call / shape / device as transformer.py:168 -> same generator stream
            rand = torch.rand(self.num_heads, nm, device=h2d.device)
            keep = int(nm * self.mask_drop)
            n_masked = self.n_masked_patch if keep > 0 else 0
        op: GatedPool = self._op
        need_grad = torch.is_grad_enabled() and (h2d.requires_grad or wv.requires_grad)

        def runner(xin):
            packed = op.pack(None, None, wv.detach(), None, None, None, ww, bw.detach())
            return op.run(packed, xin.detach(), [0, n], n_masked=n_masked, keep=[keep],
                          rand=None if (rand is None or n_masked == 0) else rand[None])

        runner.spec = op.spec
        if not need_grad:
            res = runner(h2d)
            return res.afeat[0], res.scores
        names = ("wv", "ww", "bw")
        afeat, _, scores = _PoolFn.apply(runner, h2d, len(names), *names, wv, ww, bw)
        return afeat[0], scores

    def _finish(self, pooled):
        """pooled h per head [heads, E] -> out1[0] [1, E] (v_proj per head, recombine, out_proj, dropout, LayerNorm)."""
        hd, d = self.num_heads, self.internal_dim // self.num_heads
        wv = self.v_proj.weight.reshape(hd, d, self.embedding_dim)
        out = torch.einsum("he,hde->hd", pooled, wv) + self.v_proj.bias.reshape(hd, d)
        out = self.out_proj(out.reshape(1, hd * d))
        return self.layer_norm(self.dropout(out))

    def forward(self, q, k, v):
        if k is not v:
            raise NotImplementedError("acmil_b200.MutiHeadAttention: k and v must be the same tensor (every reference call site)")
        if not k.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        pooled, scores = self.pool(k[0].to(torch.float32).contiguous(), q)
        return self._finish(pooled), scores.unsqueeze(1)


class MutiHeadAttention_modify(nn.Module):
    """transformer.py:184-236 -- value projection + pooling under a GIVEN attention [heads, 1, N]."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1, dropout: float = 0.1) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)
        self.layer_norm = nn.LayerNorm(embedding_dim, eps=1e-6)
        self.dropout = nn.Dropout(dropout)

    def from_pooled(self, pooled):
        """pooled [heads, E] = attn[hd] @ h (the caller already has it) -> [1, E]."""
        hd, d = self.num_heads, self.internal_dim // self.num_heads
        wv = self.v_proj.weight.reshape(hd, d, self.embedding_dim)
        out = torch.einsum("he,hde->hd", pooled, wv) + self.v_proj.bias.reshape(hd, d) * 1.0
        out = self.out_proj(out.reshape(1, hd * d))
        return self.layer_norm(self.dropout(out))

    def forward(self, v, attn):
        # attn rows sum to 1 (a mean of softmaxes at the only call site), so the value bias passes through unchanged
        pooled = torch.einsum("hn,ne->he", attn.reshape(self.num_heads, -1).to(v.dtype), v[0])
        return self.from_pooled(pooled)


def _front(dimreduction: DimReduction, x2d: torch.Tensor) -> torch.Tensor:
    """h = relu(x W1^T) (network.py:49-57): tcgen05 GEMM with the relu in its epilogue, torch ops under grad mode."""
    w1 = dimreduction.fc1.weight
    if torch.is_grad_enabled() and (x2d.requires_grad or w1.requires_grad):
        return F.relu(F.linear(x2d, w1))
    from .transmil import gemm_nt
    return gemm_nt(x2d, w1.detach(), relu=True).reshape(x2d.shape[0], w1.shape[0])


class ACMIL_MHA(nn.Module):
    """transformer.py:50-84.  forward(x [1, N, D_feat]) -> (sub [K, C], slide [1, C], attns [8, K, N])."""

    def __init__(self, conf, n_token=1, n_masked_patch=0, mask_drop=0):
        super().__init__()
        self.dimreduction = DimReduction(conf.D_feat, conf.D_inner)
        self.sub_attention = nn.ModuleList()
        for _ in range(n_token):
            self.sub_attention.append(MutiHeadAttention(conf.D_inner, 8, n_masked_patch=n_masked_patch, mask_drop=mask_drop))
        self.bag_attention = MutiHeadAttention_modify(conf.D_inner, 8)
        self.q = nn.Parameter(torch.zeros((1, n_token, conf.D_inner)))
        nn.init.normal_(self.q, std=1e-6)
        self.n_class = conf.n_class
        self.classifier = nn.ModuleList()
        for _ in range(n_token):
            self.classifier.append(Classifier_1fc(conf.D_inner, conf.n_class, 0.0))
        self.n_token = n_token
        self.Slide_classifier = Classifier_1fc(conf.D_inner, conf.n_class, 0.0)

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        h = _front(self.dimreduction, input[0].to(torch.float32).contiguous())
        outputs, attns, pooled = [], [], []
        for i in range(self.n_token):
            att = self.sub_attention[i]
            p_i, a_i = att.pool(h, self.q[:, i].unsqueeze(0))
            outputs.append(self.classifier[i](att._finish(p_i)))
            attns.append(a_i.unsqueeze(1))
            pooled.append(p_i)
        attns = torch.cat(attns, 1)
        # bag_attention(v, attns.softmax(-1).mean(1)): the mean over tokens of softmax(attn_token) @ h per head
        feat_bag = self.bag_attention.from_pooled(torch.stack(pooled).mean(0))
        return torch.cat(outputs, dim=0), self.Slide_classifier(feat_bag), attns


class MHA(nn.Module):
    """transformer.py:87-103.  forward(x [1, N, D_feat]) -> [1, C]."""

    def __init__(self, conf):
        super().__init__()
        self.dimreduction = DimReduction(conf.D_feat, conf.D_inner)
        self.attention = MutiHeadAttention(conf.D_inner, 8)
        self.q = nn.Parameter(torch.zeros((1, 1, conf.D_inner)))
        nn.init.normal_(self.q, std=1e-6)
        self.n_class = conf.n_class
        self.classifier = Classifier_1fc(conf.D_inner, conf.n_class, 0.0)

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        h = _front(self.dimreduction, input[0].to(torch.float32).contiguous())
        att = self.attention
        pooled, _ = att.pool(h, self.q)
        return self.classifier(att._finish(pooled))
