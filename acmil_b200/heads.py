"""Drop-in nn.Modules for the reference's gated-attention MIL heads.

Constructor arguments, forward signatures, return shapes and parameter names are the reference's
(checkpoints written by the reference's ``save_model`` load with ``load_state_dict``):

    ACMIL_GA, ABMIL, Attention_Gated          architecture/transformer.py:239-352
    DimReduction, Classifier_1fc              architecture/network.py:6-57
    Attention2, Attention_Gated (isNorm),
    Attention_with_Classifier                 architecture/Attention.py:6-71
    AttentionGated, DAttention                architecture/attmil.py:45-146

Every forward runs through libacmil_b200.so (acmil_b200.gated_pool.GatedPool); inputs must live on
a CUDA device -- there is no CPU path.  Gradients: the forward is the fused kernel; the backward
recomputes the (cheap to express) graph with torch ops on the same device using the mask the
kernel selected, so parameters train with the same losses as the reference.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .gated_pool import GatedPool, GatedPoolSpec

MASK_FILL = -1e9


# --------------------------------------------------------------------------------------------
# small building blocks (parameter containers; their stand-alone forwards are plain nn.Linear calls)
class Classifier_1fc(nn.Module):
    """network.py:6-19 -- optional dropout followed by one Linear."""

    def __init__(self, n_channels, n_classes, droprate=0.0):
        super().__init__()
        self.fc = nn.Linear(n_channels, n_classes)
        self.droprate = droprate
        if droprate != 0.0:
            self.dropout = nn.Dropout(p=droprate)

    def forward(self, x):
        return self.fc(self.dropout(x) if self.droprate != 0.0 else x)


class DimReduction(nn.Module):
    """network.py:37-57 -- Linear(no bias) + ReLU (residual blocks unused by every caller: numLayer_Res=0)."""

    def __init__(self, n_channels, m_dim=512, numLayer_Res=0):
        super().__init__()
        if numLayer_Res != 0:
            raise NotImplementedError("numLayer_Res > 0 is never used by the reference's heads")
        self.fc1 = nn.Linear(n_channels, m_dim, bias=False)
        self.relu1 = nn.ReLU(inplace=True)
        self.numRes = numLayer_Res
        self.resBlocks = nn.Sequential()

    def forward(self, x):
        return self.relu1(self.fc1(x))


def _gate_params(L_, D, K):
    v = nn.Sequential(nn.Linear(L_, D), nn.Tanh())
    u = nn.Sequential(nn.Linear(L_, D), nn.Sigmoid())
    w = nn.Linear(D, K)
    return v, u, w


def _torch_gate(h, m, act_a="tanh", gated=True):
    """torch-op replica of the gate, used only inside backward recomputation."""
    wv, bv = m["wv"], m["bv"]
    z = F.linear(h, wv, bv)
    a = torch.tanh(z) if act_a == "tanh" else (F.relu(z) if act_a == "relu" else F.gelu(z))
    if gated:
        a = a * torch.sigmoid(F.linear(h, m["wu"], m["bu"]))
    return F.linear(a, m["ww"], m["bw"]).transpose(0, 1)


class _PoolFn(torch.autograd.Function):
    """forward = fused kernels; backward = the kernel path of gp_backward.py (recompute-based: tcgen05 GEMMs + csrc/gp_bwd.cu)
    where it supports the head's shape, else a torch-op recomputation with the kernel's mask (ACMIL_POOL_BACKWARD=torch
    forces the latter: it is the yardstick of tests/test_gated_pool_gpu.py)."""

    @staticmethod
    def forward(ctx, runner, x, n_named, *params):
        names, tensors = params[:n_named], params[n_named:]
        res = runner(x)
        ctx.runner = runner
        ctx.names = names
        ctx.masked = res.masked_idx
        ctx.save_for_backward(x, *tensors)
        outs = [res.afeat, res.bag_feat, res.scores]
        ctx.mark_non_differentiable(*(t for t in (res.topk_idx, res.masked_idx) if t is not None))
        ctx.res = res
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_afeat, g_bag, g_scores):
        x, *tensors = ctx.saved_tensors
        spec = ctx.runner.spec
        from . import gp_backward as B
        if B.supported(spec) and os.environ.get("ACMIL_POOL_BACKWARD", "kernel") != "torch":
            # kernel path: tcgen05 GEMMs + csrc/gp_bwd.cu, recomputed from x, the raw scores and the softmax statistics
            res = ctx.res
            got = B.pool_backward(spec, x, dict(zip(ctx.names, tensors)), res.scores, res.lse_m[0], res.lse_l[0], res.afeat[0],
                                  g_afeat, g_bag, g_scores, need_dx=x.requires_grad)
            gp = [got.get(nm) if t.requires_grad else None for nm, t in zip(ctx.names, tensors)]
            return (None, got.get("x"), None) + (None,) * len(ctx.names) + tuple(gp)
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in tensors]
            m = dict(zip(ctx.names, leaves))
            xin = x.detach().float().requires_grad_(x.requires_grad)
            h = xin
            if spec.front:
                z = F.linear(xin, m["w1"], m.get("b1"))
                h = F.relu(z) if spec.front_act == "relu" else F.gelu(z)
            gm = {"wv": m["wv"], "bv": m.get("bv"), "wu": m.get("wu"), "bu": m.get("bu"), "ww": m["ww"], "bw": m.get("bw")}
            a = _torch_gate(h, gm, spec.act_a, spec.gated)           # [K, N]
            if ctx.masked is not None:
                mi = ctx.masked[0]
                valid = mi >= 0
                mask = torch.zeros_like(a, dtype=torch.bool)
                rows = torch.arange(a.shape[0], device=a.device).unsqueeze(-1).expand_as(mi)
                mask[rows[valid], mi[valid]] = True
                a = a.masked_fill(mask, MASK_FILL)
            p = torch.softmax(a, dim=1)
            afeat = p @ h
            bag = afeat.mean(0, keepdim=True)
            outs, grads = [], []
            for o, g in ((afeat.unsqueeze(0), g_afeat), (bag, g_bag), (a, g_scores)):
                if g is not None:
                    outs.append(o)
                    grads.append(g.reshape(o.shape))
            wanted = ([xin] if xin.requires_grad else []) + leaves
            got = torch.autograd.grad(outs, wanted, grads, allow_unused=True)
        it = iter(got)
        gx = next(it) if xin.requires_grad else None
        gp = [next(it) for _ in leaves]
        return (None, gx, None) + (None,) * len(ctx.names) + tuple(gp)


class _GatedPoolModule(nn.Module):
    """Shared plumbing: spec, packed-weight cache, single-bag runner."""

    _impl = L.IMPL_AUTO

    def _make_op(self, spec: GatedPoolSpec):
        object.__setattr__(self, "_op", GatedPool(spec, self._impl))

    def _weights(self):  # -> ordered dict name -> tensor or None (nn.Linear layout)
        raise NotImplementedError

    def invalidate_packed_weights(self) -> None:
        """Call after writing weights through ``param.data`` (which does not bump the tensor version the packed-weight cache
        keys on); optimizer steps, ``load_state_dict`` and in-place ops on the parameter itself are picked up automatically."""
        self._op.invalidate()

    def _pool(self, x2d: torch.Tensor, *, n_masked=0, keep=0, rsel=None, branch=None, head=None,
              slide_head=False, shared_head=False, rand=None):
        """x2d [N, d_in] -> GatedPoolResult for one bag, differentiable w.r.t. params (and x)."""
        if not x2d.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only: move the module and its input to a GPU "
                               "(the CPU reference lives in the upstream repository, not here)")
        w = self._weights()
        op: GatedPool = self._op
        # fp16 features (the dtype the reference stores them in) go to the kernels as they are: widening is exact
        x2d = (x2d if x2d.dtype == torch.float16 else x2d.to(torch.float32)).contiguous()
        n = x2d.shape[0]

        def runner(xin):
            packed = op.pack(w.get("w1"), w.get("b1"), w["wv"], w.get("bv"), w.get("wu"), w.get("bu"), w["ww"], w.get("bw"))
            bw_, bb_ = (None, None) if branch is None else branch
            hw_, hb_ = (None, None) if head is None else head
            return op.run(packed, xin.detach(), [0, n], n_masked=n_masked, keep=[keep], rsel=rsel, rand=rand, branch_w=bw_,
                          branch_b=bb_, head_w=hw_, head_b=hb_, slide_head=slide_head, shared_head=shared_head)

        runner.spec = op.spec
        need_grad = torch.is_grad_enabled() and (x2d.requires_grad or any(
            t is not None and t.requires_grad for t in w.values()))
        if not need_grad:
            return runner(x2d), None
        names = tuple(k for k, t in w.items() if t is not None)
        tensors = tuple(w[k] for k in names)
        afeat, bag, scores = _PoolFn.apply(runner, x2d, len(names), *names, *tensors)
        return None, (afeat, bag, scores)


# --------------------------------------------------------------------------------------------
class Attention_Gated(nn.Module):
    """transformer.py:239-267 (forward(x) -> raw [K, N]) and Attention.py:29-59
    (forward(x, isNorm=True) -> softmax over N unless isNorm is False).  `norm_default` selects which of
    the two reference classes this instance mirrors."""

    def __init__(self, L=512, D=128, K=1, norm_default: Optional[bool] = None):
        super().__init__()
        self.L, self.D, self.K = L, D, K
        self.attention_V, self.attention_U, self.attention_weights = _gate_params(L, D, K)
        self._norm_default = norm_default
        object.__setattr__(self, "_op", GatedPool(GatedPoolSpec(d_in=L, d_inner=L, n_branch=K, d_attn=D, front=False)))

    def _w(self):
        return dict(wv=self.attention_V[0].weight, bv=self.attention_V[0].bias, wu=self.attention_U[0].weight,
                    bu=self.attention_U[0].bias, ww=self.attention_weights.weight, bw=self.attention_weights.bias)

    def forward(self, x, isNorm=None):
        if isNorm is None:
            isNorm = bool(self._norm_default)
        if not x.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        if torch.is_grad_enabled() and (x.requires_grad or self.attention_weights.weight.requires_grad):
            a = _torch_gate(x, self._w())     # differentiable path for stand-alone use of the gate
        else:
            w = self._w()
            packed = self._op.pack(None, None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
            a = self._op.run(packed, x.to(torch.float32).contiguous(), [0, x.shape[0]]).scores
        if not isNorm:
            return a
        return F.softmax(a, dim=1) if a.requires_grad else GatedPool.softmax_rows(a)


class Attention2(nn.Module):
    """Attention.py:6-26 -- non-gated tanh attention."""

    def __init__(self, L=512, D=128, K=1):
        super().__init__()
        self.L, self.D, self.K = L, D, K
        self.attention = nn.Sequential(nn.Linear(L, D), nn.Tanh(), nn.Linear(D, K))
        object.__setattr__(self, "_op", GatedPool(GatedPoolSpec(d_in=L, d_inner=L, n_branch=K, d_attn=D, front=False,
                                                                gated=False)))

    def forward(self, x, isNorm=True):
        if not x.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        a0, a2 = self.attention[0], self.attention[2]
        if torch.is_grad_enabled() and (x.requires_grad or a0.weight.requires_grad):
            a = F.linear(torch.tanh(F.linear(x, a0.weight, a0.bias)), a2.weight, a2.bias).transpose(0, 1)
            return F.softmax(a, dim=1) if isNorm else a
        packed = self._op.pack(None, None, a0.weight, a0.bias, None, None, a2.weight, a2.bias)
        a = self._op.run(packed, x.to(torch.float32).contiguous(), [0, x.shape[0]]).scores
        return GatedPool.softmax_rows(a) if isNorm else a


class Attention_with_Classifier(_GatedPoolModule):
    """Attention.py:62-71 -- softmaxed gate, A @ x, Classifier_1fc -> [K, num_cls]."""

    def __init__(self, L=512, D=128, K=1, num_cls=2, droprate=0):
        super().__init__()
        self.attention = Attention_Gated(L, D, K, norm_default=True)
        self.classifier = Classifier_1fc(L, num_cls, droprate)
        self._make_op(GatedPoolSpec(d_in=L, d_inner=L, n_branch=K, d_attn=D, front=False))

    def _weights(self):
        return self.attention._w()

    def forward(self, x):
        plain_head = self.classifier.droprate == 0.0 or not self.training
        fc = self.classifier.fc
        res, diff = self._pool(x, head=(fc.weight, fc.bias) if plain_head else None, shared_head=plain_head)
        if res is not None and plain_head:
            return res.sub[0]
        afeat = res.afeat[0] if res is not None else diff[0][0]
        return self.classifier(afeat)


class ABMIL(_GatedPoolModule):
    """transformer.py:270-286 -- single-branch gated attention MIL -> [1, n_class]."""

    def __init__(self, conf, D=128, droprate=0):
        super().__init__()
        self.dimreduction = DimReduction(conf.D_feat, conf.D_inner)
        self.attention = Attention_Gated(conf.D_inner, D, 1)
        self.classifier = Classifier_1fc(conf.D_inner, conf.n_class, droprate)
        self._make_op(GatedPoolSpec(d_in=conf.D_feat, d_inner=conf.D_inner, n_branch=1, d_attn=D))

    def _weights(self):
        return dict(w1=self.dimreduction.fc1.weight, **self.attention._w())

    def forward(self, x):
        plain_head = self.classifier.droprate == 0.0 or not self.training
        fc = self.classifier.fc
        res, diff = self._pool(x[0], head=(fc.weight, fc.bias) if plain_head else None, shared_head=plain_head)
        if res is not None and plain_head:
            return res.sub[0]
        afeat = res.afeat[0] if res is not None else diff[0][0]
        return self.classifier(afeat)


class ACMIL_GA(_GatedPoolModule):
    """transformer.py:291-352 -- multi-branch gated attention with stochastic top-k masking.

    forward(x [1, N, D_feat]) -> (sub_preds [K, n_class], slide_pred [1, n_class], A_out [1, K, N]);
    masking is active iff ``self.training`` and n_masked_patch > 0 (transformer.py:311) and draws
    ``torch.rand(K, min(n_masked_patch, N), device=x.device)`` exactly once, like the reference.
    """

    def __init__(self, conf, D=128, droprate=0, n_token=1, n_masked_patch=0, mask_drop=0):
        super().__init__()
        self.dimreduction = DimReduction(conf.D_feat, conf.D_inner)
        self.attention = Attention_Gated(conf.D_inner, D, n_token)
        self.classifier = nn.ModuleList(Classifier_1fc(conf.D_inner, conf.n_class, droprate) for _ in range(n_token))
        self.n_masked_patch = n_masked_patch
        self.n_token = conf.n_token
        self.Slide_classifier = Classifier_1fc(conf.D_inner, conf.n_class, droprate)
        self.mask_drop = mask_drop
        self._droprate = droprate
        self._make_op(GatedPoolSpec(d_in=conf.D_feat, d_inner=conf.D_inner, n_branch=n_token, d_attn=D))

    def _weights(self):
        return dict(w1=self.dimreduction.fc1.weight, **self.attention._w())

    def _mask_args(self, n, device, use_mask):
        if not (self.n_masked_patch > 0 and use_mask):
            return 0, 0, None
        k = self.attention.K
        nm = min(self.n_masked_patch, n)
        keep = int(nm * self.mask_drop)
        # same call, shape, device and order as transformer.py:316 -> same generator stream
        # (the reference's argsort(...)[:, :keep] of the draw is taken inside acmil_gp_finish_rand: it consumes no
        # random numbers, so only the rand call has to stay here)
        rand = torch.rand(k, nm, device=device)
        if keep == 0:       # int(nm * mask_drop) == 0: the reference masks nothing (the draw above still happened)
            return 0, 0, None
        return self.n_masked_patch, keep, rand

    def _run(self, x, use_mask, with_heads):
        x0 = x[0]
        if self.n_masked_patch > L.MAX_MASKED and use_mask:
            raise ValueError(f"n_masked_patch > {L.MAX_MASKED} is not supported by the kernels")
        n_masked, keep, rand = self._mask_args(x0.shape[0], x0.device, use_mask)
        plain = with_heads and (self._droprate == 0.0 or not self.training)
        branch = head = None
        if plain:
            branch = (torch.stack([c.fc.weight for c in self.classifier]), torch.stack([c.fc.bias for c in self.classifier]))
            head = (self.Slide_classifier.fc.weight, self.Slide_classifier.fc.bias)
        return self._pool(x0, n_masked=n_masked, keep=keep, rand=None if rand is None else rand[None], branch=branch, head=head,
                          slide_head=plain), plain

    def forward(self, x):
        (res, diff), plain = self._run(x, self.training, True)
        if res is not None:
            a_out = res.scores.unsqueeze(0)
            if plain:
                return res.sub[0], res.slide, a_out
            afeat, bag = res.afeat[0], res.bag_feat
        else:
            afeat, bag, scores = diff[0][0], diff[1], diff[2]
            a_out = scores.unsqueeze(0)
        sub = torch.stack([head(afeat[i]) for i, head in enumerate(self.classifier)], dim=0)
        return sub, self.Slide_classifier(bag), a_out

    def forward_feature(self, x, use_attention_mask=False):
        (res, diff), _ = self._run(x, use_attention_mask, False)
        return res.bag_feat if res is not None else diff[1]

    @torch.no_grad()
    def forward_bags(self, x_cat, row_offsets, *, shard_begin=None, n_total=None, exchange=None, group=None,
                     want_scores=True, rand=None):
        """Several bags in ONE launch (inference / the forward of a training step; no autograd): ``x_cat`` [R, D_feat] =
        the rows of S bags back to back, ``row_offsets`` S + 1 host ints.  Returns (sub [S, K, C], slide [S, C],
        scores [K, R] or None) -- bag s equals ``forward(x_cat[row_offsets[s]:row_offsets[s + 1]][None])``, including
        the mask draw: one ``torch.rand(K, min(n_masked_patch, N_s))`` per bag, in bag order (transformer.py:316).

        Sharded bags (each rank holds a contiguous run of rows of every bag): ``shard_begin[s]`` = global index of this
        rank's first row of bag s, ``n_total[s]`` = rows of the whole bag, and either ``exchange`` (sharding.PeerExchange:
        records travel inside the kernels over NVLink) or ``group`` (NCCL all-gather of the records).  Every rank must
        mask the same patches: the draw of rank 0 is broadcast unless ``rand`` [S, K, n_masked_patch] is given."""
        if not x_cat.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        S = len(row_offsets) - 1
        sizes = [int(row_offsets[i + 1] - row_offsets[i]) for i in range(S)] if n_total is None else [int(v) for v in n_total]
        use_mask = self.training and self.n_masked_patch > 0
        n_masked, keep = 0, [0] * S
        if use_mask:
            if self.n_masked_patch > L.MAX_MASKED:
                raise ValueError(f"n_masked_patch > {L.MAX_MASKED} is not supported by the kernels")
            k = self.attention.K
            nm_max = min(self.n_masked_patch, max(sizes) if sizes else 0)
            for i, n in enumerate(sizes):
                keep[i] = int(min(self.n_masked_patch, n) * self.mask_drop)
            if rand is None:
                rand = torch.ones(S, k, max(nm_max, 1), device=x_cat.device)
                for i, n in enumerate(sizes):      # the reference's draw, bag by bag (same generator stream)
                    nm = min(self.n_masked_patch, n)
                    rand[i, :, :nm] = torch.rand(k, nm, device=x_cat.device)
                if shard_begin is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
                    torch.distributed.broadcast(rand, src=0)
            n_masked = self.n_masked_patch if any(keep) else 0
        w = self._weights()
        op = self._op
        packed = op.pack(w.get("w1"), w.get("b1"), w["wv"], w.get("bv"), w.get("wu"), w.get("bu"), w["ww"], w.get("bw"))
        res = op.run(packed, (x_cat if x_cat.dtype == torch.float16 else x_cat.to(torch.float32)).contiguous(), [int(v) for v in row_offsets], n_masked=n_masked, keep=keep,
                     rand=rand if (n_masked and use_mask) else None,
                     branch_w=torch.stack([c.fc.weight for c in self.classifier]),
                     branch_b=torch.stack([c.fc.bias for c in self.classifier]),
                     head_w=self.Slide_classifier.fc.weight, head_b=self.Slide_classifier.fc.bias, slide_head=True,
                     want_scores=want_scores, shard_begin=shard_begin, exchange=exchange, group=group)
        return res.sub, res.slide, res.scores


# --------------------------------------------------------------------------------------------
def _xavier_like_reference(module):
    """attmil.py:6-15: xavier_normal_ on every Linear weight, zero bias."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


class AttentionGated(_GatedPoolModule):
    """attmil.py:45-98 -- Linear(1024,512)+ReLU(+Dropout .25) feature layer, gate a (relu|gelu|tanh) x
    sigmoid gate b, one branch, Linear(512, 2).  forward(x [1, N, 1024]) -> [1, 2]."""

    def __init__(self, input_dim=512, act='relu', bias=False, dropout=False):
        super().__init__()
        self.L, self.D, self.K = 512, 128, 1
        self.feature = nn.Sequential(nn.Linear(1024, 512), nn.ReLU(), nn.Dropout(0.25))
        self.classifier = nn.Sequential(nn.Linear(self.L * self.K, 2))
        a = [nn.Linear(self.L, self.D, bias=bias)]
        a += [nn.GELU()] if act == 'gelu' else [nn.ReLU()] if act == 'relu' else [nn.Tanh()] if act == 'tanh' else []
        b = [nn.Linear(self.L, self.D, bias=bias), nn.Sigmoid()]
        if dropout:
            a += [nn.Dropout(0.25)]
            b += [nn.Dropout(0.25)]
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(self.D, self.K, bias=bias)
        self.apply(_xavier_like_reference)
        if act not in ("relu", "gelu", "tanh"):
            raise ValueError("act must be relu, gelu or tanh")
        self._gate_dropout = bool(dropout)
        self._make_op(GatedPoolSpec(d_in=1024, d_inner=512, n_branch=1, front_bias=True, act_a=act,
                                    gate_bias=bool(bias), score_bias=bool(bias)))

    def _weights(self):
        return dict(w1=self.feature[0].weight, b1=self.feature[0].bias, wv=self.attention_a[0].weight,
                    bv=self.attention_a[0].bias, wu=self.attention_b[0].weight, bu=self.attention_b[0].bias,
                    ww=self.attention_c.weight, bw=self.attention_c.bias)

    def forward(self, x):
        if self.training:
            raise NotImplementedError("attmil.AttentionGated: training-mode Dropout(0.25) inside the fused pass is not "
                                      "implemented; call .eval()")
        fc = self.classifier[0]
        res, diff = self._pool(x.squeeze(0), head=(fc.weight, fc.bias), shared_head=True)
        if res is not None:
            return res.sub[0]
        return self.classifier(diff[0][0])


class DAttention(_GatedPoolModule):
    """attmil.py:100-146 -- non-gated tanh attention over a Linear(1024,512)+(ReLU|GELU) feature layer.
    forward(x [1,N,1024], return_attn=False, no_norm=False)."""

    def __init__(self, n_classes, dropout, act):
        super().__init__()
        self.L, self.D, self.K = 512, 128, 1
        feat = [nn.Linear(1024, 512), nn.GELU() if act.lower() == 'gelu' else nn.ReLU()]
        if dropout:
            feat += [nn.Dropout(0.25)]
        self.feature = nn.Sequential(*feat)
        self.attention = nn.Sequential(nn.Linear(self.L, self.D), nn.Tanh(), nn.Linear(self.D, self.K))
        self.classifier = nn.Sequential(nn.Linear(self.L * self.K, n_classes))
        self.apply(_xavier_like_reference)
        self._feat_dropout = bool(dropout)
        self._make_op(GatedPoolSpec(d_in=1024, d_inner=512, n_branch=1, front_bias=True,
                                    front_act='gelu' if act.lower() == 'gelu' else 'relu', gated=False))

    def _weights(self):
        return dict(w1=self.feature[0].weight, b1=self.feature[0].bias, wv=self.attention[0].weight,
                    bv=self.attention[0].bias, ww=self.attention[2].weight, bw=self.attention[2].bias)

    def forward(self, x, return_attn=False, no_norm=False):
        if self.training and self._feat_dropout:
            raise NotImplementedError("attmil.DAttention: training-mode Dropout(0.25) inside the fused pass is not "
                                      "implemented; call .eval()")
        fc = self.classifier[0]
        res, diff = self._pool(x.squeeze(0), head=(fc.weight, fc.bias), shared_head=True)
        if res is not None:
            y, scores = res.sub[0], res.scores
        else:
            y, scores = self.classifier(diff[0][0]), diff[2]
        if not return_attn:
            return y
        if no_norm:
            return y, scores.transpose(0, 1).clone()      # A_ori: raw [N, K]
        return y, (GatedPool.softmax_rows(scores) if not scores.requires_grad else F.softmax(scores, dim=-1))
