"""The other consumers of the gated-attention pool in the reference, on the same kernels:

    Attn_Net, Attn_Net_Gated, CLAM_SB, CLAM_MB        architecture/clam.py:18-280
    IBMIL (+ its Attention_Gated)                     architecture/ibmil.py:7-117

Constructor arguments, forward signatures, return values and parameter names are the reference's, so its checkpoints
load with ``load_state_dict``.  The bag pass -- Linear(+bias)+ReLU front layer, tanh x sigmoid gate, softmax over N,
A @ h -- is one GatedPool call; the instance-level clustering loss of CLAM (clam.py:128-156) touches 2 x k_sample rows
of h only, which are recomputed from their x rows instead of materialising h [N, D_inner].
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .gated_pool import GatedPool, GatedPoolSpec
from .heads import Attention_Gated as _AttentionGatedBase
from .heads import Classifier_1fc, DimReduction, _GatedPoolModule


def initialize_weights(module):
    """utils/utils.py:519-527."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def softmax_one(x, dim=-1):
    """utils/utils.py:54-64 (no max subtraction, +1 in the denominator)."""
    e = torch.exp(x)
    return e / (e.sum(dim=dim, keepdim=True) + 1)


def _scores_only(op: GatedPool, x, wv, bv, wu, bu, ww, bw):
    packed = op.pack(None, None, wv, bv, wu, bu, ww, bw)
    return op.run(packed, x.to(torch.float32).contiguous(), [0, x.shape[0]]).scores      # [K, N]


class Attn_Net(nn.Module):
    """clam.py:18-34 -- Linear, Tanh, [Dropout .25], Linear.  forward(x [N, L]) -> (A [N, n_classes], x)."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        mods = [nn.Linear(L, D), nn.Tanh()]
        if dropout:
            mods.append(nn.Dropout(0.25))
        mods.append(nn.Linear(D, n_classes))
        self.module = nn.Sequential(*mods)
        self._dropout = bool(dropout)
        self._dims = (L, D, n_classes)
        object.__setattr__(self, "_op", None)

    def _w(self):
        a, c = self.module[0], self.module[-1]
        return dict(wv=a.weight, bv=a.bias, ww=c.weight, bw=c.bias)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        w = self._w()
        if (self.training and self._dropout) or (torch.is_grad_enabled() and (x.requires_grad or w["wv"].requires_grad)):
            return self.module(x), x            # stand-alone differentiable use of the small module
        if self._op is None:
            L_, D, K = self._dims
            object.__setattr__(self, "_op", GatedPool(GatedPoolSpec(d_in=L_, d_inner=L_, n_branch=K, d_attn=D, front=False,
                                                                    gated=False)))
        return _scores_only(self._op, x, w["wv"], w["bv"], None, None, w["ww"], w["bw"]).transpose(0, 1), x


class Attn_Net_Gated(nn.Module):
    """clam.py:46-69 -- tanh branch a, sigmoid branch b, Linear(D, n_classes) on a*b.  forward(x) -> (A [N, K], x)."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        a = [nn.Linear(L, D), nn.Tanh()]
        b = [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_classes)
        self._dropout = bool(dropout)
        self._dims = (L, D, n_classes)
        object.__setattr__(self, "_op", None)

    def _w(self):
        a, b, c = self.attention_a[0], self.attention_b[0], self.attention_c
        return dict(wv=a.weight, bv=a.bias, wu=b.weight, bu=b.bias, ww=c.weight, bw=c.bias)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("acmil_b200 modules run on CUDA only")
        w = self._w()
        if (self.training and self._dropout) or (torch.is_grad_enabled() and (x.requires_grad or w["wv"].requires_grad)):
            return self.attention_c(self.attention_a(x).mul(self.attention_b(x))), x
        if self._op is None:
            L_, D, K = self._dims
            object.__setattr__(self, "_op", GatedPool(GatedPoolSpec(d_in=L_, d_inner=L_, n_branch=K, d_attn=D, front=False)))
        return _scores_only(self._op, x, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"]).transpose(0, 1), x


class CLAM_SB(_GatedPoolModule):
    """clam.py:85-209.  forward(h [1, N, D_feat], label=None, instance_eval=False, return_features=False,
    attention_only=False) -> logits [1, n_class] (, instance loss) | raw A [K, N] when attention_only."""

    _multi_branch = False

    def __init__(self, conf, gate=True, size_arg="small", k_sample=8, dropout=True,
                 instance_loss_fn=nn.CrossEntropyLoss()):
        super().__init__()
        n_classes = conf.n_class
        self.size_dict = {"small": [conf.D_feat, conf.D_inner, 128], "big": [conf.D_feat, 512, 384]}
        size = self.size_dict[size_arg]
        k_att = n_classes if self._multi_branch else 1
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        fc.append((Attn_Net_Gated if gate else Attn_Net)(L=size[1], D=size[2], dropout=dropout, n_classes=k_att))
        self.attention_net = nn.Sequential(*fc)
        if self._multi_branch:
            self.classifiers = nn.ModuleList([nn.Linear(size[1], 1) for _ in range(n_classes)])
        else:
            self.classifiers = nn.Linear(size[1], n_classes)
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample = k_sample
        self.instance_loss_fn = instance_loss_fn
        self.n_classes = n_classes
        self.subtyping = conf.n_class > 2
        initialize_weights(self)
        self._dropout = bool(dropout)
        self._gate = bool(gate)
        self._make_op(GatedPoolSpec(d_in=size[0], d_inner=size[1], n_branch=k_att, d_attn=size[2], front_bias=True,
                                    gated=bool(gate)))

    def relocate(self):
        """clam.py:113-117."""
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.attention_net = self.attention_net.to(device)
        self.classifiers = self.classifiers.to(device)
        self.instance_classifiers = self.instance_classifiers.to(device)

    @staticmethod
    def create_positive_targets(length, device):
        return torch.full((length,), 1, device=device).long()

    @staticmethod
    def create_negative_targets(length, device):
        return torch.full((length,), 0, device=device).long()

    def _weights(self):
        front, att = self.attention_net[0], self.attention_net[-1]
        return dict(w1=front.weight, b1=front.bias, **att._w())

    # ---- instance-level clustering (clam.py:128-156): h rows of the selected patches only
    def _h_rows(self, x2d, ids):
        front = self.attention_net[0]
        return F.relu(F.linear(x2d.index_select(0, ids), front.weight, front.bias))

    def inst_eval(self, A, x2d, classifier):
        if A.dim() == 1:
            A = A.view(1, -1)
        top_p_ids = torch.topk(A, self.k_sample)[1][-1]
        top_n_ids = torch.topk(-A, self.k_sample, dim=1)[1][-1]
        inst = torch.cat([self._h_rows(x2d, top_p_ids), self._h_rows(x2d, top_n_ids)], dim=0)
        targets = torch.cat([self.create_positive_targets(self.k_sample, x2d.device),
                             self.create_negative_targets(self.k_sample, x2d.device)], dim=0)
        logits = classifier(inst)
        preds = torch.topk(logits, 1, dim=1)[1].squeeze(1)
        return self.instance_loss_fn(logits, targets), preds, targets

    def inst_eval_out(self, A, x2d, classifier):
        if A.dim() == 1:
            A = A.view(1, -1)
        top_p_ids = torch.topk(A, self.k_sample)[1][-1]
        targets = self.create_negative_targets(self.k_sample, x2d.device)
        logits = classifier(self._h_rows(x2d, top_p_ids))
        preds = torch.topk(logits, 1, dim=1)[1].squeeze(1)
        return self.instance_loss_fn(logits, targets), preds, targets

    def _instance_loss(self, A, x2d, label):
        total, inst_labels = 0.0, F.one_hot(label, num_classes=self.n_classes).squeeze()
        for i, classifier in enumerate(self.instance_classifiers):
            Ai = A[i] if self._multi_branch else A
            if inst_labels[i].item() == 1:
                loss, _, _ = self.inst_eval(Ai, x2d, classifier)
            elif self.subtyping:
                loss, _, _ = self.inst_eval_out(Ai, x2d, classifier)
            else:
                continue
            total = total + loss
        if self.subtyping:
            total = total / len(self.instance_classifiers)
        return total

    def _normalise(self, scores):
        return GatedPool.softmax_rows(scores) if not scores.requires_grad else F.softmax(scores, dim=-1)

    def _pooled(self, res, diff):
        """-> (M [K, D_inner] under this class's normalisation, raw scores [K, N])."""
        if res is not None:
            return res.afeat[0], res.scores
        return diff[0][0], diff[2]

    def _logits(self, M):
        return self.classifiers(M)

    def forward(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        if self.training and self._dropout:
            raise NotImplementedError("acmil_b200 CLAM: training-mode Dropout(0.25) inside the fused pass is not implemented; "
                                      "construct with dropout=False or call .eval()")
        x2d = h[0]
        res, diff = self._pool(x2d)
        M, scores = self._pooled(res, diff)
        if attention_only:
            return scores
        if instance_eval:
            inst_loss = self._instance_loss(self._normalise(scores), x2d.to(torch.float32), label)
        logits = self._logits(M)
        return (logits, inst_loss) if instance_eval else logits


class CLAM_MB(CLAM_SB):
    """clam.py:212-280 -- one attention branch and one Linear(D_inner, 1) per class, ``softmax_one`` over N."""

    _multi_branch = True

    def _normalise(self, scores):
        return softmax_one(scores, dim=1)

    def _pooled(self, res, diff):
        # softmax_one(A) @ h = softmax(A) @ h * S / (1 + S) with S = sum exp(A) = l * exp(m)   (utils.py:54-64)
        if res is not None:
            s = res.lse_l[0] * torch.exp(res.lse_m[0])
            return res.afeat[0] * (s / (1.0 + s)).unsqueeze(-1), res.scores
        afeat, scores = diff[0][0], diff[2]
        s = torch.exp(scores).sum(dim=1)
        return afeat * (s / (1.0 + s)).unsqueeze(-1), scores

    def _logits(self, M):
        return torch.cat([self.classifiers[c](M[c]) for c in range(self.n_classes)]).reshape(1, self.n_classes).float()


# --------------------------------------------------------------------------------------------
class Attention_Gated(_AttentionGatedBase):
    """ibmil.py:7-35 -- same gate as transformer.py:239-267, raw [K, N] scores."""

    def __init__(self, L=512, D=128, K=1):
        super().__init__(L, D, K, norm_default=False)


class IBMIL(_GatedPoolModule):
    """ibmil.py:38-117.  forward(x [1, N, D_feat]) -> (Y_prob [1, C], M [1, D_inner (+ conf)], A [1, N] softmaxed)
    (the third value is ``deconf_A`` on the deconfounded path, like the reference)."""

    def __init__(self, conf, confounder_dim=128, confounder_merge='cat'):
        super().__init__()
        self.confounder_merge = confounder_merge
        assert confounder_merge in ['cat', 'add', 'sub']
        self.dimreduction = DimReduction(conf.D_feat, conf.D_inner)
        self.attention = Attention_Gated(conf.D_inner, 128, 1)
        self.classifier = Classifier_1fc(conf.D_inner, conf.n_class, 0)
        self.confounder_path = None
        if getattr(conf, "c_path", None):
            self.confounder_path = conf.c_path
            conf_tensor = torch.cat([torch.from_numpy(np.load(i)).view(-1, conf.D_inner).float() for i in conf.c_path], 0)
            conf_tensor_dim = conf_tensor.shape[-1]
            if conf.c_learn:
                self.confounder_feat = nn.Parameter(conf_tensor, requires_grad=True)
            else:
                self.register_buffer("confounder_feat", conf_tensor)
            self.W_q = nn.Linear(conf.D_inner, confounder_dim)
            self.W_k = nn.Linear(conf_tensor_dim, confounder_dim)
            if confounder_merge == 'cat':
                self.classifier = nn.Linear(conf.D_inner + conf_tensor_dim, conf.n_class)
            else:
                self.classifier = nn.Linear(conf.D_inner, conf.n_class)
            self.dropout = nn.Dropout(0.5)
        self._make_op(GatedPoolSpec(d_in=conf.D_feat, d_inner=conf.D_inner, n_branch=1, d_attn=128))

    def _weights(self):
        return dict(w1=self.dimreduction.fc1.weight, **self.attention._w())

    def forward(self, x):
        res, diff = self._pool(x[0])
        if res is not None:
            M, A = res.afeat[0], GatedPool.softmax_rows(res.scores)
        else:
            M, A = diff[0][0], F.softmax(diff[2], dim=1)
        if self.confounder_path:
            bag_q = self.W_q(M)
            conf_k = self.W_k(self.confounder_feat)
            deconf_A = torch.mm(conf_k, bag_q.transpose(0, 1))
            deconf_A = F.softmax(deconf_A / torch.sqrt(torch.tensor(conf_k.shape[1], dtype=torch.float32, device=M.device)), 0)
            conf_feats = torch.mm(deconf_A.transpose(0, 1), self.confounder_feat)
            if self.confounder_merge == 'cat':
                M = torch.cat((M, conf_feats), dim=1)
            elif self.confounder_merge == 'add':
                M = M + conf_feats
            else:
                M = M - conf_feats
            return self.classifier(M), M, deconf_A
        return self.classifier(M), M, A
