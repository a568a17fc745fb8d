"""TEST / BASELINE INFRASTRUCTURE ONLY -- CPU port of the reference's ACMIL_GA forward written with
the same torch op sequence the reference executes (architecture/transformer.py:305-330,
network.py:49-57), so that timing it on the host cores measures what the reference's CPU path costs.

Parity status: PINNED through tests/test_oracle_golden.py::test_torch_port_matches_golden (same
reference-generated vectors as the numpy oracle).  The reference itself cannot travel to the GPU box
(/root/reference does not exist there), hence cpu_baseline.kind == "port".
"""
import torch
import torch.nn.functional as F


def acmil_ga_forward(p: dict, x: torch.Tensor, training: bool = False, n_masked_patch: int = 0,
                     mask_drop: float = 0.0, rand: torch.Tensor = None):
    """p: reference-named state_dict of torch tensors; x [1, N, D_feat] -> (sub, slide, A_out[1,K,N])."""
    h = F.relu(F.linear(x[0], p["dimreduction.fc1.weight"]))                                     # :306-307
    a_v = torch.tanh(F.linear(h, p["attention.attention_V.0.weight"], p["attention.attention_V.0.bias"]))
    a_u = torch.sigmoid(F.linear(h, p["attention.attention_U.0.weight"], p["attention.attention_U.0.bias"]))
    a = F.linear(a_v * a_u, p["attention.attention_weights.weight"], p["attention.attention_weights.bias"])
    a = torch.transpose(a, 1, 0)                                                                 # :264 -> [K, N]
    if n_masked_patch > 0 and training:                                                          # :311-320
        k, n = a.shape
        nm = min(n_masked_patch, n)
        _, idx = torch.topk(a, nm, dim=-1)
        r = torch.rand(*idx.shape) if rand is None else rand
        rsel = torch.argsort(r, dim=-1)[:, :int(nm * mask_drop)]
        masked = idx[torch.arange(k).unsqueeze(-1), rsel]
        keep = torch.ones(k, n)
        keep.scatter_(-1, masked, 0)
        a = a.masked_fill(keep == 0, -1e9)
    a_out = a
    pr = F.softmax(a, dim=1)                                                                     # :323
    afeat = torch.mm(pr, h)                                                                      # :324
    k = a.shape[0]
    sub = torch.stack([F.linear(afeat[i], p[f"classifier.{i}.fc.weight"], p[f"classifier.{i}.fc.bias"])
                       for i in range(k)], dim=0)                                                # :325-327
    bag_a = F.softmax(a_out, dim=1).mean(0, keepdim=True)                                        # :328
    bag_feat = torch.mm(bag_a, h)                                                                # :329
    slide = F.linear(bag_feat, p["Slide_classifier.fc.weight"], p["Slide_classifier.fc.bias"])
    return sub, slide, a_out.unsqueeze(0)


def random_state(d_feat=384, d_inner=128, d_attn=128, k=5, n_class=2, seed=0):
    """nn.Linear-style uniform init (same bounds as the reference's default initialisers)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i, bias=True):
        b = 1.0 / i ** 0.5
        w = (torch.rand(o, i, generator=g) * 2 - 1) * b
        return (w, (torch.rand(o, generator=g) * 2 - 1) * b) if bias else (w, None)

    p = {"dimreduction.fc1.weight": lin(d_inner, d_feat, False)[0]}
    for nm, (o, i) in {"attention.attention_V.0": (d_attn, d_inner), "attention.attention_U.0": (d_attn, d_inner),
                       "attention.attention_weights": (k, d_attn)}.items():
        p[nm + ".weight"], p[nm + ".bias"] = lin(o, i)
    for c in range(k):
        p[f"classifier.{c}.fc.weight"], p[f"classifier.{c}.fc.bias"] = lin(n_class, d_inner)
    p["Slide_classifier.fc.weight"], p["Slide_classifier.fc.bias"] = lin(n_class, d_inner)
    return p
