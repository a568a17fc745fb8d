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
        r = torch.rand(*idx.shape) if rand is None else rand                                     # :316 (drawn on the host)
        rsel = torch.argsort(r, dim=-1)[:, :int(nm * mask_drop)]
        masked = idx[torch.arange(k).unsqueeze(-1), rsel.to(idx.device)]
        keep = torch.ones(k, n).to(a.device)                                                     # :318
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


# ------------------------------------------------------------------------------------------ TransMIL
def _nystrom_attention(p, pre, x, heads, m, iters=6):
    """architecture/nystrom_attention.py:67-143 (mask=None, eval), torch ops in the reference's order."""
    b, n, dim = x.shape
    d = p[pre + "to_qkv.weight"].shape[0] // 3 // heads
    rem = n % m
    if rem > 0:
        x = F.pad(x, (0, 0, m - rem, 0), value=0)                                                # :72-76
    q, k, v = F.linear(x, p[pre + "to_qkv.weight"]).chunk(3, dim=-1)                             # :83
    q, k, v = (t.reshape(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))                    # :84
    q = q * d ** -0.5                                                                            # :93
    l = -(-n // m)
    ql = q.reshape(b, heads, m, l, d).sum(3) / l                                                 # :98-114
    kl = k.reshape(b, heads, m, l, d).sum(3) / l
    a1 = torch.softmax(q @ kl.transpose(-1, -2), -1)                                             # :119-133
    a2 = torch.softmax(ql @ kl.transpose(-1, -2), -1)
    a3 = torch.softmax(ql @ k.transpose(-1, -2), -1)
    ax = a2.abs()                                                                                # :12-27
    z = a2.transpose(-1, -2) / (ax.sum(-1).max() * ax.sum(-2).max())
    eye = torch.eye(m).unsqueeze(0)
    for _ in range(iters):
        xz = a2 @ z
        z = 0.25 * z @ (13 * eye - (xz @ (15 * eye - (xz @ (7 * eye - xz)))))
    out = (a1 @ z) @ (a3 @ v)                                                                    # :135
    wc = p[pre + "res_conv.weight"]
    out = out + F.conv2d(v, wc, padding=(wc.shape[2] // 2, 0), groups=heads)                     # :137-138
    out = out.transpose(1, 2).reshape(b, -1, heads * d)                                          # :141
    out = F.linear(out, p[pre + "to_out.0.weight"], p[pre + "to_out.0.bias"])                    # :142
    return out[:, -n:]


def transmil_forward(p: dict, x: torch.Tensor):
    """architecture/transMIL.py:60-91, eval mode; p: reference-named state_dict; x [B, n, D_feat] -> logits."""
    h = F.relu(F.linear(x, p["_fc1.0.weight"], p["_fc1.0.bias"]))                                # :61
    n, dim = h.shape[1], h.shape[2]
    side = int(-(-n ** 0.5 // 1))
    while side * side < n:
        side += 1
    h = torch.cat([h, h[:, :side * side - n]], dim=1)                                            # :64-67
    h = torch.cat([p["cls_token"].expand(h.shape[0], -1, -1), h], dim=1)                         # :70-72
    for name in ("layer1.", "layer2."):
        if name == "layer2.":                                                                    # :78 PPEG between the layers
            cls, feat = h[:, :1], h[:, 1:]
            g = feat.transpose(1, 2).reshape(h.shape[0], dim, side, side)
            y = g
            for cn, kk in (("proj", 7), ("proj1", 5), ("proj2", 3)):
                y = y + F.conv2d(g, p[f"pos_layer.{cn}.weight"], p[f"pos_layer.{cn}.bias"], padding=kk // 2, groups=dim)
            h = torch.cat([cls, y.flatten(2).transpose(1, 2)], dim=1)
        xn = F.layer_norm(h, (dim,), p[name + "norm.weight"], p[name + "norm.bias"])             # :27
        h = h + _nystrom_attention(p, name + "attn.", xn, 8, dim // 2)
    h = F.layer_norm(h, (dim,), p["norm.weight"], p["norm.bias"])[:, 0]                          # :84
    return F.linear(h, p["_fc2.weight"], p["_fc2.bias"])                                         # :87
