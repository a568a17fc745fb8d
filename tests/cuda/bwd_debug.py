import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from acmil_b200 import gp_backward as B, _lib as L
from acmil_b200.gated_pool import GatedPool, GatedPoolSpec
torch.backends.cuda.matmul.allow_tf32 = False

def run(d_in, Li, K, n, masked, impl, act_a="tanh", fb=False, biases=True, gated=True, gs_scale=1e-3):
    spec = GatedPoolSpec(d_in=d_in, d_inner=Li, n_branch=K, front_bias=fb, act_a=act_a, gated=gated, gate_bias=biases, score_bias=biases)
    g = torch.Generator().manual_seed(1000 + n)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
    w = dict(w1=rnd(Li, d_in, scale=d_in ** -0.5), b1=rnd(Li, scale=0.1) if fb else None, wv=rnd(128, Li, scale=Li ** -0.5),
             bv=rnd(128, scale=0.1) if biases else None, wu=rnd(128, Li, scale=Li ** -0.5) if gated else None,
             bu=rnd(128, scale=0.1) if (gated and biases) else None, ww=rnd(K, 128, scale=0.3), bw=rnd(K, scale=0.1) if biases else None)
    x = rnd(n, d_in)
    op = GatedPool(spec, impl)
    packed = op.pack(w["w1"], w["b1"], w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    n_masked, keep, rand = (10, [6], torch.rand(1, K, 10, generator=g).cuda()) if masked else (0, [0], None)
    res = op.run(packed, x, [0, n], n_masked=n_masked, keep=keep, rand=rand)
    g_afeat, g_bag, g_scores = rnd(K, Li), rnd(1, Li), rnd(K, n, scale=gs_scale)
    got = B.pool_backward(spec, x, {k: v for k, v in w.items() if v is not None}, res.scores, res.lse_m[0], res.lse_l[0], res.afeat[0], g_afeat, g_bag, g_scores, need_dx=True)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items() if v is not None}
    xr = x.clone().requires_grad_(True)
    h = F.relu(F.linear(xr, leaves["w1"], leaves.get("b1")))
    zv = F.linear(h, leaves["wv"], leaves.get("bv"))
    a = torch.tanh(zv) if act_a == "tanh" else (F.relu(zv) if act_a == "relu" else F.gelu(zv))
    if gated: a = a * torch.sigmoid(F.linear(h, leaves["wu"], leaves.get("bu")))
    s = F.linear(a, leaves["ww"], leaves.get("bw")).t()
    s = s.masked_fill(res.scores == -1e9, -1e9)
    p = torch.softmax(s, 1)
    af = p @ h
    names = list(leaves)
    ref = torch.autograd.grad([af, af.mean(0, keepdim=True), s], [xr] + [leaves[k] for k in names], [g_afeat, g_bag, g_scores])
    errs = {k: float((got[k] - r).abs().max()) / (float(r.abs().max()) + 1e-30) for k, r in zip(["x"] + names, ref)}
    lse = torch.logsumexp(s, 1)
    print(f"d_in {d_in} Li {Li} K {K} n {n} masked {masked} impl {impl} act {act_a}: afeat err {float((res.afeat[0]-af).abs().max()):.2e} "
          f"lse err {float((res.lse_m[0] + torch.log(res.lse_l[0]) - lse).abs().max()):.2e} n_masked_pos {int((res.scores == -1e9).sum())} pmax {float(p.max()):.3f}")
    print("   ", {k: f"{v:.1e}" for k, v in errs.items()})

for impl in (L.IMPL_FFMA, L.IMPL_UMMA):
    for masked in (False, True):
        run(384, 128, 5, 2999, masked, impl)
run(384, 128, 5, 3000, True, L.IMPL_AUTO)
run(384, 128, 5, 2999, True, L.IMPL_AUTO, gs_scale=0.0)
run(1024, 512, 1, 700, False, L.IMPL_AUTO, act_a="gelu", fb=True, biases=False)
run(1024, 512, 1, 700, False, L.IMPL_AUTO, act_a="tanh", fb=True, biases=False)
run(1024, 512, 1, 704, False, L.IMPL_AUTO, act_a="gelu", fb=True, biases=False)
