import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from acmil_b200 import gp_backward as B, _lib as L
from acmil_b200.gated_pool import GatedPool, GatedPoolSpec
torch.backends.cuda.matmul.allow_tf32 = False
d_in, Li, K, n = 384, 128, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 2999
spec = GatedPoolSpec(d_in=d_in, d_inner=Li, n_branch=K)
g = torch.Generator().manual_seed(1000 + n)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
w = dict(w1=rnd(Li, d_in, scale=d_in ** -0.5), wv=rnd(128, Li, scale=Li ** -0.5), bv=rnd(128, scale=0.1), wu=rnd(128, Li, scale=Li ** -0.5),
         bu=rnd(128, scale=0.1), ww=rnd(K, 128, scale=0.3), bw=rnd(K, scale=0.1))
x = rnd(n, d_in)
op = GatedPool(spec)
packed = op.pack(w["w1"], None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
res = op.run(packed, x, [0, n])
g_afeat, g_bag, g_scores = rnd(K, Li), rnd(1, Li), rnd(K, n, scale=1e-3)
dbg = {}
got = B.pool_backward(spec, x, w, res.scores, res.lse_m[0], res.lse_l[0], res.afeat[0], g_afeat, g_bag, g_scores, need_dx=True, _debug=dbg)
xr = x.clone().requires_grad_(True)
h = F.relu(F.linear(xr, w["w1"])); h.retain_grad()
z = torch.cat([F.linear(h, w["wv"], w["bv"]), F.linear(h, w["wu"], w["bu"])], 1); z.retain_grad()
a = torch.tanh(z[:, :128]) * torch.sigmoid(z[:, 128:])
s = F.linear(a, w["ww"], w["bw"]).t()
af = torch.softmax(s, 1) @ h
torch.autograd.backward([af, af.mean(0, keepdim=True), s], [g_afeat, g_bag, g_scores])
def rep(name, mine, ref):
    e = (mine - ref).abs()
    rows = e.max(1).values if e.dim() == 2 else e
    bad = torch.nonzero(rows > 1e-3 * ref.abs().max()).flatten()
    print(f"{name:6s} max err {float(e.max()):.3e} (ref max {float(ref.abs().max()):.3e}); bad rows: {bad[:12].tolist()} (count {len(bad)})")
rep("h", dbg["h"], h.detach()); rep("ht", dbg["ht"][:, :n].t(), h.detach()); rep("z", dbg["z"], z.detach()); rep("dz", dbg["dz"], z.grad)
rep("dzt", dbg["dzt"][:, :n].t(), z.grad); rep("dh", dbg["dh"], h.grad)
dz1_ref = h.grad * (h.detach() > 0)
rep("dz1", dbg["dz1"], dz1_ref); rep("dz1t", dbg["dz1t"][:, :n].t(), dz1_ref); rep("xt", dbg["xt"][:, :n].t(), x); rep("dx", got["x"], xr.grad)
