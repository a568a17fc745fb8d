"""Bring-up driver for the tcgen05 kernel: compares IMPL_UMMA with IMPL_FFMA on the same device.
usage: python tests/cuda/umma_debug.py N [K] [n_masked]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200 import ACMIL_GA, Struct, _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nmask = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.manual_seed(0)
m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=K), n_token=K, n_masked_patch=nmask, mask_drop=0.6).cuda().eval()
x = torch.randn(n, 384, generator=torch.Generator().manual_seed(1)).cuda()
nm = min(nmask, n); keep = int(nm * 0.6)
rsel = torch.argsort(torch.rand(K, max(nm, 1), generator=torch.Generator().manual_seed(2)), dim=-1)[:, :keep].cuda() if keep else None
branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
out = {}
with torch.no_grad():
    for name, impl in (("ffma", L.IMPL_FFMA), ("umma", L.IMPL_UMMA)):
        m._op.impl = impl
        res, _ = m._pool(x, n_masked=nmask if keep else 0, keep=keep, rsel=rsel, branch=branch, head=head, slide_head=True)
        torch.cuda.synchronize()
        out[name] = res
        print(name, "slide", res.slide.cpu().numpy().ravel(), "lse_m", res.lse_m.cpu().numpy().ravel()[:3], flush=True)
a, b = out["ffma"], out["umma"]
d = (a.scores - b.scores).abs()
print("scores max abs diff", float(d.max()), "at", int(d.argmax()) % n, "ref absmax", float(a.scores.abs().max()))
print("afeat max abs diff", float((a.afeat - b.afeat).abs().max()), "slide diff", float((a.slide - b.slide).abs().max()),
      "sub diff", float((a.sub - b.sub).abs().max()))
if keep:
    print("masked equal", bool(torch.equal(a.masked_idx.sort(-1).values, b.masked_idx.sort(-1).values)),
          "topk equal", bool(torch.equal(a.topk_idx, b.topk_idx)))
bad = float(d.max()) > 1e-4 or not torch.isfinite(b.slide).all()
print("RESULT", "FAIL" if bad else "PASS")
sys.exit(1 if bad else 0)
