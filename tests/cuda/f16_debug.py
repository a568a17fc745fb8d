import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import acmil_b200._lib as L
from acmil_b200 import ACMIL_GA, Struct
torch.manual_seed(41)
m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).cuda().train()
g = torch.Generator().manual_seed(6)
x16 = torch.randn(3000 + 1 + 4097, 384, generator=g).half().cuda()
off = [0, 3000, 3001, 3001 + 4097]
rand = torch.rand(3, 5, 10, generator=g).cuda()
op = m._op
w = m._weights()
packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
for impl in (L.IMPL_FFMA, L.IMPL_UMMA):
    rs = []
    for x in (x16.float(), x16.float().clone(), x16):
        r = op.run(packed, x, off, n_masked=10, keep=[6, 0, 6], rand=rand, impl=impl)
        rs.append(r)
    for i in (1, 2):
        print("impl", impl, "vs", i, "masked equal", torch.equal(rs[0].masked_idx.sort(-1).values, rs[i].masked_idx.sort(-1).values),
              "topk equal", torch.equal(rs[0].topk_idx, rs[i].topk_idx),
              "afeat diff", float((rs[0].afeat - rs[i].afeat).abs().max()), "scores diff", float((rs[0].scores - rs[i].scores).abs().max()))
    print(rs[0].masked_idx[0, 0], rs[2].masked_idx[0, 0], rs[0].topk_idx[0, 0], rs[2].topk_idx[0, 0])
