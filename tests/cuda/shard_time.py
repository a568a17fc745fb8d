"""Times the pieces of a bag-sharded step at the shape one rank sees at N ranks (S = 16 N bags of 50000 / N rows), with the
peers' gather buffers on the same GPU: python tests/cuda/shard_time.py [ranks]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from acmil_b200 import ACMIL_GA, Struct, _lib as L
ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S, n = 16 * ranks, 50000 // ranks
torch.manual_seed(0)
m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).cuda().train()
op = m._op
w = m._weights()
packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
branch = (torch.stack([c.fc.weight for c in m.classifier]).detach(), torch.stack([c.fc.bias for c in m.classifier]).detach())
head = (m.Slide_classifier.fc.weight.detach(), m.Slide_classifier.fc.bias.detach())
xs = [torch.randn(S * n, 384, device="cuda") for _ in range(3)]
off = [i * n for i in range(S + 1)]
cap = L.record_floats(5, 128, 10) * 4 * S
bufs = [torch.zeros((256 + 2 * ranks * cap) // 4, device="cuda") for _ in range(ranks)]


class X:
    def __init__(self, r):
        self.world, self.rank, self.state = ranks, r, torch.zeros(8, dtype=torch.int32, device="cuda")

    def c_struct(self, pb):
        x = L.GpExchange()
        x.n_ranks, x.rank = ranks, self.rank
        for r, b in enumerate(bufs):
            x.d_flags[r] = b.data_ptr()
            x.d_gather[r] = b.data_ptr() + 256
        x.d_epoch, x.d_ticket, x.gather_bytes = self.state.data_ptr(), self.state.data_ptr() + 16, 2 * ranks * cap
        return x


xch = [X(r) for r in range(ranks)]
rand = torch.rand(S, 5, 10, device="cuda")
fkw = dict(keep=[6] * S, rand=rand, branch_w=branch[0], branch_b=branch[1], head_w=head[0], head_b=head[1], slide_head=True)
lib = L.load()
ev = lambda: torch.cuda.Event(enable_timing=True)
tp = tf = 0.0
with torch.no_grad():
    for it in range(8):
        ctxs = []
        e0, e1, e2 = ev(), ev(), ev()
        for r in range(ranks):
            if r == 0:
                e0.record()
            _, ctx = op.partial(packed, xs[it % 3], off, n_masked=10, shard_begin=[r * n] * S, exchange=xch[r], want_scores=(r == 0))
            if r == 0:
                e1.record()
            ctxs.append(ctx)
        outs = []
        for r in range(ranks):
            if r == 0:
                e2.record()
            outs.append(op.finish(ctxs[r], None, ranks, **fkw))
            if r == 0:
                e3 = ev(); e3.record()
        torch.cuda.synchronize()
        if it >= 3:
            tp += e0.elapsed_time(e1); tf += e2.elapsed_time(e3)
print(f"ranks={ranks}: S={S} bags x {n} rows: partial_x (memset + row pass + rescue + 2 reduce, eager launches) {tp / 5 * 1e3:.1f} us, "
      f"finish {tf / 5 * 1e3:.1f} us")
