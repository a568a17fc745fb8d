"""Prints the cycle counters of a GP_UMMA_PROF build (ACMIL_NVCC_EXTRA=-DGP_UMMA_PROF=1 python -m acmil_b200.build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200 import ACMIL_GA, Struct, _lib as L
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nmask = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(os.environ.get('UMMA_PROF_ROWS', '50000'))
torch.manual_seed(0)
m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).cuda().eval()
x = torch.randn(S * n, 384, device="cuda")
op = m._op; op.impl = L.IMPL_UMMA
w = m._weights()
packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
off = [i * n for i in range(S + 1)]
for it in range(3):
    rec, ctx = op.partial(packed, x, off, n_masked=nmask)
torch.cuda.synchronize()
lib = L.load()
buf = (C.c_longlong * (148 * 88))()
lib.acmil_debug_umma_prof.argtypes = [C.POINTER(C.c_longlong), C.c_int]
print("rc", lib.acmil_debug_umma_prof(buf, 148 * 88))
import numpy as np
a = np.array(buf[:], dtype=np.int64).reshape(148, 88)
names = {0: "TMA wait_empty_x", 8: "MMA wait_w_ready", 9: "MMA idle: G1 blk d1_empty", 10: "MMA idle: G1 blk xop_full", 11: "MMA idle: G2 blk hop_full",
         12: "MMA idle: G2 blk d2_empty", 13: "MMA idle: G1 done", 14: "MMA idle: G2 done", 15: "MMA total", 16: "CVT wait_full_x", 17: "CVT wait_xop_empty", 23: "CVT total",
         24: "EPI wait_d1_full", 25: "EPI wait_d2_full", 26: "EPI epi1", 27: "EPI epi2(incl wait)", 28: "EPI softmax/cand",
         29: "EPI pool", 30: "EPI flush (bag ends)", 31: "EPI total"}
for cta in (0, 1):
    sel = a[cta::2]
    print(f"--- cta rank {cta} (mean over {len(sel)} CTAs; tiles per CTA ~{S * ((n + 255) // 256) / 74:.1f})")
    for k, v in names.items():
        print(f"{v:28s} mean {sel[:, k].mean():12.0f}  max {sel[:, k].max():12.0f}")

print(f"SM clock during the kernel: {a[:, 23].mean() / a[:, 22].mean() * 1e3:.0f} MHz  (converter loop: {a[:, 23].mean():.0f} cycles in {a[:, 22].mean() / 1e3:.1f} us)")
enames = ["wait_d1_full", "wait_d2_full", "epi1", "epi2(incl wait)", "softmax/cand", "pool", "flush", "total"]
print("--- epilogue warps (mean over all CTAs), cycles per launch")
print(" " * 18 + "".join(f"{'e' + str(e):>10s}" for e in range(8)))
for i, nme in enumerate(enames):
    print(f"{nme:18s}" + "".join(f"{a[:, 24 + 8 * e + i].mean():10.0f}" for e in range(8)))
