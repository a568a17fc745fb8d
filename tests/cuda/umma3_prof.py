"""Cycle counters of the v3 row pass (GP_UMMA_PROF build): python tests/cuda/umma3_prof.py [bags] [n_masked]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from acmil_b200 import ACMIL_GA, Struct, _lib as L
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nmask = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 50000
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
buf = (C.c_longlong * (148 * 20 * 8))()
lib.acmil_debug_umma3_prof.argtypes = [C.POINTER(C.c_longlong), C.c_int]
print("rc", lib.acmil_debug_umma3_prof(buf, 148 * 20 * 8))
a = np.array(buf[:], dtype=np.int64).reshape(148, 20, 8)
tiles = S * 196 / 74
print(f"tiles per CTA ~{tiles:.1f}; cycles per launch, mean over CTAs (per tile in brackets)")
def row(name, v):
    print(f"{name:28s} {v.mean():10.0f}  [{v.mean() / tiles:7.0f}]  max {v.max():10.0f}")
row("MMA1 wait dh_free", a[0::2, 1, 0]); row("MMA1 wait xop_full", a[0::2, 1, 1]); row("MMA1 total", a[0::2, 1, 7])
row("MMA2 wait h_full", a[0::2, 2, 0]); row("MMA2 wait d2_empty", a[0::2, 2, 1]); row("MMA2 total", a[0::2, 2, 7])
row("CVT wait_full_x", a[:, 4:8, 0]); row("CVT wait_xop_empty", a[:, 4:8, 1]); row("CVT wait_dh_full", a[:, 4:8, 2]); row("CVT epi1", a[:, 4:8, 3]); row("CVT total", a[:, 4:8, 7])
row("G wait_d2_full", a[:, 8:16, 0]); row("G wait_sc_empty", a[:, 8:16, 1]); row("G total", a[:, 8:16, 7])
for i, nme in [(2, "P wait_sc_full"), (3, "P softmax/cand"), (4, "P pool"), (5, "P flush")]:
    row(nme, a[:, 16:20, i])
row("P total", a[:, 16:20, 7])
