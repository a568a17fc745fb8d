"""Times the row-pass (+ reduce) of the tcgen05 path with CUDA events: python tests/cuda/umma_time.py [bags] [n_masked]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from acmil_b200 import ACMIL_GA, Struct, _lib as L
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 50000
torch.manual_seed(0)
m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).cuda().eval()
xs = [torch.randn(S * n, 384, device="cuda") for _ in range(3)]
op = m._op; op.impl = L.IMPL_UMMA
w = m._weights()
packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
off = [i * n for i in range(S + 1)]
lib = L.load()
for nmask in [int(a) for a in sys.argv[2:]] or [0, 10]:
    for it in range(5):
        op.partial(packed, xs[it % 3], off, n_masked=nmask)
    torch.cuda.synchronize()
    lib.acmil_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(20):
        op.partial(packed, xs[it % 3], off, n_masked=nmask)
    e1.record()
    torch.cuda.synchronize()
    ms, nl = C.c_double(0), C.c_int64(0)
    lib.acmil_prof_collect(C.byref(ms), C.byref(nl))
    lib.acmil_prof_enable(0)
    print(f"n_masked={nmask}: row pass {ms.value / nl.value * 1e3:.1f} us, partial() {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
