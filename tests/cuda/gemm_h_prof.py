"""Where the warps of the fp16-split GEMM kernels wait (needs a library built with -DTM_GEMM_PROF=1:
tests/cuda/build_variants.sh gprof "-DTM_GEMM_PROF=1"; ACMIL_B200_LIB_DIR=acmil_b200/lib_gprof python tests/cuda/gemm_h_prof.py)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from acmil_b200.transmil import SplitImage, gemm_nt

torch.manual_seed(0)
NAMES = ["producer: wait free stage", "producer total", "mma: wait operands", "mma: wait accumulator", "mma total",
         "converter: wait TMA", "converter total", "epilogue: wait accumulator", "epilogue total", "converter: tcgen05.wait::st"]
for m, n, k, note in ((50432, 384, 1536, "ViT fc2"), (50432, 1536, 384, "ViT fc1"), (50176, 512, 512, "TransMIL 512^2")):
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") * 0.02
    out = torch.empty(m, n, device="cuda")
    img = SplitImage(w)
    for _ in range(3):
        gemm_nt(a, w, out=out, b_split=img)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gemm_nt(a, w, out=out, b_split=img, _prof=prof)
    e1.record()
    torch.cuda.synchronize()
    pr = prof.cpu().tolist()
    nchunk = (k + 63) // 64
    print(f"[{m}x{n}x{k}] {note}: {e0.elapsed_time(e1)*1e3:.1f} us, pair mode env ACMIL_GEMM_PAIR={os.environ.get('ACMIL_GEMM_PAIR', 'default')}")
    for i, nm in enumerate(NAMES):
        print(f"    {nm:32s} {pr[i]:10d} cycles")
    if pr[4]:
        print(f"    => mma warp busy issuing {100 * (pr[4] - pr[2] - pr[3]) / pr[4]:.1f} %, per K=64 stage {pr[1] / max(1, nchunk):.0f} x tiles cycles (producer total / chunks per tile)")
