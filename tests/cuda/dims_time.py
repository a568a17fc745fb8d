"""Eval-mode pool throughput for the dimension table of Step3_WSI_classification_ACMIL.py:69-87 (N = 50 000, K = 5, 8 bags per call):
fused tcgen05 kernel where it applies, else GEMM-engine front + FFMA pool (AUTO) against the all-FFMA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import acmil_b200._lib as L
from acmil_b200.gated_pool import GatedPool, GatedPoolSpec
S, n, K = 8, 50000, 5
for d_in, Li in ((384, 128), (512, 256), (768, 384), (1024, 512)):
    spec = GatedPoolSpec(d_in=d_in, d_inner=Li, n_branch=K)
    g = torch.Generator().manual_seed(d_in)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
    w = (rnd(Li, d_in, scale=d_in ** -0.5), None, rnd(128, Li, scale=Li ** -0.5), rnd(128, scale=0.1), rnd(128, Li, scale=Li ** -0.5),
         rnd(128, scale=0.1), rnd(K, 128, scale=0.3), rnd(K, scale=0.1))
    xs = [torch.randn(S * n, d_in, device="cuda") for _ in range(2)]
    off = [i * n for i in range(S + 1)]
    for name, impl in (("auto", L.IMPL_AUTO), ("ffma", L.IMPL_FFMA)):
        if name == "ffma" and Li == 128 and False:
            continue
        op = GatedPool(spec, impl)
        packed = op.pack(*w)
        for i in range(2):
            op.run(packed, xs[i % 2], off)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6 if name == "auto" else 2
        e0.record()
        for i in range(reps):
            op.run(packed, xs[i % 2], off)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gb = S * n * d_in * 4 / 1e9
        print(f"D_feat {d_in:4d} D_inner {Li:3d} {name}: {ms:8.3f} ms per {S} bags = {S / ms * 1e3:8.0f} slides/s, {gb / ms * 1e3:6.0f} GB/s of x "
              f"({gb / ms * 1e3 / 6553 * 100:4.1f} % of the HBM peak)", flush=True)
