"""fp16-split (pre-split weight image) against 3xTF32 on the weight products of ViT-S/16, TransMIL and ResNet18:
python tests/cuda/gemm_split_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from acmil_b200.transmil import SplitImage, gemm_nt

torch.manual_seed(0)


def timeit(m, n, k, note="", reps=10, gelu=False):
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") * 0.02
    out = torch.empty(m, n, device="cuda")
    img = SplitImage(w)
    res = []
    for split in (None, img):
        for _ in range(3):
            gemm_nt(a, w, out=out, b_split=split, gelu=gelu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gemm_nt(a, w, out=out, b_split=split, gelu=gelu)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps)
    ref = a.double() @ w.double().T
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    tf = 2.0 * m * n * k / 1e9
    print(f"  [{m}x{n}x{k}] 3xTF32 {res[0]*1e3:7.1f} us ({tf/res[0]:6.1f} TFLOP/s)   fp16-split {res[1]*1e3:7.1f} us ({tf/res[1]:6.1f} TFLOP/s)"
          f"   x{res[0]/res[1]:.2f}  max err {err:.1e}  {note}")


timeit(50432, 1536, 384, "ViT fc1", gelu=True)
timeit(50432, 384, 1536, "ViT fc2")
timeit(50432, 768, 384, "ViT q,k")
timeit(50432, 384, 384, "ViT proj")
timeit(50176, 384, 768, "ViT patch embedding")
timeit(50176, 512, 512, "TransMIL q (one of 3), to_out")
timeit(50000, 512, 384, "TransMIL fc1")
timeit(64 * 56 * 56, 64, 576, "ResNet18 layer1 3x3")
timeit(64 * 28 * 28, 128, 1152, "ResNet18 layer2 3x3")
timeit(64 * 7 * 7, 512, 4608, "ResNet18 layer4 3x3")
