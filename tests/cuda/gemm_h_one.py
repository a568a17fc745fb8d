"""One weight product on the fp16-split kernel (for ncu): python tests/cuda/gemm_h_one.py [m n k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from acmil_b200.transmil import SplitImage, gemm_nt

m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (50432, 384, 1536)
torch.manual_seed(0)
a = torch.randn(m, k, device="cuda")
w = torch.randn(n, k, device="cuda") * 0.02
out = torch.empty(m, n, device="cuda")
img = SplitImage(w)
for _ in range(6):
    gemm_nt(a, w, out=out, b_split=img)
torch.cuda.synchronize()
