"""GEMM engine timing on the ViT / TransMIL shapes: python tests/cuda/gemm_shapes.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200.transmil import gemm_nt
torch.manual_seed(0)
def timeit(m, n, k, batch=1, mode=1, reps=10, note=""):
    a = torch.randn(batch, m, k, device="cuda"); b = torch.randn(batch, n, k, device="cuda")
    out = torch.empty(batch, m, n, device="cuda")
    for _ in range(3): gemm_nt(a, b, precise=mode, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gemm_nt(a, b, precise=mode, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {batch}x[{m}x{n}x{k}] mode {mode}: {ms*1e3:8.1f} us  {2.0*batch*m*n*k/ms/1e9:7.1f} TFLOP/s  out {batch*m*n*4/ms/1e6:7.1f} GB/s {note}")
timeit(197, 197, 64, 1536, note="ViT scores")
timeit(256, 256, 64, 1536, note="no tails")
timeit(256, 256, 64, 1536, mode=0, note="no tails, plain tf32")
timeit(128, 128, 64, 6144, note="one tile per batch entry")
timeit(256, 197, 64, 1536, note="N tail only")
timeit(197, 256, 64, 1536, note="M tail only")
timeit(200, 200, 64, 1536, note="tails, multiple of 8")
timeit(197, 64, 200, 1536, note="ViT PV (k padded)")
timeit(50432, 1536, 384, note="fc1")
timeit(50432, 384, 1536, note="fc2")
timeit(50432, 768, 384, note="qk")
timeit(50432, 128, 64, note="thin")
