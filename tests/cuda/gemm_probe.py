"""GEMM engine probe: accuracy of the precise modes and timing of representative shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200.transmil import gemm_nt
torch.manual_seed(0)
a = torch.randn(1024, 512, device="cuda"); b = torch.randn(512, 512, device="cuda")
ref = a.double() @ b.double().T
for mode in (1, 2, 0):
    out = gemm_nt(a, b, precise=mode)
    print("precise", mode, "max rel err", float((out.double() - ref).abs().max() / ref.abs().max()))
def timeit(m, n, k, batch=1, mode=1, reps=20):
    a = torch.randn(batch, m, k, device="cuda"); b = torch.randn(batch, n, k, device="cuda")
    out = torch.empty(batch, m, n, device="cuda")
    for _ in range(3): gemm_nt(a, b, precise=mode, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gemm_nt(a, b, precise=mode, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {batch}x[{m}x{n}x{k}] mode {mode}: {ms*1e3:8.1f} us  {2.0*batch*m*n*k/ms/1e9:7.1f} TFLOP/s")
for mode in (1, 2, 0):
    timeit(50432, 512, 512, 1, mode)
    timeit(51200, 1536, 384, 1, mode)
    timeit(256, 256, 256, 8, mode)
    timeit(256, 50432, 64, 8, mode, reps=5)
