"""Times TransMIL forward (BASELINE config 3 dims) on the GPU. usage: python tests/cuda/transmil_time.py [n] [D] [precise]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200 import Struct
from acmil_b200.transmil import TransMIL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
precise = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0)
m = TransMIL(Struct(D_feat=D, D_inner=D, n_class=2)).cuda().eval()
for l in (m.layer1, m.layer2):
    l.attn.precise = bool(precise)
xs = [torch.randn(1, n, D, device="cuda") for _ in range(3)]
with torch.no_grad():
    for i in range(3):
        y = m(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for i in range(reps):
        y = m(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"TransMIL n={n} D={D} precise={precise}: {ms:.3f} ms/slide  {1e3/ms:.1f} slides/s  logits {y.cpu().numpy().ravel()}")
