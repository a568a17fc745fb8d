"""Times the ViT-S/16 encoder forward. usage: python tests/cuda/vit_time.py [batch] [precise]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from acmil_b200.vit import vit_small
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precise = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
m = vit_small(False, False, None).cuda().eval()
m.precise = bool(precise)
xs = [torch.randn(B, 3, 224, 224, device="cuda") for _ in range(2)]
with torch.no_grad():
    for i in range(2):
        y = m(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    e0.record()
    for i in range(reps):
        y = m(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"ViT-S/16 batch {B} precise={precise}: {ms:.2f} ms  {B / ms * 1e3:.0f} patches/s  {9.2e9 * B / ms / 1e9:.1f} TFLOP/s  feat[0,:3] {y[0,:3].cpu().numpy()}")
