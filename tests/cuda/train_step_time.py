"""Times one training step (forward + loss + backward + AdamW) of ACMIL_GA at N = 50 000 per bag on the GPU:
kernel backward (acmil_b200.gp_backward), torch-op recompute backward (ACMIL_POOL_BACKWARD=torch), and the reference's op
sequence in stock eager PyTorch (oracle/torch_port.py is not used here: plain nn ops below)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from acmil_b200 import ACMIL_GA, Struct

torch.backends.cuda.matmul.allow_tf32 = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dev = "cuda"
xs = [torch.randn(1, n, 384, device=dev) for _ in range(4)]
y = torch.tensor([1], device=dev)


def loss_fn(sub, slide, a):
    p = torch.softmax(a, dim=-1)
    d = sum(torch.cosine_similarity(p[:, i], p[:, j], dim=-1).mean() for i in range(5) for j in range(i + 1, 5)) / 10
    return F.cross_entropy(sub, y.repeat_interleave(5)) + F.cross_entropy(slide, y) + d


def fused_loss(sub, slide, a):      # the same loss with the diversity term as one kernel each way (acmil_b200.losses)
    from acmil_b200.losses import diversity_loss
    return F.cross_entropy(sub, y.repeat_interleave(5)) + F.cross_entropy(slide, y) + diversity_loss(a)


def eager_forward(m, x):      # transformer.py:305-330 with torch ops
    h = F.relu(F.linear(x[0], m.dimreduction.fc1.weight))
    g = m.attention
    a = F.linear(torch.tanh(g.attention_V[0](h)) * torch.sigmoid(g.attention_U[0](h)), g.attention_weights.weight,
                 g.attention_weights.bias).t()
    k, nn_ = a.shape
    _, idx = torch.topk(a, 10, dim=-1)
    rand = torch.argsort(torch.rand(k, 10, device=a.device), dim=-1)[:, :6]
    mi = idx[torch.arange(k, device=a.device).unsqueeze(-1), rand]
    mask = torch.ones(k, nn_, device=a.device)
    mask.scatter_(-1, mi, 0)
    a = a.masked_fill(mask == 0, -1e9)
    ao = a
    p = F.softmax(a, dim=1)
    af = p @ h
    sub = torch.stack([c.fc(af[i]) for i, c in enumerate(m.classifier)])
    bag = torch.mm(F.softmax(ao, dim=1).mean(0, keepdim=True), h)
    return sub, m.Slide_classifier.fc(bag), ao.unsqueeze(0)


def measure(mode, graph):
    os.environ["ACMIL_POOL_BACKWARD"] = "torch" if mode == "torch" else "kernel"
    torch.manual_seed(0)
    m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, capturable=graph)
    xin = torch.empty_like(xs[0])

    def step():
        opt.zero_grad(set_to_none=True)
        out = eager_forward(m, xin) if mode == "eager" else m(xin)
        loss = fused_loss(*out) if mode == "fused" else loss_fn(*out)
        loss.backward()
        opt.step()
        return loss

    g = None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                xin.copy_(xs[i % 4])
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            loss = step()
    run = g.replay if graph else step
    for i in range(4):
        xin.copy_(xs[i % 4])
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        xin.copy_(xs[i % 4])
        run()
    e1.record()
    torch.cuda.synchronize()
    w = m.dimreduction.fc1.weight
    print(f"{mode:7s} {'graph' if graph else 'eager-launch':12s}: {e0.elapsed_time(e1) / 12:.3f} ms per training step (N = {n}); "
          f"|W1| after {float(w.abs().sum()):.4f} finite {bool(torch.isfinite(w).all())}", flush=True)


only = sys.argv[2].split(",") if len(sys.argv) > 2 else None      # e.g. "kernel:eager" = kernel backward, eager launches
for graph in (False, True):
    for mode in ("fused", "kernel", "torch", "eager"):
        if only and f"{mode}:{'graph' if graph else 'eager'}" not in only:
            continue
        try:
            measure(mode, graph)
        except Exception as exc:      # noqa: BLE001
            print(f"{mode:7s} {'graph' if graph else 'eager-launch':12s}: FAILED {type(exc).__name__}: {str(exc)[:300]}", flush=True)
            torch.cuda.synchronize()
