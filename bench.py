#!/usr/bin/env python
"""Benchmark of the ACMIL gated-attention pool hot path (BASELINE.json metric: slides/sec, ACMIL ga,
N=50k, D=384; HBM GB/s on the fused pool kernel).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the CPU port of the reference
                                                             # forward on the box's host cores

One "step" = one pass of the hot path over one batch of `--slides` synthetic bags per GPU
(config 2 of BASELINE.json: n_token=5, n_masked_patch=10, mask_drop=0.6, train-mode forward so that the
stochastic top-k masking is inside the step).  At N GPUs every bag is row-sharded over the N ranks (each
rank owns N_rows/N patches of each bag; the only exchange is the all-gather of the per-bag partial
records) and a step processes N * slides bags, so per-GPU work is fixed: "scaling": "weak".

Timing: W >= 3 warm-up steps, then K steps bracketed by barrier + cuda.synchronize, CUDA events on the
launch stream, max over ranks.  Consecutive steps read different device-resident bag groups, each group
(slides * 76.8 MB) larger than the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, D_FEAT, D_INNER, K_BRANCH, N_CLASS, N_MASKED, MASK_DROP = 50000, 384, 128, 5, 2, 10, 0.6
# x 76.8 MB + A_out 1.0 MB + params 0.337 MB per bag
ALGO_BYTES_PER_SLIDE = N_ROWS * D_FEAT * 4 + K_BRANCH * N_ROWS * 4 + 84369 * 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slides", type=int, default=16, help="bags per step per GPU")
    ap.add_argument("--groups", type=int, default=3, help="distinct device-resident bag groups rotated over steps")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ffma", "umma"])
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="graph: the step (mask draw + row pass + reduce + finish) is captured once per bag group in a CUDA "
                         "graph and replayed (falls back to eager launches if capture fails)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp16", action="store_true", help="skip the fp16-feature variant of the step (N = 1 only)")
    ap.add_argument("--no-train-step", action="store_true", help="skip the training-step comparison (N = 1 only)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the stock-PyTorch-on-this-GPU comparator (N = 1 only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the per-bag partial records travel: inside the kernels over peer memory, or NCCL")
    ap.add_argument("--workload", default="acmil", choices=["acmil", "transmil", "vit", "resnet", "stream"],
                    help="acmil = the headline metric (BASELINE.json configs[1]); transmil = configs[2], see bench_transmil.py")
    ap.add_argument("--patch-batch", type=int, default=256, help="vit / resnet: patches per step per GPU")
    ap.add_argument("--stream-slides", type=int, default=398, help="stream: slides in the Camelyon16-shaped set")
    ap.add_argument("--stream-scale", type=float, default=1 / 64, help="stream: patch-count scale (1 = 10k..100k patches per slide)")
    ap.add_argument("--transmil-replicas", action="store_true", help="transmil at N > 1: independent replicas instead of one sharded bag")
    ap.add_argument("--dim", type=int, default=512, help="transmil: D_inner")
    ap.add_argument("--d-feat", type=int, default=512, help="transmil: D_feat")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_port_rate(n_rows, slides, warmup, mode):
    """slides/sec of the CPU port of the reference forward (oracle/torch_port.py) on all host cores."""
    from oracle import torch_port as T
    torch.set_num_threads(os.cpu_count() or 1)
    p = T.random_state(D_FEAT, D_INNER, 128, K_BRANCH, N_CLASS, seed=0)
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(1, n_rows, D_FEAT, generator=g) for _ in range(2)]
    times = []
    with torch.no_grad():
        for i in range(warmup + slides):
            t0 = time.perf_counter()
            T.acmil_ga_forward(p, xs[i % 2], mode == "train", N_MASKED, MASK_DROP)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return len(times) / sum(times), sum(times) / len(times)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, sec = cpu_port_rate(a.rows, a.steps, max(a.warmup, 1), a.mode)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "slides/sec (ACMIL ga, N=50k, D=384)", "value": rate, "unit": "slides/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, 1, cpu=True),
        "cpu_baseline": {"value": rate, "unit": "slides/s", "cores": cores, "kind": "port",
                         "sample": f"{a.steps} bags of {a.rows}x{D_FEAT} fp32, one bag per step, torch CPU ops in the "
                                   f"reference's op order (oracle/torch_port.py)"},
        "e2e": {"value": rate, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(a, world, cpu=False):
    return {"workload": f"ACMIL ga n_token={K_BRANCH} n_masked_patch={N_MASKED} mask_drop={MASK_DROP} "
                        f"D_feat={D_FEAT} D_inner={D_INNER}, synthetic N(0,1) fp32 bags of N={a.rows} rows "
                        f"(BASELINE.json configs[1]), {a.mode}-mode forward",
            "bags_per_step": 1 if cpu else a.slides * world, "rows_per_bag": a.rows,
            "parallelism": "cpu" if cpu else (f"bag rows sharded over {world} GPUs" if world > 1 else "1 GPU"),
            "l2_policy": "each step reads a different resident bag group; one group > 126 MB L2"}


# ------------------------------------------------------------------------------------------ our arm
def gpu_eager_rate(dev, n_rows, mode, steps=10, warmup=3):
    """slides/sec of the reference's op sequence (oracle/torch_port.py: the same torch calls as transformer.py:305-330)
    run by stock PyTorch ON THE B200, fp32 with TF32 off, device-resident bags, CUDA-event timed: what a user of the
    reference gets on this GPU today (BASELINE.md section 4 / SURVEY.md section 8d)."""
    from oracle import torch_port as T
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        p = {k: v.to(dev) for k, v in T.random_state(D_FEAT, D_INNER, 128, K_BRANCH, N_CLASS, seed=0).items()}
        g = torch.Generator(device=dev).manual_seed(0)
        xs = [torch.randn(1, n_rows, D_FEAT, generator=g, device=dev) for _ in range(3)]      # 3 x 76.8 MB > L2
        with torch.no_grad():
            for i in range(warmup):
                T.acmil_ga_forward(p, xs[i % 3], mode == "train", N_MASKED, MASK_DROP)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                T.acmil_ga_forward(p, xs[i % 3], mode == "train", N_MASKED, MASK_DROP)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"value": 1e3 / ms, "unit": "slides/s", "ms_per_slide": ms, "kind": "stock PyTorch eager on this GPU, fp32 (TF32 off)",
                "sample": f"{steps} bags of {n_rows}x{D_FEAT} fp32, device-resident, one bag per call, after {warmup} warm-ups "
                          f"(oracle/torch_port.py = the reference's op sequence)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def train_step_rates(dev, n_rows, steps=12, warmup=4):
    """One TRAINING step per bag (Step3_WSI_classification_ACMIL.py:188-221: forward, sub/slide cross-entropy + branch-diversity
    loss, backward, AdamW) at N = n_rows, CUDA-event timed: ours (fused forward kernels, kernel backward of
    acmil_b200.gp_backward, acmil_b200.losses.diversity_loss, the whole step replayed from one CUDA graph) next to the
    reference's own op sequence in stock eager PyTorch on this GPU (fp32, TF32 off)."""
    import torch.nn.functional as F
    from acmil_b200 import ACMIL_GA, Struct
    from acmil_b200.losses import diversity_loss
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(5)
    xs = [torch.randn(1, n_rows, D_FEAT, generator=g, device=dev) for _ in range(3)]
    y = torch.tensor([1], device=dev)

    def ref_loss(sub, slide, a):       # Step3_WSI_classification_ACMIL.py:201-216
        p = torch.softmax(a, dim=-1)
        d = sum(torch.cosine_similarity(p[:, i], p[:, j], dim=-1).mean() for i in range(K_BRANCH) for j in range(i + 1, K_BRANCH))
        return F.cross_entropy(sub, y.repeat_interleave(K_BRANCH)) + F.cross_entropy(slide, y) + d / (K_BRANCH * (K_BRANCH - 1) / 2)

    def ref_forward(m, x):             # transformer.py:305-330 with torch ops on the module's own parameters
        h = F.relu(F.linear(x[0], m.dimreduction.fc1.weight))
        gt = m.attention
        a = F.linear(torch.tanh(gt.attention_V[0](h)) * torch.sigmoid(gt.attention_U[0](h)), gt.attention_weights.weight,
                     gt.attention_weights.bias).t()
        k, n = a.shape
        _, idx = torch.topk(a, N_MASKED, dim=-1)
        rsel = torch.argsort(torch.rand(k, N_MASKED, device=a.device), dim=-1)[:, :int(N_MASKED * MASK_DROP)]
        mi = idx[torch.arange(k, device=a.device).unsqueeze(-1), rsel]
        mask = torch.ones(k, n, device=a.device)
        mask.scatter_(-1, mi, 0)
        a = a.masked_fill(mask == 0, -1e9)
        af = F.softmax(a, dim=1) @ h
        sub = torch.stack([c.fc(af[i]) for i, c in enumerate(m.classifier)])
        bag = torch.mm(F.softmax(a, dim=1).mean(0, keepdim=True), h)
        return sub, m.Slide_classifier.fc(bag), a.unsqueeze(0)

    def measure(ours):
        torch.manual_seed(0)
        conf = Struct(D_feat=D_FEAT, D_inner=D_INNER, n_class=N_CLASS, n_token=K_BRANCH)
        m = ACMIL_GA(conf, n_token=K_BRANCH, n_masked_patch=N_MASKED, mask_drop=MASK_DROP).to(dev).train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, capturable=ours)
        xin = torch.empty_like(xs[0])

        def step():
            opt.zero_grad(set_to_none=True)
            if ours:
                sub, slide, a = m(xin)
                loss = F.cross_entropy(sub, y.repeat_interleave(K_BRANCH)) + F.cross_entropy(slide, y) + diversity_loss(a)
            else:
                loss = ref_loss(*ref_forward(m, xin))
            loss.backward()
            opt.step()

        run, launch = step, "eager launches"
        if ours:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    xin.copy_(xs[i % 3])
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                step()
            run, launch = graph.replay, "one CUDA graph per step, replayed"
        for i in range(warmup):
            xin.copy_(xs[i % 3])
            run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            xin.copy_(xs[i % 3])      # a fresh bag every step (device copy, inside the timed region for both arms)
            run()
        e1.record()
        torch.cuda.synchronize(dev)
        ok = all(bool(torch.isfinite(p).all()) for p in m.parameters())
        return e0.elapsed_time(e1) / steps, launch, ok

    try:
        ms_ours, launch, ok = measure(True)
        ms_ref, _, _ = measure(False)
        return {"ours_ms_per_step": ms_ours, "stock_pytorch_ms_per_step": ms_ref, "speedup": ms_ref / ms_ours,
                "launch": launch, "weights_finite": ok,
                "what": f"forward + CE/diversity loss + backward + AdamW on one bag of {n_rows}x{D_FEAT} fp32 per step "
                        f"(Step3_WSI_classification_ACMIL.py:188-221), masking on; stock = the reference's op sequence in eager "
                        f"PyTorch on this GPU, TF32 off"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_ours(a):
    import torch.distributed as dist
    from acmil_b200 import ACMIL_GA, Struct, _lib
    from acmil_b200.sharding import PeerExchange, ShardedACMIL, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    impl = {"auto": _lib.IMPL_AUTO, "ffma": _lib.IMPL_FFMA, "umma": _lib.IMPL_UMMA}[a.kernel]

    torch.manual_seed(0)
    conf = Struct(D_feat=D_FEAT, D_inner=D_INNER, n_class=N_CLASS, n_token=K_BRANCH)
    model = ACMIL_GA(conf, n_token=K_BRANCH, n_masked_patch=N_MASKED, mask_drop=MASK_DROP).to(dev)
    model.train(a.mode == "train")
    op = model._op
    op.impl = impl
    w = model._weights()
    packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    branch_w = torch.stack([c.fc.weight for c in model.classifier]).detach()
    branch_b = torch.stack([c.fc.bias for c in model.classifier]).detach()
    head_w, head_b = model.Slide_classifier.fc.weight.detach(), model.Slide_classifier.fc.bias.detach()

    S = a.slides * world                      # bags per step (each sharded over `world` ranks)
    b = shard_bounds(a.rows, world)
    n_loc = b[rank + 1] - b[rank]
    offsets = [i * n_loc for i in range(S + 1)]
    shard_begin = [b[rank]] * S
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    groups = [torch.randn(S * n_loc, D_FEAT, generator=gen, device=dev) for _ in range(a.groups)]
    masking = a.mode == "train"
    nm = min(N_MASKED, a.rows)
    keep = int(nm * MASK_DROP)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the exchange of the sharded bags' partial records: inside the kernels over peer memory (NVLink stores + flags,
    # include/acmil_b200.h acmil_gp_exchange), or -- if symmetric memory cannot be set up on this box -- an NCCL all-gather
    exchange, exchange_kind = None, "none (1 GPU)"
    if world > 1:
        exchange_kind = "NCCL all_gather_into_tensor of the records (torch.distributed)"
        if a.exchange == "peer":
            try:
                part_bytes = _lib.record_floats(K_BRANCH, D_INNER, N_MASKED) * 4 * S
                exchange = PeerExchange(part_bytes, dev)
                exchange_kind = ("in-kernel: the reduce kernel stores the records into every peer's buffer over NVLink and raises "
                                 "a flag, the finish kernel waits for the flags (symmetric memory; no collective per step)")
            except Exception as exc:      # no P2P / symmetric memory on this box
                print(f"[rank {rank}] peer exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                exchange = None
            ok = torch.tensor([1.0 if exchange is not None else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 0.5:
                exchange = None

    rng_in_sync = [world == 1]

    def step(i):
        r = None
        if masking:
            # per bag: the draw of transformer.py:316; its argsort is taken inside the finish kernel.  Every rank seeds its
            # generator identically (torch.manual_seed(0) above) and makes the same calls, so all ranks draw the same
            # numbers without an exchange; `rng_in_sync` checks that once and falls back to a broadcast otherwise.
            r = torch.rand(S, K_BRANCH, nm, device=dev)
            if world > 1 and not rng_in_sync[0]:
                dist.broadcast(r, src=0)
        return op.run(packed, groups[i % a.groups], offsets, n_masked=N_MASKED if masking else 0,
                      keep=[keep if masking else 0] * S, rand=r, branch_w=branch_w, branch_b=branch_b,
                      head_w=head_w, head_b=head_b, slide_head=True, shard_begin=shard_begin if world > 1 else None,
                      group=dist.group.WORLD if (world > 1 and exchange is None) else None, exchange=exchange)

    if world > 1:      # do the ranks' default CUDA generators agree?  (compare one probe draw with rank 0's)
        probe = torch.rand(64, device=dev)
        ref0 = probe.clone()
        dist.broadcast(ref0, src=0)
        same = torch.tensor([1.0 if torch.equal(probe, ref0) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rng_in_sync[0] = bool(same.item() > 0.5)

    # ---- sharded-versus-unsharded parity on real hardware: the same seeded bags once row-sharded over the ranks (the
    # path timed below) and once whole on every rank; logits within 1e-3, mask indices identical
    parity = None
    if world > 1:
        with torch.no_grad():
            Sc = min(S, 4)
            gfull = torch.Generator(device=dev).manual_seed(99)
            xfull = torch.randn(Sc * a.rows, D_FEAT, generator=gfull, device=dev)          # same on every rank
            rnd = torch.rand(Sc, K_BRANCH, nm, generator=gfull, device=dev)
            xloc = torch.cat([xfull[s * a.rows + b[rank]: s * a.rows + b[rank + 1]] for s in range(Sc)])
            kw = dict(n_masked=N_MASKED if masking else 0, keep=[keep if masking else 0] * Sc, rand=rnd if masking else None,
                      branch_w=branch_w, branch_b=branch_b, head_w=head_w, head_b=head_b, slide_head=True)
            whole = op.run(packed, xfull, [i * a.rows for i in range(Sc + 1)], **kw)
            # the parity step uses its own (smaller) batch; a PeerExchange is sized for the timed batch, which is larger
            shard = op.run(packed, xloc, [i * n_loc for i in range(Sc + 1)], shard_begin=[b[rank]] * Sc,
                           group=dist.group.WORLD if exchange is None else None, exchange=exchange, **kw)
            err = float(((shard.slide - whole.slide).abs() / (whole.slide.abs() + 1e-5)).max())
            err = max(err, float(((shard.sub - whole.sub).abs() / (whole.sub.abs() + 1e-5)).max()))
            mask_ok = True
            if masking:
                mask_ok = bool(torch.equal(torch.sort(shard.masked_idx, -1).values, torch.sort(whole.masked_idx, -1).values))
            loc = torch.cat([whole.scores[:, s * a.rows + b[rank]: s * a.rows + b[rank + 1]] for s in range(Sc)], dim=1)
            sc_err = float(((shard.scores - loc).abs() / (loc.abs() + 1e-5)).clamp(max=1.0).max())
            t = torch.tensor([err, sc_err, 0.0 if mask_ok else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity = {"parity_ok": bool(t[0] < 1e-3 and t[1] < 1e-3 and t[2] == 0), "logits_max_rel_err": float(t[0]),
                      "scores_max_rel_err": float(t[1]), "mask_indices_identical": bool(t[2] == 0), "bags": Sc,
                      "what": "row-sharded over the ranks (this run's exchange) versus the whole bags on one GPU, same seeds"}

    with torch.no_grad():
        # NCCL opens its channels lazily over the first collectives of each size (peer-memory exchange: none): a few
        # untimed communication pre-warm steps, reported separately from the W warm-up steps the command line asked for
        comm_prewarm = 8 if (world > 1 and exchange is None) else 0
        for i in range(comm_prewarm):
            step(i)
        for i in range(a.warmup):
            step(i)
        barrier()
        l0 = _lib.launch_count()
        step(0)
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count() - l0

        # the step is launch-bound next to its row pass (small reduce / finish launches): capture it once per bag group.
        # With the in-kernel exchange a multi-GPU step is plain kernel launches too and is captured the same way.
        graphs, graph_res, launch_mode = None, None, "eager"
        if a.launch == "graph" and world > 1 and exchange is None and os.environ.get("ACMIL_BENCH_NCCL_GRAPH", "0") != "1":
            launch_mode = "eager (the step holds NCCL collectives; graph replay needs the peer-memory exchange)"
        elif a.launch == "graph":
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for gi in range(a.groups):
                        step(gi)
                torch.cuda.current_stream().wait_stream(side)
                barrier()
                graphs, graph_res = [], []
                for gi in range(a.groups):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        r = step(gi)
                    graphs.append(g)
                    graph_res.append(r)
                barrier()
                for gi in range(a.groups):
                    graphs[gi].replay()
                barrier()
                launch_mode = "cuda graph (one per bag group), replayed"
            except Exception as exc:      # e.g. a collective that cannot be captured: eager launches
                graphs, graph_res = None, None
                launch_mode = f"eager (graph capture failed: {type(exc).__name__})"
                print(f"[rank {rank}] graph capture failed: {exc!r}", file=sys.stderr)
                torch.cuda.synchronize()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        settle = 3
        for i in range(settle):      # the GPU idled while the clock sampler started: back to steady state before the timed region
            if graphs is None:
                step(i)
            else:
                graphs[i % a.groups].replay()
        if graphs is None:
            lib.acmil_prof_enable(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            if graphs is None:
                res = step(i)
            else:
                graphs[i % a.groups].replay()
                res = graph_res[i % a.groups]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if graphs is not None:
            # graph nodes cannot carry timing events: the row-pass kernel is timed by CUDA events around its launch in an
            # eager pass over the same steps, right behind the timed region (clock sampler still running)
            lib.acmil_prof_enable(1)
            for i in range(a.steps):
                step(i)
            barrier()
        import ctypes as C
        main_ms, n_main = C.c_double(0), C.c_int64(0)
        lib.acmil_prof_collect(C.byref(main_ms), C.byref(n_main))
        lib.acmil_prof_enable(0)
        clocks = sampler.stop() if rank == 0 else None
        # ---- the same step on fp16 features (the dtype the reference stores them in, Step2_feature_extract.py:165; widening
        # is exact, the kernels read the halves directly): reported next to the fp32 headline, with its own roofline
        fp16_line = None
        if world == 1 and not a.no_fp16:
            try:
                g16 = [g.half() for g in groups]
                def step16(i):
                    r = torch.rand(S, K_BRANCH, nm, device=dev) if masking else None
                    return op.run(packed, g16[i % a.groups], offsets, n_masked=N_MASKED if masking else 0,
                                  keep=[keep if masking else 0] * S, rand=r, branch_w=branch_w, branch_b=branch_b,
                                  head_w=head_w, head_b=head_b, slide_head=True)
                for i in range(max(a.warmup, 3)):
                    step16(i)
                torch.cuda.synchronize()
                launch16, gr16 = "eager launches", None
                if graphs is not None:      # same launch mode as the headline: one captured graph per bag group, replayed
                    try:
                        gr16 = []
                        for gi in range(a.groups):
                            gg = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(gg):
                                rr = step16(gi)
                            gr16.append((gg, rr))
                        launch16 = "cuda graph (one per bag group), replayed"
                    except Exception:      # noqa: BLE001
                        gr16 = None
                        torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for i in range(a.steps):
                    if gr16 is None:
                        r16 = step16(i)
                    else:
                        gr16[i % a.groups][0].replay()
                        r16 = gr16[i % a.groups][1]
                f1.record()
                torch.cuda.synchronize()
                lib.acmil_prof_enable(1)      # the row-pass kernel alone: CUDA events around its launch in an eager pass
                for i in range(a.steps):
                    step16(i)
                torch.cuda.synchronize()
                k_ms, k_n = C.c_double(0), C.c_int64(0)
                lib.acmil_prof_collect(C.byref(k_ms), C.byref(k_n))
                lib.acmil_prof_enable(0)
                ms16 = f0.elapsed_time(f1) / a.steps
                kms16 = k_ms.value / max(k_n.value, 1)
                bytes16 = (ALGO_BYTES_PER_SLIDE - N_ROWS * D_FEAT * 2) * (a.rows / N_ROWS) * a.slides
                try:
                    peak16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
                except Exception:
                    peak16 = 6650.0
                fp16_line = {"value": S / (ms16 * 1e-3), "unit": "slides/s", "ms_per_step": ms16, "launch": launch16,
                             "roofline": {"bound": "hbm", "achieved": bytes16 / (kms16 * 1e-3) / 1e9, "peak": peak16, "unit": "GB/s",
                                          "frac": bytes16 / (kms16 * 1e-3) / 1e9 / peak16, "kernel_ms": kms16,
                                          "algorithmic_bytes_per_launch": bytes16},
                             "max_abs_logit_diff_vs_fp32_of_same_values": float((r16.slide - op.run(
                                 packed, g16[(a.steps - 1) % a.groups].float(), offsets, n_masked=0, keep=[0] * S, branch_w=branch_w,
                                 branch_b=branch_b, head_w=head_w, head_b=head_b, slide_head=True).slide).abs().max()) if not masking else None,
                             "note": "x_f16 = 1: fp16 rows read by TMA as the hi operand, no x_lo products, half the HBM bytes"}
                del g16
            except Exception as exc:      # noqa: BLE001  (an extra: never take the headline line down with it)
                fp16_line = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
                torch.cuda.synchronize()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        checksum = float(res.slide.sum().item())

        # ---- end to end: pinned host bags -> H2D -> the public batched forward (the call a user makes) -> logits D2H ----
        e2e = None
        if a.e2e_steps > 0:
            Se = a.slides                      # bags per call (each sharded over the ranks at world > 1)
            shard = ShardedACMIL(model, dist.group.WORLD if world > 1 else None, exchange) if world > 1 else None
            e_off = [i * n_loc for i in range(Se + 1)]
            e_tot, e_beg = [a.rows] * Se, [b[rank]] * Se
            if exchange is not None:
                pass      # (sized for S >= Se bags per step)

            def timed_loop(host_bufs, dtype, calls):
                """`calls` calls of Se bags each: H2D of the call's bags from pinned memory -> forward_bags -> logits to the
                host.  A double-buffered feeder (copy stream + events) lets the copy of call i + 1 overlap the kernels and
                the read-back of call i; every call's logits are on the host before the next-but-one call starts."""
                xb = [torch.empty(Se * n_loc, D_FEAT, device=dev, dtype=dtype) for _ in range(2)]
                cstream = torch.cuda.Stream(device=dev)
                ready = [torch.cuda.Event() for _ in range(2)]
                free = [torch.cuda.Event() for _ in range(2)]
                out_host = [torch.empty(Se, N_CLASS).pin_memory() for _ in range(2)]
                done = [torch.cuda.Event() for _ in range(2)]

                def issue(i):
                    k = i % 2
                    with torch.cuda.stream(cstream):
                        cstream.wait_event(free[k])                 # the forward that read this buffer has finished
                        xb[k].copy_(host_bufs[k], non_blocking=True)
                        ready[k].record(cstream)

                def loop(n):
                    issue(0)
                    last = None
                    for i in range(n):
                        k = i % 2
                        if i + 1 < n:
                            issue(i + 1)
                        torch.cuda.current_stream().wait_event(ready[k])
                        rnd = torch.rand(Se, K_BRANCH, nm, device=dev) if (masking and world > 1 and rng_in_sync[0]) else None
                        if shard is None:
                            _, slide, _ = model.forward_bags(xb[k], e_off, want_scores=False)
                        else:
                            _, slide, _ = model.forward_bags(xb[k], e_off, shard_begin=e_beg, n_total=e_tot, exchange=exchange,
                                                             group=dist.group.WORLD if exchange is None else None,
                                                             want_scores=False, rand=rnd)
                        free[k].record()
                        if i >= 2:
                            done[k].synchronize()                   # logits of call i - 2 are on the host
                        out_host[k].copy_(slide, non_blocking=True)
                        done[k].record()
                        last = k
                    torch.cuda.synchronize()
                    return out_host[last]

                loop(2)
                barrier()
                t0 = time.perf_counter()
                loop(calls)
                barrier()
                return time.perf_counter() - t0

            host = [torch.randn(Se * n_loc, D_FEAT).pin_memory() for _ in range(2)]
            calls = a.e2e_steps
            dt = timed_loop(host, torch.float32, calls)
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            n_e2e = calls * Se
            # the same call fed with fp16 features, the dtype the reference stores them in (Step2_feature_extract.py:165;
            # Step3_WSI_classification_ACMIL.py:193 casts after the copy): half the PCIe bytes, cast on the device
            fp16 = None
            if world == 1:
                dt16 = timed_loop([h.half().pin_memory() for h in host], torch.float16, calls)
                fp16 = {"value": n_e2e / dt16, "unit": "slides/s", "h2d_bytes_per_step": Se * n_loc * D_FEAT * 2,
                        "note": "fp16 features in pinned host memory (the H5 storage dtype), read by the kernels as they are (x_f16)"}
            e2e = {"value": n_e2e / dt, "unit": "slides/s", "fp16_features": fp16,
                   "h2d_bytes_per_step": Se * n_loc * D_FEAT * 4 * world,
                   "d2h_bytes_per_step": Se * N_CLASS * 4 * world,
                   "api": f"ACMIL_GA.forward_bags(x_cat, row_offsets): {Se} bags per call (sharded over the ranks at N > 1): pinned "
                          "host -> device copy (double-buffered feeder on a copy stream), fused kernels, logits copied to pinned "
                          "host memory per call",
                   "bags": n_e2e, "bags_per_call": Se}

    eager = None
    if rank == 0 and world == 1 and not a.no_gpu_eager:
        try:
            eager = gpu_eager_rate(dev, a.rows, a.mode)
        except Exception as exc:      # never let the comparator take the bench line down
            eager = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    slides_total = S * a.steps
    value = slides_total / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    main_avg_ms = main_ms.value / max(n_main.value, 1)
    algo_bytes = ALGO_BYTES_PER_SLIDE * (a.rows / N_ROWS) * a.slides        # per launch (this rank's share)
    achieved = algo_bytes / (main_avg_ms * 1e-3) / 1e9 if main_avg_ms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        wl = tr.get("workload", {})
        if wl.get("bags_per_launch") == a.slides and wl.get("rows_per_bag") == a.rows and wl.get("d_feat") == D_FEAT:
            traffic = tr.get(a.mode, {}).get("dram_bytes_per_launch")      # ncu capture of this exact launch shape
    except Exception:
        pass
    line = {
        "metric": "slides/sec (ACMIL ga, N=50k, D=384)", "value": value, "unit": "slides/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "row pass (gp_main_*)", "kernel_ms": main_avg_ms,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * a.steps), "launch": launch_mode,
        "untimed_extra_steps": {"comm_prewarm_steps": comm_prewarm, "settle_steps_after_clock_sampler_start": settle,
                                "note": "outside --warmup and outside the timed region"},
        "exchange": exchange_kind,
        "mask_draw": "identical generator state on every rank (checked), no exchange" if (world > 1 and rng_in_sync[0]) else
                     ("broadcast from rank 0" if world > 1 else "local"),
        "kernel_impl": a.kernel, "checksum": checksum,
    }
    if parity is not None:
        line["parity"] = parity
        line["parity_ok"] = parity["parity_ok"]
    if eager is not None:
        line["gpu_eager_baseline"] = eager
    if fp16_line is not None:
        line["fp16_features"] = fp16_line
    if rank == 0 and world == 1 and not a.no_train_step:
        try:
            line["train_step"] = train_step_rates(dev, a.rows)
        except Exception as exc:      # noqa: BLE001  (an extra: never take the headline line down with it)
            line["train_step"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    if not a.no_cpu_baseline:
        rate, sec = cpu_port_rate(a.rows, 6, 2, a.mode)
        line["cpu_baseline"] = {"value": rate, "unit": "slides/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"6 bags of {a.rows}x{D_FEAT} fp32 after 2 warm-ups (oracle/torch_port.py, "
                                          f"torch CPU ops in the reference's op order)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.workload in ("transmil", "vit", "resnet", "stream"):
        mod = __import__("bench_" + ("vit" if args.workload == "resnet" else args.workload))
        if args.impl == "reference":
            mod.run_reference(args)
        elif not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (acmil_b200 has no CPU path); use --impl reference for the CPU arm")
        else:
            mod.run_ours(args, ClockSampler)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (acmil_b200 has no CPU path); use --impl reference for the CPU arm")
        run_ours(args)
