#!/usr/bin/env python
"""Secondary bench: the end-to-end stream of BASELINE.json configs[4] -- a Camelyon16-shaped set of slides (398 slides,
10k-100k patches each: SURVEY section 8d C5) goes  uint8 256x256 patches (pinned host memory) -> H2D -> Resize(224) +
normalise -> ViT-S/16 -> fp16 feature bag (Step2_feature_extract.py:35-71, 165) -> ACMIL head on the fp16 bag
(Step3_WSI_classification_ACMIL.py:193; the kernels widen the halves exactly) -> slide logits on the host.
`--workload stream` of bench.py dispatches here; same JSON contract.

Slides are independent units: rank r of N takes slides r, r + N, ... (no collective; the features never leave the GPU that
produced them, which is where the head wants them).  The job is the whole slide set, so N GPUs share a FIXED amount of work:
"scaling": "strong".  A step = one slide.  `--stream-slides S` and `--stream-scale f` bound the run: the slide sizes are the
seeded Camelyon16-shaped draw multiplied by f (default 398 slides at f = 1/64 so that the default run ends within a minute at
one GPU; f = 1 is the full-size stream, about half an hour of ViT-S/16 at one GPU); `value` is slides/s of the set as run and
`patches_per_sec` the size-independent rate behind it.  The timed region holds everything (copies included), so `value` and
`e2e` coincide; `cpu_baseline` is the same chain on the host cores (torch CPU ops in the reference's op order) on a bounded
sample, converted to slides/s of the same slide set.
"""
import json
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
N_SLIDES_C16 = 398
BATCH = 256                       # the reference's extraction batch size (Step2_feature_extract.py:25)


def slide_sizes(n_slides, scale, seed=16):
    rng = np.random.default_rng(seed)
    full = rng.integers(10_000, 100_001, size=n_slides)
    return [max(1, int(round(v * scale))) for v in full], int(full.sum())


def config(a, world, sizes, cpu=False):
    return {"workload": "Camelyon16-shape stream: uint8 256x256 patches -> Resize(224)+normalise -> ViT-S/16 -> fp16 bag -> ACMIL ga "
                        "head (n_token 5, eval) -> slide logits (BASELINE.json configs[4])",
            "slides": len(sizes), "patches": int(sum(sizes)), "patch_count_scale": a.stream_scale,
            "patches_per_slide": f"{min(sizes)}..{max(sizes)} (seeded uniform 10k..100k x scale)",
            "parallelism": "cpu" if cpu else f"slides round-robin over {world} GPU(s), no collective", "batch": BATCH}


def cpu_chain_rate(n_patches, reps=1):
    """patches/s of the host chain: PIL transform + ViT forward (torch CPU ops, timm's op order) + fp16 store; the head's
    share is measured per bag by bench.py --impl reference and is negligible next to the encoder."""
    from bench_vit import cpu_rate
    return cpu_rate(n_patches, reps)


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sizes, _ = slide_sizes(a.stream_slides, a.stream_scale)
    rate = cpu_chain_rate(64, max(1, min(a.steps, 2)))
    slides_s = rate / (sum(sizes) / len(sizes))
    print(json.dumps({
        "impl": "reference", "metric": "slides/sec (Camelyon16-shape stream, ViT-S/16 extract -> ACMIL head)", "value": slides_s,
        "unit": "slides/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": 1, "ms_per_step": 1e3 / slides_s, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(a, 1, sizes, cpu=True),
        "patches_per_sec": rate,
        "cpu_baseline": {"value": slides_s, "unit": "slides/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "batches of 64 patches (PIL/torchvision transform + timm-equivalent torch CPU forward), converted "
                                   "with the mean patches per slide of the set"},
        "e2e": {"value": slides_s, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_ours(a, ClockSampler):
    import torch.distributed as dist
    from acmil_b200 import ACMIL_GA, Struct, _lib
    from acmil_b200.extract import preprocess, to_fp16
    from acmil_b200.vit import CustomModel, vit_small
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    extractor = CustomModel(Struct(n_class=2), vit_small(False, False, None)).to(dev).eval()
    head = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).to(dev).eval()
    sizes, full_patches = slide_sizes(a.stream_slides, a.stream_scale)
    mine = list(range(rank, len(sizes), world))
    # a pool of distinct pinned host batches stands for the decoded patches of the slides (50 MB each; a full slide would
    # be 19.7 GB of pixels); the copies are real: every batch of every slide crosses PCIe inside the timed region
    pool = [torch.randint(0, 256, (BATCH, 256, 256, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
    stage = [torch.empty(BATCH, 256, 256, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    cstream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    max_n = max(sizes)
    bag = torch.empty(max_n, 384, dtype=torch.float16, device=dev)
    out_host = torch.empty(len(mine) + 1, 2).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    counter = [0]

    def run_slide(n, slot):
        """one slide of n patches: batches through the encoder into the fp16 bag, then the head"""
        nb = (n + BATCH - 1) // BATCH

        def issue(j):
            k = counter[0] % 2
            counter[0] += 1
            with torch.cuda.stream(cstream):
                cstream.wait_event(free[k])
                stage[k].copy_(pool[j % len(pool)], non_blocking=True)
                ready[k].record(cstream)
            return k

        k = issue(0)
        for j in range(nb):
            k_next = issue(j + 1) if j + 1 < nb else None
            torch.cuda.current_stream().wait_event(ready[k])
            rows = min(BATCH, n - j * BATCH)
            _, f = extractor(preprocess(stage[k][:rows]), return_feature=True)
            bag[j * BATCH:j * BATCH + rows] = to_fp16(f)
            free[k].record()
            k = k_next
        _, slide, _ = head.forward_bags(bag[:n], [0, n], want_scores=False)
        out_host[slot].copy_(slide[0], non_blocking=True)

    with torch.no_grad():
        for _ in range(2):
            run_slide(min(2 * BATCH, max_n), len(mine))      # warm-up: two batches and a head call
        barrier()
        l0 = _lib.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for slot, s in enumerate(mine):
            run_slide(sizes[s], slot)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
            tl = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
            dist.all_reduce(tl)
            launches = int(tl.item())
        checksum = float(out_host[:len(mine)].sum().item()) if mine else 0.0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_patches = int(sum(sizes))
    value = len(sizes) / (ms * 1e-3)
    from bench_vit import PATCH_FLOPS
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1345.7))
    achieved = PATCH_FLOPS * n_patches / (ms * 1e-3) / 1e12 / world
    line = {
        "metric": "slides/sec (Camelyon16-shape stream, ViT-S/16 extract -> ACMIL head)", "value": value, "unit": "slides/s",
        "n_gpus": world, "steps": len(sizes), "warmup": 2, "ms_per_step": ms / max(len(sizes), 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8 -> f32 (tensor-core products: fp16 hi/lo split on weight GEMMs, 3xTF32 elsewhere; ACMIL_GEMM_SPLIT=tf32: 3xTF32 everywhere) -> f16 bag -> f32 head",
        "data": "synthetic", "config": config(a, world, sizes),
        "patches_per_sec": n_patches / (ms * 1e-3),
        "full_size_estimate": {"slides": N_SLIDES_C16, "patches": full_patches if len(sizes) == N_SLIDES_C16 else None,
                               "seconds_at_this_rate": (full_patches / (n_patches / (ms * 1e-3))) if len(sizes) == N_SLIDES_C16 else None,
                               "note": "the stream is encoder-bound: time scales with the patch count"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "whole stream per GPU (tm_gemm_kernel of the encoder dominates)",
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1345.7") +
                                    "; fp32-faithful products cost 3 fp16 MMAs (weight GEMMs, pre-split images) or 3 TF32 MMAs = 6 bf16-equivalent (the rest)"},
        "clocks": clocks,
        "e2e": {"value": len(sizes) / wall, "unit": "slides/s", "h2d_bytes_per_step": int(n_patches / len(sizes) * 256 * 256 * 3),
                "d2h_bytes_per_step": 8,
                "api": "per slide: pinned uint8 batches -> H2D (double-buffered copy stream) -> extract.preprocess -> CustomModel(ViT-S/16) -> "
                       "extract.to_fp16 into the bag -> ACMIL_GA.forward_bags(fp16 bag) -> logits to pinned host memory"},
        "gpu_launches": int(launches), "checksum": checksum,
    }
    if not a.no_cpu_baseline:
        r = cpu_chain_rate(64, 1)
        line["cpu_baseline"] = {"value": r / (n_patches / len(sizes)), "unit": "slides/s", "cores": torch.get_num_threads(), "kind": "port",
                                "patches_per_sec": r,
                                "sample": "1 batch of 64 patches after 1 warm-up (PIL/torchvision transform + torch CPU ViT-S/16 forward), "
                                          "converted with the mean patches per slide of the set"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
