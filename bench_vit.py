#!/usr/bin/env python
"""Secondary bench: ViT-S/16 patch-feature extraction (BASELINE.json configs[3]: 256x256 RGB uint8 patches -> Resize(224)
+ normalise -> ViT-S/16 -> fp16 features).  Same JSON contract as bench.py (`--workload vit` dispatches here).
A step = one batch of `--patch-batch` patches (default 256, the reference's batch size) per GPU; `value` = patches/s over
all GPUs with the uint8 patches resident in HBM; `e2e` = extract_feature() from pinned host memory (H2D of the uint8
patches, features read back); roofline bound "tensor": useful FLOPs (2 M N K, split not counted) / time against the
measured bf16 peak.  Patches are independent: N GPUs = N data-parallel replicas, no collective.
"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PATCH_FLOPS = 2 * (196 * 768 * 384 + 12 * (197 * 384 * 1152 + 2 * 6 * 197 * 197 * 64 + 197 * 384 * 384 + 2 * 197 * 384 * 1536))
# ResNet18 at 224x224 (models.py:13-77): conv1 + 4 stages of 2 BasicBlocks (+ 3 strided 1x1 downsample convolutions)
RESNET_FLOPS = 2 * (112 * 112 * 147 * 64 + 4 * 56 * 56 * 576 * 64
                    + 28 * 28 * (576 * 128 + 3 * 1152 * 128 + 64 * 128) + 14 * 14 * (1152 * 256 + 3 * 2304 * 256 + 128 * 256)
                    + 7 * 7 * (2304 * 512 + 3 * 4608 * 512 + 256 * 512))


def is_resnet(a):
    return getattr(a, "workload", "vit") == "resnet"


def names(a):
    return ("ResNet18", RESNET_FLOPS, 512) if is_resnet(a) else ("ViT-S/16", PATCH_FLOPS, 384)


def cpu_rate(batch, reps, resnet=False):
    from acmil_b200.vit import vit_small
    from oracle import preprocess as OP
    import numpy as np
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    p = {k: v.detach() for k, v in vit_small(False, False, None).state_dict().items()}

    def fwd(x):      # timm VisionTransformer.forward with torch CPU ops
        B = x.shape[0]
        t = F.conv2d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=16).flatten(2).transpose(1, 2)
        t = torch.cat([p["cls_token"].expand(B, -1, -1), t], 1) + p["pos_embed"]
        for i in range(12):
            b = f"blocks.{i}."
            y = F.layer_norm(t, (384,), p[b + "norm1.weight"], p[b + "norm1.bias"], 1e-6)
            qkv = F.linear(y, p[b + "attn.qkv.weight"], p[b + "attn.qkv.bias"]).reshape(B, 197, 3, 6, 64).permute(2, 0, 3, 1, 4)
            a = torch.softmax((qkv[0] * 0.125) @ qkv[1].transpose(-1, -2), -1)
            t = t + F.linear((a @ qkv[2]).transpose(1, 2).reshape(B, 197, 384), p[b + "attn.proj.weight"], p[b + "attn.proj.bias"])
            y = F.layer_norm(t, (384,), p[b + "norm2.weight"], p[b + "norm2.bias"], 1e-6)
            t = t + F.linear(F.gelu(F.linear(y, p[b + "mlp.fc1.weight"], p[b + "mlp.fc1.bias"])), p[b + "mlp.fc2.weight"], p[b + "mlp.fc2.bias"])
        return F.layer_norm(t, (384,), p["norm.weight"], p["norm.bias"], 1e-6)[:, 0]

    if resnet:      # the reference's own module graph (torchvision BasicBlocks) on the host cores
        from acmil_b200.resnet import resnet18
        rn = resnet18(pretrained=False).eval()

        def fwd(x):      # models.py:54-73 with class_classifier = Identity (models.py:201-204)
            x = rn.maxpool(rn.relu(rn.bn1(rn.conv1(x))))
            x = rn.layer4(rn.layer3(rn.layer2(rn.layer1(x))))
            return rn.avgpool(x).flatten(1)

    from PIL import Image
    from torchvision import transforms
    tr = transforms.Compose([transforms.Resize(224), transforms.ToTensor(),
                             transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))])
    rng = np.random.default_rng(0)
    patches = rng.integers(0, 256, (batch, 256, 256, 3), dtype=np.uint8)
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            x = torch.stack([tr(Image.fromarray(pt)) for pt in patches])      # the reference's DataLoader workers
            fwd(x).half()
            if i:
                times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times)


def config(a, world, batch, cpu=False):
    return {"workload": f"{names(a)[0]} patch-feature extraction: uint8 256x256 RGB patches -> Resize(224)+normalise -> encoder -> fp16 "
                        "features (BASELINE.json configs[3]; 100k patches = one slide), fp32-faithful split products",
            "patches_per_step": batch * (1 if cpu else world), "parallelism": "cpu" if cpu else f"{world} data-parallel replica(s)",
            "l2_policy": "rotating resident patch batches (50 MB uint8 each, 154 MB fp32 after resize); activations of a batch exceed L2"}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    batch = 64
    rate = cpu_rate(batch, max(1, min(a.steps, 2)), is_resnet(a))
    print(json.dumps({
        "impl": "reference", "metric": f"patches/sec ({names(a)[0]} feature extraction)", "value": rate, "unit": "patches/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": 1, "ms_per_step": batch / rate * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(a, 1, batch, cpu=True),
        "cpu_baseline": {"value": rate, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"batches of {batch} patches: PIL/torchvision transform + timm-equivalent torch CPU forward"},
        "e2e": {"value": rate, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_ours(a, ClockSampler):
    import torch.distributed as dist
    from acmil_b200 import Struct, _lib
    from acmil_b200.extract import extract_feature, preprocess, to_fp16
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
    if is_resnet(a):
        from acmil_b200.resnet import resnet18
        enc = resnet18(pretrained=False)
        enc.class_classifier = torch.nn.Identity()
        enc.embed_dim = enc.inplanes
    else:
        enc = vit_small(False, False, None)
    model = CustomModel(Struct(n_class=2), enc).to(dev).eval()
    enc_name, flops, feat_dim = names(a)
    batch = getattr(a, "patch_batch", 256)      # the reference's extraction batch size (Step2_feature_extract.py:25)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    bags = [torch.randint(0, 256, (batch, 256, 256, 3), device=dev, dtype=torch.uint8, generator=gen) for _ in range(3)]

    def step(i):
        _, f = model(preprocess(bags[i % 3]), return_feature=True)
        return to_fp16(f)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(a.warmup, 3)):
            f = step(i)
        barrier()
        l0 = _lib.launch_count()
        step(0)
        torch.cuda.synchronize()
        launches = _lib.launch_count() - l0
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            f = step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        checksum = float(f.float().sum().item())
        host = torch.randint(0, 256, (2 * batch, 256, 256, 3), dtype=torch.uint8).pin_memory()
        extract_feature(host[:batch], model, batch_size=batch)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, a.e2e_steps)
        for i in range(n_e2e):
            extract_feature(host, model, batch_size=batch)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1345.7))
    sec = ms * 1e-3 / a.steps
    achieved = flops * batch / sec / 1e12
    rate = world * batch * a.steps / (ms * 1e-3)
    line = {
        "metric": f"patches/sec ({enc_name} feature extraction)", "value": rate, "unit": "patches/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 -> f32 (tensor-core products: fp16 hi/lo split on weight GEMMs, 3xTF32 elsewhere; ACMIL_GEMM_SPLIT=tf32: 3xTF32 everywhere) -> f16", "data": "synthetic",
        "config": config(a, world, batch), "slides_per_sec_100k_patches": rate / 1e5,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "whole step (tm_gemm_kernel dominates)", "algorithmic_flops_per_patch": flops,
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1345.7") +
                                    "; fp32-faithful products cost 3 fp16 MMAs (weight GEMMs, pre-split images) or 3 TF32 MMAs = 6 bf16-equivalent (the rest)"},
        "clocks": clocks,
        "e2e": {"value": world * n_e2e * 2 * batch / dt, "unit": "patches/s", "h2d_bytes_per_step": world * batch * 256 * 256 * 3,
                "d2h_bytes_per_step": world * batch * feat_dim * 4,
                "api": "extract_feature(uint8 patches in pinned host memory, model, batch_size): H2D, preprocess, encoder, features .cpu()"},
        "gpu_launches": int(launches * a.steps), "checksum": checksum,
    }
    if not a.no_cpu_baseline:
        r = cpu_rate(64, 1, is_resnet(a))
        line["cpu_baseline"] = {"value": r, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"1 batch of 64 patches after 1 warm-up: PIL/torchvision transform + torch CPU {enc_name} forward"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
