"""Feature-extraction loop of the reference (Step2_feature_extract.py:35-71, 164-167) on the GPU:
uint8 RGB patches -> Resize(224) + ToTensor + Normalize (one kernel, byte-exact with Pillow) -> encoder -> features,
stored as fp16 like the reference's H5 writer.  Slide reading (openslide) is outside the hot path and absent from this image:
``extract_feature`` takes the decoded patches.  The H5 container: ``store_features`` mirrors the reference's h5py calls (for
an open h5py file); ``acmil_b200.h5bag.write_bags`` / ``H5BagFile`` write and read the same layout without h5py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .transmil import _ptr, _stream

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/dataset_h5.py:22-23
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess(patches_u8: torch.Tensor, out_size: int = 224, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None) -> torch.Tensor:
    """[B, H, W, 3] uint8 CUDA tensor -> [B, 3, out, out] fp32 = eval_transforms(pretrained=True) of dataset_h5.py:20-37."""
    if not patches_u8.is_cuda:
        raise RuntimeError("preprocess: acmil_b200 runs on CUDA tensors only (no CPU path)")
    if patches_u8.dtype != torch.uint8 or patches_u8.dim() != 4 or patches_u8.shape[3] != 3:
        raise TypeError(f"preprocess: expected uint8 [B, H, W, 3], got {patches_u8.dtype} {tuple(patches_u8.shape)}")
    x = patches_u8.contiguous()
    B, H, W, _ = x.shape
    lib = L.load()
    nbytes = C.c_size_t(0)
    L.check(lib.acmil_preprocess_workspace_bytes(H, W, out_size, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
    if out is None:
        out = torch.empty(B, 3, out_size, out_size, device=x.device, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    L.check(lib.acmil_preprocess_u8(_ptr(x), B, H, W, out_size, m3, s3, _ptr(out), _ptr(ws), nbytes.value, _stream(x.device)))
    return out


def to_fp16(x: torch.Tensor) -> torch.Tensor:
    """The ``.astype(np.float16)`` of Step2_feature_extract.py:165 on the device."""
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    L.check(L.load().acmil_f32_to_f16(_ptr(x), _ptr(y), x.numel(), _stream(x.device)))
    return y


@torch.no_grad()
def extract_feature(patches_u8, model, batch_size: int = 256, device=None, target_patch_size: int = 224):
    """Step2_feature_extract.py:35-71 for already-decoded patches: batches of `batch_size` through
    ``model(batch, return_feature=True)``; returns the [N, D] fp32 feature matrix on the host (like the reference)."""
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    src = torch.as_tensor(patches_u8)
    feats = []
    for i in range(0, src.shape[0], batch_size):
        chunk = src[i:i + batch_size]
        chunk = (chunk.pin_memory() if not chunk.is_cuda else chunk).to(dev, non_blocking=True)
        _, f = model(preprocess(chunk, target_patch_size), return_feature=True)
        feats.append(f.cpu())
    return torch.cat(feats, dim=0).numpy()


def store_features(h5file, slide_id, features: np.ndarray, coords: np.ndarray, label):
    """Step2_feature_extract.py:163-167: group per slide with 'feat' (fp16), 'coords', attrs['label']."""
    grp = h5file.create_group(slide_id)
    grp.create_dataset('feat', data=features.astype(np.float16))
    grp.create_dataset('coords', data=coords)
    grp.attrs['label'] = label
    return grp
