import glob
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when someone runs the whole suite on a CPU-only box
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """-> (weights dict without the 'w::' prefix, everything else dict)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    rest = {k: z[k] for k in z.files if not k.startswith("w::")}
    return w, rest


def golden_x(meta):
    """Regenerate the seeded input exactly as tests/golden/make_golden.py did and check its hash."""
    import torch
    g = torch.Generator().manual_seed(int(meta["meta_x_seed"]))
    shape = tuple(int(v) for v in meta["meta_x_shape"])
    x = torch.randn(*shape, generator=g) * float(meta.get("meta_x_scale", 1.0))
    if int(meta.get("meta_x_fp16", 0)):
        x = x.half().float()
    h = hashlib.sha256(x.contiguous().numpy().tobytes()).hexdigest()
    assert h == str(meta["meta_x_sha"]), "seeded input differs from the one the golden vector was made with"
    return x


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def np_seeded_state(shapes: dict, seed: int, scale: float = 0.05):
    """Same draw order as make_golden.np_state (dict order = state_dict order)."""
    rng = np.random.default_rng(seed)
    return {k: (rng.standard_normal(tuple(s)) * scale).astype(np.float32) for k, s in shapes.items()}
