"""Builds libacmil_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C-ABI).

    python -m acmil_b200.build            # incremental
    python -m acmil_b200.build --force
"""
from __future__ import annotations

import hashlib
import os
from concurrent.futures import ThreadPoolExecutor
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# ACMIL_B200_LIB_DIR: build / load a variant library from another in-tree directory (kernel experiments: a second
# build with other -D flags next to the default one); the default is acmil_b200/lib
LIB_DIR = os.environ.get("ACMIL_B200_LIB_DIR") or os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libacmil_b200.so")
STAMP = os.path.join(LIB_DIR, "libacmil_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _flags() -> list[str]:
    extra = os.environ.get("ACMIL_NVCC_EXTRA", "").split()
    return NVCC_FLAGS + extra


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: acmil_b200 needs the CUDA toolkit to build its kernels")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += sorted(os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def up_to_date() -> bool:
    """True when the in-tree library was built from the current sources and flags."""
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and up_to_date():
        return LIB
    # one builder at a time (torchrun ranks all call load() at start-up); the others wait and then find the stamp
    import fcntl
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    dig = _digest()
    log = []

    def compile_one(src: str) -> str:
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), *_flags(), "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        out = "$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + out)
        log.append(out)
        return obj

    # one nvcc per translation unit, all at once (the files are independent)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    # link under a temporary name and rename: a concurrent load() (torchrun ranks) never sees a half-written library
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [_nvcc(), "-shared", "-o", tmp, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    os.replace(tmp, LIB)
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
