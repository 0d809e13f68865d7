"""Build libqvit_b200.so (sm_100a only) in-tree with nvcc.

    python -m quantized_vit_b200.build [--force]

Every .cu under csrc/ is compiled with ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (the ``a`` target is
required: tcgen05 / TMEM instructions do not assemble for plain compute_100) and linked into ONE shared
library next to this file, so the built artefact travels with the repository snapshot to the GPU box.
No torch headers are involved: the library has a plain C ABI (include/qvit_b200.h).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libqvit_b200.so")
BUILD_DIR = os.path.join(HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libqvit_b200.so")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    include = os.path.join(os.path.dirname(HERE), "include", "qvit_b200.h")
    for p in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [include]:
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(BUILD_DIR, "fingerprint")
    return os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == _fingerprint()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of libqvit_b200.so."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(os.path.join(BUILD_DIR, "fingerprint"), "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
