"""
Build ``libswiftly_b200.so`` in-tree with nvcc for sm_100a.

    python -m ska_sdp_distributed_fourier_transform_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is placed next to this file so
that it travels with the source tree; it is git-ignored.
"""

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libswiftly_b200.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build the CUDA extension")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(CSRC, "*.h")) + [os.path.join(ROOT, "include", "swiftly_b200.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in _deps())


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
    hdr_time = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
        return obj, ""
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=False)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + p.stdout.decode())
    return obj, p.stdout.decode()


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libswiftly_b200.so``."""
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJDIR, "*.o")):
            os.remove(f)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile(s, verbose), sources()))
    if verbose:
        for _, log in results:
            sys.stdout.write(log)
    objs = [o for o, _ in results]
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + [
        "-lcudart"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
