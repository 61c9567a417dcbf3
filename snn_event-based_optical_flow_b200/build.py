"""Build libsnnflow.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python snn_event-based_optical_flow_b200/build.py [--force] [-v]

The shared object is git-ignored but travels to the GPU box with the working tree.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsnnflow.so")
STAMP = os.path.join(HERE, ".libsnnflow.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(os.path.dirname(HERE), "include", "snnflow.h")]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libsnnflow.so (skipped when sources are unchanged)."""
    dig = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsnnflow.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
