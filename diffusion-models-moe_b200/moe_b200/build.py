"""Build libmoe_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python diffusion-models-moe_b200/moe_b200/build.py [--force]

The .so lands in moe_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# MOE_LIB_VARIANT=trace builds / loads a second library with the in-kernel timeline stamps compiled in (tools/ only)
VARIANT = os.environ.get("MOE_LIB_VARIANT", "")
_SUFFIX = ("_" + VARIANT) if VARIANT else ""
LIB_PATH = os.path.join(LIB_DIR, f"libmoe_b200{_SUFFIX}.so")
STAMP = os.path.join(LIB_DIR, f"libmoe_b200{_SUFFIX}.stamp")
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "..", "include"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--shared", "-cudart", "static",
] + (["-DMOE_TRACE=1"] if VARIANT == "trace" else [])


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    hsh = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(INCLUDE, "moe_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            hsh.update(fh.read())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        if open(STAMP).read().strip() == digest:
            return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", INCLUDE, "-o", LIB_PATH] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
