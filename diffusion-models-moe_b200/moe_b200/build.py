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


def _headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
        [os.path.join(INCLUDE, "moe_b200.h")]


def _file_digest(src):
    """One translation unit's digest: its own text, every header, the flags."""
    hsh = hashlib.sha256()
    for f in [src] + _headers():
        with open(f, "rb") as fh:
            hsh.update(fh.read())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.hexdigest()


def _digest():
    hsh = hashlib.sha256()
    for f in sources():
        hsh.update(_file_digest(f).encode())
    return hsh.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _compile_one(src, force, verbose):
    """src.cu -> lib/obj<variant>/src.o, skipped when the object's stamp matches (per-file incremental build)."""
    obj_dir = os.path.join(LIB_DIR, "obj" + _SUFFIX)
    os.makedirs(obj_dir, exist_ok=True)
    base = os.path.splitext(os.path.basename(src))[0]
    obj, stamp = os.path.join(obj_dir, base + ".o"), os.path.join(obj_dir, base + ".stamp")
    digest = _file_digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return obj, ""
    flags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
    cmd = [nvcc_path()] + flags + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE, "-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return obj, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a (in parallel, one object per file) and link libmoe_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        if open(STAMP).read().strip() == digest:
            return LIB_PATH
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, force, verbose), srcs))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    link = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB_PATH] + [obj for obj, _ in results]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
