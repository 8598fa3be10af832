"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, and exports exactly the
symbols include/moe_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "moe_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MOE_API\s+[\w\s\*]+?\b(moe_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["moe_geglu_up", "moe_router_topk", "moe_down_proj", "moe_hist_accumulate", "moe_mask_weights",
                 "moe_mask_union", "moe_mask_pack", "moe_colmax_f32", "moe_last_error", "moe_abi_version"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from moe_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in moe_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared_symbols(), "ctypes signature table out of sync with the header"
    assert lib.moe_abi_version() == _lib.ABI_VERSION


def test_no_extra_exports(lib):
    from moe_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()


def test_sass_uses_blackwell_tensor_path(lib):
    """The GEMMs must be tcgen05 (UTCHMMA) fed by TMA (UTMALDG) with TMEM loads (LDTM), not mma.sync."""
    from moe_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout


def test_argument_validation_without_gpu(lib):
    """Argument checks run before any CUDA call, so error paths are testable on CPU."""
    rc = lib.moe_router_topk(None, None, 1, None, None, None, None, None, 0, 0, 4, 8, 0, 0, None)
    assert rc == -1 and b"scores is NULL" in lib.moe_last_error()
    rc = lib.moe_geglu_up(1, 1, None, None, 0.0, 1, None, None, 4, 32, 100, 10, 10, 0, None)
    assert rc == -2 and b"multiples of 8" in lib.moe_last_error()
    rc = lib.moe_geglu_up(1, 1, None, None, 0.0, 16, None, None, 4, 32, 96, 16, 6, 0, None)
    assert rc == -2 and b"expert size" in lib.moe_last_error()
    rc = lib.moe_mask_weights(16, 16, 16, 4, 48, None)  # h % 32
    assert rc == -2


def test_missing_library_fails_loudly(monkeypatch):
    from moe_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmoe_b200.so")
    with pytest.raises(_lib.MoeLibraryError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_ops_refuse_cpu_tensors(lib):
    import torch
    import moe_b200 as M
    with pytest.raises(ValueError, match="CUDA tensor"):
        M.router_topk(torch.zeros(4, 8), 2)
    with pytest.raises(ValueError, match="CUDA tensor"):
        M.down_proj(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(16, 8, dtype=torch.bfloat16), None)
