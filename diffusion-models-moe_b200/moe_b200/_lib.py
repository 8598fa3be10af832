"""ctypes loader for libmoe_b200.so (the C ABI declared in include/moe_b200.h).

There is NO fallback: if the library is missing or does not export a symbol the import of the
compute path raises.  `load()` only dlopen()s and binds signatures -- it never touches the GPU,
so it also works (and is tested) on a CPU-only host.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("MOE_LIB_VARIANT", "")     # "trace": profiling build (tools/ only), see build.py
LIB_PATH = os.path.join(_HERE, "lib", "libmoe_b200" + (("_" + _VARIANT) if _VARIANT else "") + ".so")

c_void_p, c_int, c_float, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/moe_b200.h one to one
SIGNATURES = {
    "moe_abi_version": (c_int, []),
    "moe_last_error": (ctypes.c_char_p, []),
    "moe_launch_count": (c_ll, []),
    "moe_reset_launch_count": (None, []),
    "moe_geglu_up": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "moe_router_topk": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "moe_router_topk_biased": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "moe_colsum_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "moe_rownorm_colsumsq_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "moe_wanda_score_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "moe_mask_vote": (c_int, [c_void_p, c_int, c_ll, c_float, c_void_p, c_void_p]),
    "moe_down_proj": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_size_t,
                              c_void_p]),
    "moe_down_proj_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                     ctypes.c_size_t, c_void_p]),
    "moe_down_proj_workspace_bytes": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "moe_hist_accumulate": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "moe_colmax_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "moe_colmax_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "moe_mask_pack": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "moe_mask_union": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "moe_mask_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "moe_debug_counters": (c_int, [c_void_p, c_int]),
    "moe_debug_trace": (c_int, [c_void_p, c_int]),
    "moe_ffn_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_void_p, ctypes.c_size_t, c_void_p]),
    "moe_ffn_fused_workspace_bytes": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "moe_debug_trace_fused": (c_int, [c_void_p, c_int]),
    "moe_expert_permutation": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       ctypes.c_size_t, c_void_p]),
    "moe_expert_permutation_workspace_bytes": (ctypes.c_size_t, [c_int, c_int]),
    "moe_router_topk_perm": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, ctypes.c_size_t, c_void_p]),
    "moe_down_grouped": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_void_p, ctypes.c_size_t, c_void_p]),
    "moe_cfg_ddim_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_float, c_float, c_void_p]),
    "moe_down_grouped_rows": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "moe_down_grouped_workspace_bytes": (ctypes.c_size_t, [c_int, c_int, c_int, c_int]),
}

ABI_VERSION = 3
_lib = None


class MoeLibraryError(RuntimeError):
    pass


def load():
    """dlopen the library once and bind every declared symbol; raises MoeLibraryError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MoeLibraryError(
            f"{LIB_PATH} not found: build it with `python diffusion-models-moe_b200/moe_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise MoeLibraryError(f"libmoe_b200.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    got = lib.moe_abi_version()
    if got != ABI_VERSION:
        raise MoeLibraryError(f"libmoe_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().moe_last_error().decode(errors="replace")
        raise MoeLibraryError(f"{what} failed (code {rc}): {msg}")
