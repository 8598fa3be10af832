"""moe_b200: B200-native (sm_100a) kernels for the MoEfied GEGLU feed-forward hot path.

Python is only the host: tensors live in PyTorch, the arithmetic happens in libmoe_b200.so
(hand-written CUDA, C ABI in include/moe_b200.h) reached through ctypes.
"""
from . import _lib  # noqa: F401
from .ops import (  # noqa: F401
    ACT_GELU, ACT_RELU, geglu_up, router_topk, down_proj, hist_accumulate, colmax, mask_pack, mask_union,
    mask_weights, launch_count, reset_launch_count, ffn_fused, fused_workspace, colsum, rownorm_colsumsq, wanda_score_mask, mask_vote,
    expert_permutation, down_grouped, ExpertPermutation, cfg_ddim_step,
)
from .packing import ExpertLayout, pack_ffn, bits_from_expert_list  # noqa: F401
