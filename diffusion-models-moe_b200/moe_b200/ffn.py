"""Per-FFN device state and the fused MoE feed-forward built from the C-ABI kernels.

`FFNState` holds what the kernels need for one transformer-block FFN: the packed bf16
up-projection weight, f32 biases, the (optionally packed) down-projection weight and the expert
layout.  `moe_ffn_forward` is the whole hot path for one layer call:

    moe_ffn_fused (one persistent kernel: up-projection -> routing -> down-projection), or, for geometries
    it does not cover,  K1 geglu_up -> K2 router_topk (select + histogram + zero inactive experts) -> K3 down_proj
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import ops, _lib
from .packing import ExpertLayout, pack_ffn
from .sd_modules import GEGLU


@dataclass
class FFNState:
    layout: ExpertLayout
    w1p: torch.Tensor                    # bf16 [2h, d], packed neuron order
    b1p: Optional[torch.Tensor]          # f32 [2h]
    w2p: Optional[torch.Tensor] = None   # bf16 [d, h], packed columns
    b2: Optional[torch.Tensor] = None    # f32 [d]
    k: Optional[int] = None              # experts selected per token (None: not MoEfied)
    act: int = ops.ACT_GELU
    weights_permuted_in_model: bool = False
    down_module: Optional[nn.Module] = None
    fused_unsupported: bool = False      # set once moe_ffn_fused has rejected this geometry

    @property
    def n_experts(self) -> int:
        return self.layout.n_experts

    @property
    def expert_size(self) -> int:
        return self.layout.expert_size

    @property
    def hidden(self) -> int:
        return self.layout.hidden


def activation_code(module) -> int:
    """Which activation `module.gelu` computes, decided by probing it (the reference swaps it for
    a function called `relu`, sparsity/relufy_model.py:8-40).  Anything that is neither exact
    GELU nor ReLU is rejected -- the CUDA epilogue implements only those two."""
    fn = getattr(module, "gelu", None)
    if fn is None:
        return ops.ACT_GELU
    probe = torch.tensor([-1.0, -0.25, 2.0])
    with torch.no_grad():
        got = fn(probe).float()
    if torch.equal(got, torch.relu(probe)):
        return ops.ACT_RELU
    if torch.allclose(got, torch.nn.functional.gelu(probe), atol=1e-6):
        return ops.ACT_GELU
    raise ValueError("module.gelu is neither exact (erf) GELU nor ReLU; unsupported by the CUDA path")


def default_expert_size(hidden: int) -> int:
    """Segment width for FFNs that are not MoEfied (no routing): any tile-friendly divisor of h."""
    for es in (64, 32, 16, 8, 4):
        if hidden % es == 0:
            return es
    raise ValueError(f"hidden size {hidden} is not a multiple of 4; unsupported by the CUDA path")


def find_down_proj(root: nn.Module, geglu_name: str) -> Optional[nn.Module]:
    """`...ff.net.0` -> the sibling `...ff.net.2` Linear (upstream FeedForward.net[2])."""
    if not geglu_name.endswith("net.0"):
        return None
    try:
        return root.get_submodule(geglu_name[:-1] + "2")
    except AttributeError:
        return None


@torch.no_grad()
def attach_state(module: GEGLU, layout: Optional[ExpertLayout] = None, k: Optional[int] = None,
                 down: Optional[nn.Module] = None, permute_model_weights: bool = True) -> FFNState:
    """Build (or rebuild) `module._moe_state`.

    With `permute_model_weights` the module's own parameters are permuted in place into packed
    neuron order (W1 value/gate rows + b1, and the columns of the sibling down-projection), which
    leaves the model's function unchanged but lets the GEGLU output flow into the stock
    `ff.net.2` in packed order.  bf16 parameters are then aliased by the kernels, not copied.
    """
    w1, b1 = module.proj.weight, module.proj.bias
    h = w1.shape[0] // 2
    if layout is None:
        layout = ExpertLayout.contiguous(h // default_expert_size(h), default_expert_size(h))
    if layout.hidden != h:
        raise ValueError(f"expert layout covers {layout.hidden} neurons but the FFN has {h}")
    identity = bool(torch.equal(layout.perm, torch.arange(h)))
    w2 = None if down is None else down.weight
    b2 = None if down is None else down.bias
    if permute_model_weights and not identity:
        perm = layout.perm.to(w1.device)
        rows = torch.cat([perm, perm + h])
        w1.data.copy_(w1.data[rows].clone())
        if b1 is not None:
            b1.data.copy_(b1.data[rows].clone())
        if w2 is not None:
            w2.data.copy_(w2.data[:, perm].clone())
        # the parameters now ARE in packed order; `layout.perm` is kept to translate per-neuron
        # artefacts (removal flags, Wanda mask columns, captured gates) given in original order
        packed = pack_ffn(ExpertLayout.contiguous(layout.n_experts, layout.expert_size), w1, b1, w2, b2)
        permuted = True
    else:
        packed = pack_ffn(layout, w1, b1, w2, b2)
        permuted = identity
    state = FFNState(layout=layout, w1p=packed.w1p, b1p=packed.b1p, w2p=packed.w2p, b2=packed.b2, k=k,
                     act=activation_code(module), weights_permuted_in_model=permuted, down_module=down)
    module._moe_state = state
    return state


def get_state(module: GEGLU) -> FFNState:
    """State of a hooked GEGLU; FFNs that were never MoEfied get an identity layout on first use."""
    st = getattr(module, "_moe_state", None)
    if st is None:
        st = attach_state(module, None, None, None, permute_model_weights=False)
    st.act = activation_code(module)
    return st


def as_tokens(x: torch.Tensor) -> torch.Tensor:
    """[B, S, d] (any float dtype) -> contiguous bf16 [B*S, d] on the same device."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    return x2.contiguous()


def moe_ffn_forward(state: FFNState, x: torch.Tensor, *, removed_bits=None, hist=None, count_rows=(0, 0),
                    colmax_out=None, neuron_override=None, override_value: float = -0.17, want_bits=False,
                    want_idx=False, w2_override=None, route: bool = True):
    """Whole hot path for one layer call.  x [B, S, d] -> (y [B, S, d] bf16, bits, idx)."""
    if state.w2p is None:
        raise ValueError("FFNState has no down-projection weight")
    lead = x.shape[:-1]
    xt = as_tokens(x)
    do_route = route and state.k is not None
    w2 = state.w2p if w2_override is None else w2_override
    if do_route and neuron_override is None and colmax_out is None and not state.fused_unsupported:
        # one persistent kernel for the whole layer call (K1 -> routing -> K3); geometries it does not cover
        # (MOE_ERR_UNSUPPORTED_SHAPE) use the three separate launches below -- both are the CUDA path
        try:
            y, _, _, bits, idx = ops.ffn_fused(xt, state.w1p, state.b1p, w2, state.b2, state.n_experts,
                                               state.expert_size, state.k, state.act, removed_bits=removed_bits,
                                               want_bits=want_bits, want_idx=want_idx, hist=hist, count_rows=count_rows)
            return y.view(*lead, y.shape[-1]), bits, idx
        except _lib.MoeLibraryError as e:
            if "code -2" not in str(e):
                raise
            state.fused_unsupported = True
    H, scores, _ = ops.geglu_up(xt, state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                neuron_override=neuron_override, override_value=override_value,
                                want_scores=do_route)
    bits = idx = None
    if do_route:
        bits, idx = ops.router_topk(scores, state.k, removed_bits=removed_bits, want_bits=want_bits, want_idx=want_idx,
                                    hist=hist, colmax_out=colmax_out, H=H, expert_size=state.expert_size,
                                    count_rows=count_rows)
    y = ops.down_proj(H, w2, state.b2)
    return y.view(*lead, y.shape[-1]), bits, idx
