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
    fused_down: bool = False             # while hooked: ff.net.2 is a pass-through and the hook returns Y, not H
    applied_perm: Optional[torch.Tensor] = None   # row / column permutation applied IN PLACE to the model's parameters
    src_versions: tuple = ()             # (data_ptr, _version) of the parameters the packed copies were made from
    gelu_fn: object = None               # the `module.gelu` callable `act` was probed from

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


def _param_versions(*params) -> tuple:
    return tuple((None if p is None else (p.data_ptr(), p._version)) for p in params)


def _gelu_identity(module):
    fn = getattr(module, "gelu", None)
    return getattr(fn, "__func__", fn)      # bound methods are re-created on every attribute access


@torch.no_grad()
def undo_model_permutation(module: GEGLU) -> None:
    """Put a GEGLU (and its down-projection) that `attach_state` permuted in place back into the original
    neuron order, so that a second `modify_ffn` (another label file, another top-k) starts from the model the
    reference would see (the reference's modify_ffn is idempotent: it only attaches `patterns` / `k`)."""
    st = getattr(module, "_moe_state", None)
    if st is None or st.applied_perm is None:
        return
    w1, b1 = module.proj.weight, module.proj.bias
    h = w1.shape[0] // 2
    inv = torch.empty_like(st.applied_perm)
    inv[st.applied_perm] = torch.arange(h)
    inv = inv.to(w1.device)
    rows = torch.cat([inv, inv + h])
    w1.data.copy_(w1.data[rows].clone())
    if b1 is not None:
        b1.data.copy_(b1.data[rows].clone())
    down = st.down_module
    if down is not None:
        down.weight.data.copy_(down.weight.data[:, inv.to(down.weight.device)].clone())
        if hasattr(down, "_moe_column_perm"):
            down._moe_column_perm = None
    st.applied_perm = None
    module._moe_state = None


@torch.no_grad()
def attach_state(module: GEGLU, layout: Optional[ExpertLayout] = None, k: Optional[int] = None,
                 down: Optional[nn.Module] = None, permute_model_weights: bool = True) -> FFNState:
    """Build (or rebuild) `module._moe_state`.

    With `permute_model_weights` AND a down-projection to permute along, the module's own parameters are
    permuted in place into packed neuron order (W1 value/gate rows + b1, and the columns of the sibling
    down-projection), which leaves the model's function unchanged but lets the GEGLU output flow into the
    stock `ff.net.2` in packed order; bf16 parameters are then aliased by the kernels, not copied.  Without a
    down-projection (`helper.modify_ffn(ffn, path, k)`, the reference's exact signature) the model is left
    untouched and the kernels get packed COPIES; the hook output is translated back to the model's order.
    Re-attaching (e.g. a new top-k) first undoes a permutation applied earlier.
    """
    undo_model_permutation(module)
    w1, b1 = module.proj.weight, module.proj.bias
    h = w1.shape[0] // 2
    if layout is None:
        layout = ExpertLayout.contiguous(h // default_expert_size(h), default_expert_size(h))
    if layout.hidden != h:
        raise ValueError(f"expert layout covers {layout.hidden} neurons but the FFN has {h}")
    identity = bool(torch.equal(layout.perm, torch.arange(h)))
    w2 = None if down is None else down.weight
    b2 = None if down is None else down.bias
    applied = None
    if permute_model_weights and not identity and down is not None:
        perm = layout.perm.to(w1.device)
        rows = torch.cat([perm, perm + h])
        w1.data.copy_(w1.data[rows].clone())
        if b1 is not None:
            b1.data.copy_(b1.data[rows].clone())
        w2.data.copy_(w2.data[:, perm.to(w2.device)].clone())
        # the parameters now ARE in packed order; `layout.perm` is kept to translate per-neuron
        # artefacts (removal flags, Wanda mask columns, captured gates) given in original order
        packed = pack_ffn(ExpertLayout.contiguous(layout.n_experts, layout.expert_size), w1, b1, w2, b2)
        permuted = True
        applied = layout.perm.clone()
    else:
        packed = pack_ffn(layout, w1, b1, w2, b2)
        permuted = identity
    state = FFNState(layout=layout, w1p=packed.w1p, b1p=packed.b1p, w2p=packed.w2p, b2=packed.b2, k=k,
                     act=activation_code(module), weights_permuted_in_model=permuted, down_module=down,
                     applied_perm=applied, src_versions=_param_versions(w1, b1, w2, b2),
                     gelu_fn=_gelu_identity(module))
    module._moe_state = state
    return state


@torch.no_grad()
def refresh_state(module: GEGLU, state: FFNState) -> None:
    """Re-make the packed bf16 / f32 copies after the model's parameters changed (optimizer step, LoRA merge,
    `wanda_scoring.bake`); aliased bf16 parameters need no copy but are re-aliased all the same."""
    w1, b1 = module.proj.weight, module.proj.bias
    down = state.down_module
    w2 = None if down is None else down.weight
    b2 = None if down is None else down.bias
    lay = ExpertLayout.contiguous(state.layout.n_experts, state.layout.expert_size) \
        if state.weights_permuted_in_model else state.layout
    packed = pack_ffn(lay, w1, b1, w2, b2)
    state.w1p, state.b1p, state.w2p, state.b2 = packed.w1p, packed.b1p, packed.w2p, packed.b2
    state.src_versions = _param_versions(w1, b1, w2, b2)


def attach_down(module: GEGLU, state: FFNState, down: nn.Module) -> None:
    """Give a state that was built without its down-projection (plain `modify_ffn(ffn, path, k)`, or an FFN that
    was never MoEfied) the packed W2 / b2 the native down-projection needs.  The model is not modified."""
    state.down_module = down
    refresh_state(module, state)


def get_state(module: GEGLU) -> FFNState:
    """State of a hooked GEGLU; FFNs that were never MoEfied get an identity layout on first use.  Packed copies
    are refreshed when the parameters they were made from have changed (data pointer or version counter)."""
    st = getattr(module, "_moe_state", None)
    if st is None:
        st = attach_state(module, None, None, None, permute_model_weights=False)
    down = st.down_module
    if st.src_versions != _param_versions(module.proj.weight, module.proj.bias,
                                          None if down is None else down.weight, None if down is None else down.bias):
        refresh_state(module, st)
    if st.gelu_fn is not _gelu_identity(module):      # module.gelu swapped (relufy_model.py:35): probe it again
        st.act = activation_code(module)
        st.gelu_fn = _gelu_identity(module)
    return st


def as_tokens(x: torch.Tensor) -> torch.Tensor:
    """[B, S, d] (any float dtype) -> contiguous bf16 [B*S, d] on the same device."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    return x2.contiguous()


def moe_ffn_forward(state: FFNState, x: torch.Tensor, *, removed_bits=None, hist=None, count_rows=(0, 0),
                    colmax_out=None, neuron_override=None, override_value: float = -0.17, want_bits=False,
                    want_idx=False, w2_override=None, route: bool = True, k: Optional[int] = None, H_out=None):
    """Whole hot path for one layer call.  x [B, S, d] -> (y [B, S, d] bf16, bits, idx).
    `k` overrides `state.k` (the reference reads `module.k` at hook time); `H_out` receives the masked hidden
    state [T, h] (packed neuron order)."""
    if state.w2p is None:
        raise ValueError("FFNState has no down-projection weight")
    lead = x.shape[:-1]
    xt = as_tokens(x)
    k = state.k if k is None else k
    do_route = route and k is not None
    w2 = state.w2p if w2_override is None else w2_override
    if do_route and neuron_override is None and colmax_out is None and not state.fused_unsupported:
        # one persistent kernel for the whole layer call (K1 -> routing -> K3); geometries it does not cover
        # (MOE_ERR_UNSUPPORTED_SHAPE) use the three separate launches below -- both are the CUDA path
        try:
            y, _, _, bits, idx = ops.ffn_fused(xt, state.w1p, state.b1p, w2, state.b2, state.n_experts,
                                               state.expert_size, k, state.act, removed_bits=removed_bits,
                                               want_bits=want_bits, want_idx=want_idx, hist=hist, count_rows=count_rows,
                                               H_out=H_out)
            return y.view(*lead, y.shape[-1]), bits, idx
        except _lib.MoeLibraryError as e:
            if "code -2" not in str(e):
                raise
            state.fused_unsupported = True
    H, scores, _ = ops.geglu_up(xt, state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                neuron_override=neuron_override, override_value=override_value,
                                want_scores=do_route, out=H_out)
    bits = idx = None
    if do_route:
        bits, idx = ops.router_topk(scores, k, removed_bits=removed_bits, want_bits=want_bits, want_idx=want_idx,
                                    hist=hist, colmax_out=colmax_out, H=H, expert_size=state.expert_size,
                                    count_rows=count_rows)
    y = ops.down_proj(H, w2, state.b2)
    return y.view(*lead, y.shape[-1]), bits, idx
