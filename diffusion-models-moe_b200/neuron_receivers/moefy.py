"""MOEFy: the core MoEfied GEGLU forward (reference neuron_receivers/moefy.py:10-54)."""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state, moe_ffn_forward
from moe_b200.sd_modules import GEGLU
from neuron_receivers.base_receiver import BaseNeuronReceiver


def routed_geglu(receiver, module, x, *, removed_bits=None, hist=None, count_rows=(0, 0), colmax_out=None,
                 want_idx=False):
    """K1 + K2 for one hooked GEGLU call.  Returns (H packed [T, h], idx | None, state, lead shape).
    Shared by MOEFy / FrequencyMeasure / RemoveExperts."""
    state = get_state(module)
    lead = x.shape[:-1]
    xt = as_tokens(x)
    routed = getattr(module, 'patterns', None) is not None and state.k is not None
    H, scores, gate = ops.geglu_up(xt, state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                   want_scores=routed, want_gate=receiver.capture_gates)
    idx = None
    if routed:
        _, idx = ops.router_topk(scores, module.k, removed_bits=removed_bits, want_bits=False, want_idx=want_idx,
                                 hist=hist, colmax_out=colmax_out, H=H, expert_size=state.expert_size,
                                 count_rows=count_rows)
        if gate is not None:  # captured gate gets the same mask (moefy.py:23-25)
            ops.router_topk(scores, module.k, removed_bits=removed_bits, want_bits=False, H=gate,
                            expert_size=state.expert_size)
    receiver._capture(gate, state, lead)
    return H, idx, state, lead


def routed_ffn(receiver, module, x, *, removed_bits=None, hist=None, count_rows=(0, 0), want_idx=False):
    """One hooked layer call of a routed receiver (MOEFy / FrequencyMeasure / RemoveExperts) -> (hook output, idx).

    With the down-projection fused behind the hook (`BaseNeuronReceiver.fuse_down_proj`, the default inside
    `observe_activation`) and no gate capture this is ONE launch: `moe_ffn_fused` = up-projection + activation ->
    per-token top-k, histogram, masking -> down-projection, returning Y [B, S, d].  Gate capture needs the activated
    gate as a tensor, which only the separate K1 produces: K1 -> K2 (-> K2 on the gate copy) -> K3, still native."""
    state = get_state(module)
    routed = getattr(module, 'patterns', None) is not None and state.k is not None
    if routed and state.fused_down and not receiver.capture_gates:
        y, _, idx = moe_ffn_forward(state, x, removed_bits=removed_bits, hist=hist, count_rows=count_rows,
                                    want_idx=want_idx, k=module.k)
        return (y if y.dtype == x.dtype else y.to(x.dtype)), idx
    H, idx, state, lead = routed_geglu(receiver, module, x, removed_bits=removed_bits, hist=hist, count_rows=count_rows,
                                       want_idx=want_idx)
    return receiver._finish(H, state, lead, x), idx


class MOEFy(BaseNeuronReceiver):
    def __init__(self, seed, **kw):
        super(MOEFy, self).__init__(seed, **kw)

    def hook_fn(self, module, input, output):
        return routed_ffn(self, module, input[0])[0]

    def test(self, model, ann='A brown dog in the snow', relu_condition=False):
        """Reference MOEFy.test (moefy.py:29-54) without the PNG side effects: run hooked, then
        check that gates are non-negative iff the model is ReLU-fied."""
        torch.manual_seed(self.seed)
        np.random.seed(self.seed)
        capture, self.capture_gates = self.capture_gates, True
        try:
            self.observe_activation(model, ann)
            for gate in self.gates:
                assert bool(torch.all(gate >= 0)) == relu_condition, "All gates should be positive"
        finally:
            self.capture_gates = capture
            self.gates = []
