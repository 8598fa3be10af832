"""MOEFy: the core MoEfied GEGLU forward (reference neuron_receivers/moefy.py:10-54)."""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
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


class MOEFy(BaseNeuronReceiver):
    def __init__(self, seed, **kw):
        super(MOEFy, self).__init__(seed, **kw)

    def hook_fn(self, module, input, output):
        x = input[0]
        H, _, state, lead = routed_geglu(self, module, x)
        return self._finish(H, state, lead, x)

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
