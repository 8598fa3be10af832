"""SparsityMeasure: captures the activated gate of every hooked GEGLU call and returns the plain GEGLU output
(reference neuron_receivers/sparsity_measure.py:6-18; the exact-zero ratio is computed from `gates` by the caller,
sparsity/check_sparsity.py:41-46)."""
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from neuron_receivers.base_receiver import BaseNeuronReceiver


class SparsityMeasure(BaseNeuronReceiver):
    '''
    Measure sparsity of the model
    '''

    def __init__(self, seed, **kw):
        super(SparsityMeasure, self).__init__(seed, **kw)

    def hook_fn(self, module, input, output):
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        H, _, gate = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                  want_scores=False, want_gate=True)
        self._capture(gate, state, lead)          # reference: self.gates.append(module.gelu(gate).detach().cpu())
        return self._finish(H, state, lead, x)

    def zero_fraction(self):
        """Fraction of exactly-zero activations over everything captured so far (check_sparsity.py:41-46)."""
        n = sum(g.numel() for g in self.gates)
        return float(sum(int((g == 0).sum()) for g in self.gates)) / max(n, 1)

    def test(self, model, ann='A brown dog in the snow'):
        """Reference test (sparsity_measure.py:20-44) without the image side effect: gates of a ReLU-fied model
        are non-negative."""
        torch.manual_seed(0)
        self.observe_activation(model, ann)
        for gate in self.gates:
            assert torch.all(gate >= 0), "Relu failed"
        self.gates = []
