"""NeuronPredictivity: (timestep, layer) state machine + per-neuron max activation statistics
(reference neuron_receivers/predictivity.py:9-97).  Base class of the removal receivers."""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from moe_b200.sd_modules import GEGLU, GELU  # noqa: F401
from moe_b200.stats import StatMeter
from neuron_receivers.base_receiver import BaseNeuronReceiver, original_order


class NeuronPredictivity(BaseNeuronReceiver):
    def __init__(self, seed, T, n_layers, replace_fn=GEGLU, keep_nsfw=False, hook_module='unet', **kw):
        super(NeuronPredictivity, self).__init__(seed, replace_fn, keep_nsfw, hook_module, **kw)
        self.T = T
        self.n_layers = n_layers
        self.predictivity = StatMeter(T, n_layers)
        self.max_gate = {t: {l: [] for l in range(n_layers)} for t in range(T)}
        self.timestep = 0
        self.layer = 0
        self.replace_fn = replace_fn

    def update_time_layer(self):
        if self.layer == self.n_layers - 1:
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset_time_layer(self):
        self.timestep = 0
        self.layer = 0
        for t in range(self.T):
            self.max_gate[t] = {l: [] for l in range(self.n_layers)}

    def hook_fn(self, module, input, output):
        """max over all tokens of act(gate) per neuron (predictivity.py:42-53), in the ORIGINAL neuron order
        (what skilled_neuron_ap.py writes and RemoveNeurons reads back); returns the plain GEGLU output."""
        if self.replace_fn != GEGLU:
            raise NotImplementedError("only GEGLU FFNs are implemented natively")
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        H, _, gate = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                  want_scores=False, want_gate=True)
        max_act = original_order(ops.colmax(gate), state).cpu().numpy()    # original neuron order, as the reference's
        self.max_gate[self.timestep][self.layer] = max_act
        self.predictivity.update(max_act, self.timestep, self.layer)
        self.update_time_layer()
        return self._finish(H, state, lead, x)
