"""RemoveNeurons: skilled-neuron removal, no routing
(reference neuron_receivers/remove_skilled_neurons.py:9-57).  Flagged neurons get
gate := -0.17 AFTER the activation, at every timestep; applied in the K1 epilogue."""
import json
import os

import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from moe_b200.sd_modules import GEGLU, GELU  # noqa: F401
from neuron_receivers.predictivity import NeuronPredictivity

REMOVED_GATE_VALUE = -0.17  # remove_skilled_neurons.py:39


class RemoveNeurons(NeuronPredictivity):
    def __init__(self, seed, path_expert_indx, T, n_layers, replace_fn=GEGLU, keep_nsfw=False, remove_timesteps=None,
                 weights_shape=None, **kw):
        super(RemoveNeurons, self).__init__(seed, T, n_layers, replace_fn, keep_nsfw, **kw)
        self.expert_indices = {}
        for i in range(0, T):
            self.expert_indices[i] = {}
            for j in range(0, n_layers):
                with open(os.path.join(path_expert_indx, f'predictivity_{i}_{j}.json'), 'r') as f:
                    self.expert_indices[i][j] = json.load(f)
        self._flags = {}
        self.timestep = 0
        self.layer = 0
        self.gates = []
        self.replace_fn = replace_fn
        self.remove_timesteps = remove_timesteps

    def _override(self, state, device):
        lst = self.expert_indices[self.timestep][self.layer]
        if len(lst) == 0:
            return None
        key = (self.timestep, self.layer)
        if key not in self._flags:
            flags = (np.asarray(lst) == 1)
            if flags.shape[0] != state.hidden:
                raise ValueError(f"neuron flag list has {flags.shape[0]} entries, FFN has {state.hidden} neurons")
            flags = flags[state.layout.perm.numpy()]        # original neuron order -> packed order
            self._flags[key] = torch.from_numpy(flags.astype(np.uint8)).to(device)
        return self._flags[key]

    def hook_fn(self, module, input, output):
        if self.replace_fn != GEGLU:
            raise NotImplementedError("only GEGLU FFNs are implemented natively (PixArt GELU FFNs are out of scope)")
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        H, _, gate = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                  neuron_override=self._override(state, x.device), override_value=REMOVED_GATE_VALUE,
                                  want_scores=False, want_gate=self.capture_gates)
        self._capture(gate, state, lead)
        self.update_time_layer()
        return self._finish(H, state, lead, x)
