"""RemoveExperts: MoE forward with skilled-expert removal
(reference neuron_receivers/remove_skilled_experts.py:9-55).

Listed experts have their pattern row zeroed iff the list is non-empty and timestep < 20: they
score exactly 0, still compete in the top-k, and own no neurons.  The per-(t, layer) lists are
packed once into E-bit device words; the router kernel applies them (no per-call [E, h] clone)."""
import json
import os

import torch

from moe_b200.packing import bits_from_expert_list
from moe_b200.sd_modules import GEGLU
from neuron_receivers.moefy import routed_ffn
from neuron_receivers.predictivity import NeuronPredictivity

REMOVAL_TIMESTEPS = 20  # hard-coded `self.timestep < 20`, remove_skilled_experts.py:32


class RemoveExperts(NeuronPredictivity):
    def __init__(self, seed, path_expert_indx, T, n_layers, keep_nsfw=False, *, hist=None, count_rows='row0', **kw):
        # (the reference passes keep_nsfw positionally into replace_fn -- SURVEY A.3 item 5; fixed here)
        super(RemoveExperts, self).__init__(seed, T, n_layers, GEGLU, keep_nsfw, **kw)
        self.expert_indices = {}
        for i in range(0, T):
            self.expert_indices[i] = {}
            for j in range(0, n_layers):
                with open(os.path.join(path_expert_indx, f'timestep_{i}_layer_{j}.json'), 'r') as f:
                    self.expert_indices[i][j] = json.load(f)
        self._bits = {}
        self.timestep = 0
        self.layer = 0
        self.gates = []
        # optional fused frequency counter (BASELINE config 3: removal + frequency in one pass)
        self.hist = hist            # int64 [T, n_layers, E_max] device tensor or None
        self.count_rows = count_rows

    def _removed_bits(self, n_experts, device):
        lst = self.expert_indices[self.timestep][self.layer]
        if len(lst) == 0 or self.timestep >= REMOVAL_TIMESTEPS:
            return None
        key = (self.timestep, self.layer)
        if key not in self._bits:
            self._bits[key] = bits_from_expert_list(lst, n_experts).to(device)
        return self._bits[key]

    def hook_fn(self, module, input, output):
        x = input[0]
        removed = None
        hist = None
        rows = (0, 0)
        if getattr(module, 'patterns', None) is not None:
            E = module.patterns.shape[0]
            removed = self._removed_bits(E, x.device)
            if self.hist is not None:
                hist = self.hist[self.timestep, self.layer, :E]
                rows = (0, x.shape[1]) if self.count_rows == 'row0' else (0, x.shape[0] * x.shape[1])
        out, _ = routed_ffn(self, module, x, removed_bits=removed, hist=hist, count_rows=rows)
        self.update_time_layer()
        return out
