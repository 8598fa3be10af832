"""AddExperts: MoE forward in which the listed skilled experts get their score raised by 5 x std[e] before a
top-int(0.8 k) selection (reference neuron_receivers/add_skilled_experts.py:8-62).

The boost is a per-expert score bias applied inside the router kernel (moe_router_topk_biased); per (t, layer) bias
vectors are built once and stay on the device."""
import json
import os

import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from moe_b200.sd_modules import GEGLU
from neuron_receivers.predictivity import NeuronPredictivity

BOOST = 5.0            # add_skilled_experts.py:56
K_FRACTION = 0.8       # add_skilled_experts.py:58


class AddExperts(NeuronPredictivity):
    def __init__(self, seed, path_expert_indx, T, n_layers, keep_nsfw=False, **kw):
        # (the reference passes keep_nsfw positionally into replace_fn -- SURVEY A.3 item 5; fixed here)
        super(AddExperts, self).__init__(seed, T, n_layers, GEGLU, keep_nsfw, **kw)
        adj = path_expert_indx.rstrip('/').split('/')[-3]
        base_path = path_expert_indx.split(adj)[0]
        with open(os.path.join(base_path, adj, 'predictivity_base_expert.json'), 'r') as f:
            activation_data = json.load(f)
        self.expert_indices = {}
        self.avg_activation = {}
        for i in range(0, T):
            self.expert_indices[i] = {}
            self.avg_activation[i] = {}
            for j in range(0, n_layers):
                with open(os.path.join(path_expert_indx, f'timestep_{i}_layer_{j}.json'), 'r') as f:
                    self.expert_indices[i][j] = json.load(f)
                self.avg_activation[i][j] = activation_data['time_steps'][str(i)][str(j)]['std']
        self._bias = {}
        self.timestep = 0
        self.layer = 0
        self.gates = []

    def _score_bias(self, n_experts, device):
        key = (self.timestep, self.layer)
        if key not in self._bias:
            idx = list(self.expert_indices[self.timestep][self.layer])
            bias = torch.zeros(n_experts, dtype=torch.float32)
            if idx:
                bias[idx] = BOOST * torch.tensor(self.avg_activation[self.timestep][self.layer], dtype=torch.float32)[idx]
            self._bias[key] = bias.to(device)
        return self._bias[key]

    def hook_fn(self, module, input, output):
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        routed = getattr(module, 'patterns', None) is not None and state.k is not None
        H, scores, gate = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                       want_scores=routed, want_gate=self.capture_gates)
        if routed:
            kk = int(K_FRACTION * module.k)
            bias = self._score_bias(state.n_experts, x.device)
            ops.router_topk(scores, kk, want_bits=False, H=H, expert_size=state.expert_size, score_bias=bias)
            if gate is not None:
                ops.router_topk(scores, kk, want_bits=False, H=gate, expert_size=state.expert_size, score_bias=bias)
        self._capture(gate, state, lead)
        self.update_time_layer()
        return self._finish(H, state, lead, x)
