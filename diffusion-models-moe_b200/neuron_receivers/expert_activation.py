"""ExpertPredictivity: per-expert max over tokens of the expert score, with running mean/std
across prompts; the output is the UNMASKED GEGLU (reference neuron_receivers/expert_activation.py:8-63).

The reference copies the [E] vector to the host on every layer call; here the column max lands
in a device buffer [T, n_layers, E_max] and the host statistics are updated once per prompt
(on `flush()`, called by observe_activation / reset / any read of `max_gate` or `predictivity`)."""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from moe_b200.stats import StatMeter
from neuron_receivers.base_receiver import BaseNeuronReceiver


class ExpertPredictivity(BaseNeuronReceiver):
    def __init__(self, seed, T, n_layers, keep_nsfw=False, **kw):
        kw.setdefault('capture_gates', False)
        super(ExpertPredictivity, self).__init__(seed, keep_nsfw=keep_nsfw, **kw)
        self.T = T
        self.n_layers = n_layers
        self._predictivity = StatMeter(T, n_layers)
        self._max_gate = {t: {l: [] for l in range(n_layers)} for t in range(T)}
        self._buf = None
        self._pending = []  # (timestep, layer, n_experts) written since the last flush
        self.timestep = 0
        self.layer = 0
        self.sample_id = 0

    def update_time_layer(self):
        if self.layer == self.n_layers - 1:   # expert_activation.py:28 hard-codes 15
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset_time_layer(self):
        self.flush()
        self.timestep = 0
        self.layer = 0

    def reset(self):
        self.flush()
        self._max_gate = {t: {l: [] for l in range(self.n_layers)} for t in range(self.T)}
        self.timestep = 0
        self.layer = 0

    def flush(self):
        """One D2H copy for everything recorded since the last flush; updates max_gate + Welford stats."""
        if not self._pending:
            return
        host = self._buf.cpu().numpy()
        for (t, l, E) in self._pending:
            v = host[t, l, :E].copy()
            self._max_gate[t][l] = v
            self._predictivity.update(v, t, l)
        self._pending = []

    @property
    def max_gate(self):
        self.flush()
        return self._max_gate

    @property
    def predictivity(self):
        self.flush()
        return self._predictivity

    def hook_fn(self, module, input, output):
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        routed = getattr(module, 'patterns', None) is not None
        H, scores, _ = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size,
                                    state.act, want_scores=routed)
        if routed:
            E = state.n_experts
            if self._buf is None or self._buf.shape[-1] < E:
                self.flush()
                self._buf = torch.empty(self.T, self.n_layers, max(E, 256), dtype=torch.float32, device=x.device)
            if any(p[0] == self.timestep and p[1] == self.layer for p in self._pending):
                self.flush()  # same cell written twice before a flush (new prompt without reset)
            cell = self._buf[self.timestep, self.layer, :E]
            cell.fill_(float('-inf'))
            ops.colmax(scores, out=cell)
            self._pending.append((self.timestep, self.layer, E))
        self.update_time_layer()
        return self._finish(H, state, lead, x)

    def observe_activation(self, model, ann, bboxes=None):
        out = super().observe_activation(model, ann, bboxes)
        self.flush()
        return out

    def test(self, model):
        return True
