"""FrequencyMeasure: MoE forward + per-(timestep, layer) expert-selection frequency
(reference neuron_receivers/frequency_measure.py:7-64).

The reference pulls `labels[0]` to the host every call and loops over the S tokens in Python
(`counter[labels[i]] += 1/S`).  Here the router kernel accumulates integer counts into one
device-resident int64 tensor `hist[T, n_layers, E_max]` (no sync in the sampling loop);
`label_counter` materialises the reference's dict-of-float64 view (count / S) on access, and
`int_counts()` exposes the exact integers (what is all-reduced across GPUs).
"""
import numpy as np
import torch

from neuron_receivers.base_receiver import BaseNeuronReceiver
from neuron_receivers.moefy import routed_ffn


class FrequencyMeasure(BaseNeuronReceiver):
    def __init__(self, seed, T, n_layers, experts_per_layer, layer_names, *, count_rows='row0', device=None, **kw):
        kw.setdefault('capture_gates', False)  # the reference FrequencyMeasure never stores gates
        super(FrequencyMeasure, self).__init__(seed, **kw)
        self.T = T
        self.n_layers = n_layers
        self.experts_per_layer = experts_per_layer
        self.layer_names = layer_names
        if count_rows not in ('row0', 'all'):
            raise ValueError("count_rows must be 'row0' (reference: first batch row only) or 'all'")
        self.count_rows = count_rows
        self._n_experts = [int(experts_per_layer[layer_names[i]]) for i in range(n_layers)]
        self._e_max = max(self._n_experts)
        self._device = device
        self._hist = None
        self._seq_len = {}
        self.timestep = 0
        self.layer = 0
        self.sample_id = 0

    # -- (timestep, layer) state machine: frequency_measure.py:24-33 (hard-codes 15 == n_layers-1 for SD) --
    def update_time_layer(self):
        if self.layer == self.n_layers - 1:
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset_time_layer(self):
        self.timestep = 0
        self.layer = 0

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()
        self._seq_len = {}
        self.reset_time_layer()

    def _hist_tensor(self, device):
        if self._hist is None:
            self._hist = torch.zeros(self.T, self.n_layers, self._e_max, dtype=torch.int64, device=device)
        return self._hist

    def hook_fn(self, module, input, output):
        x = input[0]
        bsz, seq_len = x.shape[0], x.shape[1]
        hist = None
        rows = (0, 0)
        if getattr(module, 'patterns', None) is not None:
            E = module.patterns.shape[0]
            hist = self._hist_tensor(x.device)[self.timestep, self.layer, :E]
            rows = (0, seq_len) if self.count_rows == 'row0' else (0, bsz * seq_len)
            self._seq_len[(self.timestep, self.layer)] = seq_len
        out, _ = routed_ffn(self, module, x, hist=hist, count_rows=rows)
        self.update_time_layer()
        return out

    # -- results ------------------------------------------------------------------------------------
    def int_counts(self) -> torch.Tensor:
        """int64 [T, n_layers, E_max] on the device: #(token, slot) selections per expert."""
        if self._hist is None:
            dev = self._device or ('cuda' if torch.cuda.is_available() else 'cpu')
            self._hist = torch.zeros(self.T, self.n_layers, self._e_max, dtype=torch.int64, device=dev)
        return self._hist

    @property
    def label_counter(self):
        """{t: {layer: float64[E]}} with value = selections / seq_len, as the reference accumulates
        (frequency_measure.py:57).  Synchronises."""
        counts = self.int_counts().cpu().numpy()
        out = {}
        for t in range(self.T):
            out[t] = {}
            for l in range(self.n_layers):
                s = self._seq_len.get((t, l))
                c = counts[t, l, :self._n_experts[l]].astype(np.float64)
                out[t][l] = c / s if s else np.zeros(self._n_experts[l])
        return out

    def all_reduce(self, group=None):
        """Sum the integer histogram over the ranks of `group` (prompts sharded across GPUs)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.int_counts(), op=dist.ReduceOp.SUM, group=group)
        return self.int_counts()
