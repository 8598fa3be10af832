"""Wanda receiver: column norms of the row-normalised GEGLU output per (timestep, layer), accumulated over prompts
(reference neuron_receivers/wanda_receiver.py:9-57, utils.py:321-370 ColumnNormCalculator / TimeLayerColumnNorm).

The reference copies every [B*S, h] activation to the host, normalises its rows and updates sqrt(old^2 + new^2).
Here one kernel (moe_rownorm_colsumsq_bf16) adds the squared column norms of the call into a device-resident
[T, n_layers, h_max] accumulator; the root is taken when the norms are read."""
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from moe_b200.sd_modules import GEGLU
from neuron_receivers.base_receiver import BaseNeuronReceiver


class TimeLayerColumnNorm:
    """Device-resident stand-in for utils.TimeLayerColumnNorm: same `get_column_norms()` / `save()` results."""

    def __init__(self, T, n_layers):
        self.T = T
        self.n_layers = n_layers
        self.sumsq = {}          # (t, layer) -> f32 [h] device tensor, in the kernels' packed neuron order
        # (t, layer) -> packed -> ORIGINAL order index; get_column_norms() returns original order (the order of the
        # reference's artefacts and of the masks WandaRemoveNeuronsFast consumes)
        self.perm = {}

    def cell(self, t, layer, h, device):
        key = (t, layer)
        if key not in self.sumsq:
            self.sumsq[key] = torch.zeros(h, dtype=torch.float32, device=device)
        return self.sumsq[key]

    def get_column_norms(self):
        results = {}
        for t in range(self.T):
            results[t] = {}
            for i in range(self.n_layers):
                v = self.sumsq.get((t, i))
                if v is None:
                    results[t][i] = torch.tensor([])
                    continue
                v = torch.sqrt(v)
                inv = self.perm.get((t, i))
                results[t][i] = (v if inv is None else v[inv]).cpu()
        return results

    def save(self, path):
        torch.save(self.get_column_norms(), path)


class Wanda(BaseNeuronReceiver):
    def __init__(self, seed, T, n_layers, replace_fn=GEGLU, keep_nsfw=False, hook_module='unet', **kw):
        kw.setdefault('capture_gates', False)
        super(Wanda, self).__init__(seed, replace_fn, keep_nsfw, hook_module, **kw)
        if hook_module != 'unet' or replace_fn != GEGLU:
            raise NotImplementedError("only the UNet GEGLU FFN path is implemented natively")
        self.T = T
        self.n_layers = n_layers
        self.predictivity = TimeLayerColumnNorm(T, n_layers)
        self.timestep = 0
        self.layer = 0

    def update_time_layer(self):
        if self.layer == self.n_layers - 1:
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset_time_layer(self):
        self.timestep = 0
        self.layer = 0

    def hook_fn(self, module, input, output):
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        H, _, _ = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                               want_scores=False)
        ops.rownorm_colsumsq(H, out=self.predictivity.cell(self.timestep, self.layer, H.shape[-1], H.device))
        if not state.layout.is_identity:
            self.predictivity.perm[(self.timestep, self.layer)] = state.layout.inv_perm.to(H.device)
        self.update_time_layer()
        return self._finish(H, state, lead, x)
