"""GetExperts: per (timestep, layer) the top-k experts of the expert score AVERAGED over tokens -- all tokens, or the
bounding-box positions of every batch row (reference neuron_receivers/get_experts.py:8-83).  The output is the
unmasked GEGLU.

Device side: K1 (scores), a column-sum kernel with an optional periodic row mask (moe_colsum_f32); the k labels of
the [E] mean vector are taken on the host in torch.topk's descending-score order, exactly the list the reference
stores (it does the same D2H with .tolist())."""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import as_tokens, get_state
from neuron_receivers.base_receiver import BaseNeuronReceiver


class GetExperts(BaseNeuronReceiver):
    def __init__(self, seed, T, n_layers, experts_per_layer, layer_names, keep_nsfw=False, **kw):
        kw.setdefault('capture_gates', False)
        super(GetExperts, self).__init__(seed, keep_nsfw=keep_nsfw, **kw)
        self.T = T
        self.n_layers = n_layers
        self.experts_per_layer = experts_per_layer
        self.layer_names = layer_names
        self.label_counter = {}
        self.freq_counter = {}
        self.mean_score = {}
        self.reset()
        self.sample_id = 0

    def update_time_layer(self):
        if self.layer == self.n_layers - 1:     # get_experts.py:31 hard-codes 15
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset_time_layer(self):
        self.timestep = 0
        self.layer = 0

    def reset(self):
        for t in range(self.T):
            self.label_counter[t] = {}
            self.freq_counter[t] = {}
            self.mean_score[t] = {}
            for i in range(self.n_layers):
                self.freq_counter[t][i] = np.zeros(self.experts_per_layer[self.layer_names[i]])
                self.label_counter[t][i] = []
        self.reset_time_layer()

    def hook_fn(self, module, input, output):
        x = input[0]
        state = get_state(module)
        lead = x.shape[:-1]
        routed = getattr(module, 'patterns', None) is not None
        H, scores, _ = ops.geglu_up(as_tokens(x), state.w1p, state.b1p, state.n_experts, state.expert_size, state.act,
                                    want_scores=routed)
        if routed:
            S = x.shape[-2]
            n_rows = scores.shape[0]
            row_mask = None
            bb = getattr(module, 'bounding_box', None)
            if bb is not None:
                keep = np.zeros(S, dtype=np.uint8)
                try:
                    keep[np.asarray(bb, dtype=np.int64)] = 1          # gate[:, bounding_box, :], get_experts.py:66
                    row_mask = torch.from_numpy(keep).to(x.device)
                    n_rows = (n_rows // S) * int(keep.sum())
                except (IndexError, ValueError):                      # the reference's bare `except`: whole sequence
                    row_mask = None
            sums = ops.colsum(scores, row_mask=row_mask)
            mean = (sums / max(n_rows, 1)).cpu()
            self.mean_score[self.timestep][self.layer] = mean.numpy()
            self.label_counter[self.timestep][self.layer] = torch.topk(mean, k=module.k, dim=-1)[1].reshape(-1).tolist()
        self.update_time_layer()
        return self._finish(H, state, lead, x)
