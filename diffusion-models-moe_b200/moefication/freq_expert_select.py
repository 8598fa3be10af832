"""Expert-selection frequency driver (reference moefication/freq_expert_select.py:20-72).

    run(model, prompts, seed, timesteps, n_layers, ffn_names, num_experts_per_ffn, topk, save_path)

For every prompt: reset the receiver, run the hooked pipeline, add `label_counter / num_images` into
`expert_counter[t][ffn_name][expert]`; finally write `expert_counter_{topk}.json` in the reference's schema
`{t: {ffn_name: [E floats]}}`.  With `world_size > 1` each rank takes prompts `rank::world_size` and the
integer histograms are summed with one all-reduce before the averages are formed."""
import json
import os

import numpy as np
import torch

from neuron_receivers import FrequencyMeasure


def run(model, prompts, seed, timesteps, n_layers, ffn_names, num_experts_per_ffn, topk, save_path=None,
        rank=0, world_size=1, count_rows='row0'):
    receiver = FrequencyMeasure(seed, timesteps, n_layers, num_experts_per_ffn, ffn_names, count_rows=count_rows)
    num_images = len(prompts)
    seq_len = {}
    total = None
    for i in range(rank, num_images, world_size):
        torch.manual_seed(seed)
        np.random.seed(seed)
        receiver.reset()
        receiver.observe_activation(model, prompts[i])
        seq_len.update(receiver._seq_len)
        counts = receiver.int_counts().clone()
        total = counts if total is None else total + counts
    if total is None:
        total = torch.zeros_like(receiver.int_counts())
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    host = total.cpu().numpy()
    expert_counter = {}
    for t in range(timesteps):
        expert_counter[t] = {}
        for l, name in enumerate(ffn_names):
            E = num_experts_per_ffn[name]
            s = seq_len.get((t, l))
            vals = host[t, l, :E].astype(np.float64) / (s * num_images) if s else np.zeros(E)
            expert_counter[t][name] = vals.tolist()
    if save_path is not None and rank == 0:
        os.makedirs(save_path, exist_ok=True)
        with open(os.path.join(save_path, f'expert_counter_{topk}.json'), 'w') as f:
            json.dump(expert_counter, f)
    return expert_counter
