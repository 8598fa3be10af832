"""MoE set-up for the FFNs of a diffusion UNet (reference moefication/helper.py:48-96).

`modify_ffn_to_experts(model, args)` has the reference's signature and return value
`(model, layer_names, num_experts_per_ffn)` and attaches `patterns [E, h]` and `k` to every GEGLU.
In addition it sorts each FFN's neurons by expert ONCE (see moe_b200/packing.py) so the kernels
see contiguous experts, and stores the device state on `module._moe_state`."""
import os

import torch

from moe_b200.ffn import attach_state, find_down_proj
from moe_b200.packing import ExpertLayout
from moe_b200.sd_modules import GEGLU


def load_labels(path):
    """label list saved by the offline split (reference moefication/moe_utils.py:54-61: torch.save(list))."""
    labels = torch.load(path, weights_only=False)
    return [int(v) for v in labels]


def modify_ffn(ffn, path_or_labels, k, down=None, permute_model_weights=True):
    """Attach experts to one GEGLU (reference helper.py:48-62).  `k` is the RATIO of selected
    experts; the module gets `k = int(E * ratio)` exactly as the reference computes it."""
    assert isinstance(ffn, GEGLU)
    labels = load_labels(path_or_labels) if isinstance(path_or_labels, (str, os.PathLike)) else path_or_labels
    layout = ExpertLayout.from_labels(labels)
    ffn.k = int(layout.n_experts * k)
    state = attach_state(ffn, layout, ffn.k, down, permute_model_weights)
    w = ffn.proj.weight
    # module.patterns[i, j] = 1 iff neuron j (in the module's current neuron order) is in expert i
    ffn.patterns = layout.patterns(dtype=w.dtype, device=w.device, packed=state.weights_permuted_in_model)
    ffn.expert_size = layout.expert_size
    ffn.neuron_perm = layout.perm
    if down is not None:
        down._moe_column_perm = layout.perm if state.weights_permuted_in_model else None
    return state


def modify_ffn_to_experts(model, args, labels_by_name=None, permute_model_weights=True):
    """Reference helper.py:65-78.  `args` needs `.res_path` and `.moefication['topk_experts']`.
    `labels_by_name` (optional, {ffn weight name: label list}) bypasses the label files."""
    num_experts_per_ffn = {}
    layer_names = []
    for name, module in model.unet.named_modules():
        if 'ff.net' in name and isinstance(module, GEGLU):
            ffn_name = name + '.proj.weight'
            src = labels_by_name[ffn_name] if labels_by_name is not None else \
                os.path.join(args.res_path, 'param_split', ffn_name)
            modify_ffn(module, src, args.moefication['topk_experts'], find_down_proj(model.unet, name),
                       permute_model_weights)
            layer_names.append(ffn_name)
            num_experts_per_ffn[ffn_name] = module.patterns.shape[0]
    layer_names.sort()   # sorted order == firing order down -> mid -> up (helper.py:76-77)
    return model, layer_names, num_experts_per_ffn


def initialise_expert_counter(model, timesteps=51):
    """Reference helper.py:80-96."""
    import numpy as np
    expert_counter = {i: {} for i in range(timesteps)}
    ffn_names_list = []
    for name, module in model.unet.named_modules():
        if 'ff.net' in name and isinstance(module, GEGLU):
            ffn_name = name + '.proj.weight'
            for t in range(timesteps):
                expert_counter[t][ffn_name] = np.zeros(module.patterns.shape[0])
            ffn_names_list.append(ffn_name)
    ffn_names_list.sort()
    return expert_counter, ffn_names_list


def average_expert_counters(per_image_counters, layer_names, timesteps):
    """freq_expert_select.main accumulation (reference moefication/freq_expert_select.py:43-64):
    {t: {ffn_name: [E floats]}} = mean over images of the per-image label_counter."""
    n = len(per_image_counters)
    out = {t: {nm: [0.0] * len(per_image_counters[0][t][i]) for i, nm in enumerate(layer_names)}
           for t in range(timesteps)}
    for counter in per_image_counters:
        for t in range(timesteps):
            for i, nm in enumerate(layer_names):
                row = out[t][nm]
                src = counter[t][i]
                for e in range(len(row)):
                    row[e] += src[e] / n
    return out
