"""MultiConceptRemoverWanda: one WandaRemoveNeuronsFast per concept + a union remover whose
per-(t, layer) mask is the OR of the selected concepts' masks
(reference neuron_receivers/multi_concept_remover.py:13-99).  The union runs on the packed
device bit words (moe_mask_union) instead of 816 dense numpy logical_or calls per prompt."""
import os

import numpy as np
import torch

from moe_b200 import ops
from moe_b200.sd_modules import GEGLU
from neuron_receivers.remove_wanda_neurons_fast import WandaRemoveNeuronsFast


class MultiConceptRemoverWanda:
    def __init__(self, root, seed, T, n_layers, replace_fn=GEGLU, keep_nsfw=False, remove_timesteps=None,
                 weights_shape=None, concepts_to_remove=None, wanda_thr=0.05, **kw):
        self.concepts_to_remove = concepts_to_remove
        self.removers = {}
        self.seed = seed
        self.T = T
        self.n_layers = n_layers
        for concept in concepts_to_remove:
            thr = wanda_thr[concept] if isinstance(wanda_thr, dict) else wanda_thr
            path_expert_indx = os.path.join((root % (seed, concept)), f'skilled_neuron_wanda/{thr}')
            self.removers[concept] = WandaRemoveNeuronsFast(
                seed=seed, path_expert_indx=path_expert_indx, T=T, n_layers=n_layers, replace_fn=replace_fn,
                keep_nsfw=keep_nsfw, remove_timesteps=remove_timesteps, weights_shape=weights_shape, **kw)
        # the union remover's masks are rebuilt per prompt
        self.union_neuron_remover = WandaRemoveNeuronsFast(
            seed=seed, path_expert_indx=None, T=T, n_layers=n_layers, replace_fn=replace_fn, keep_nsfw=keep_nsfw,
            remove_timesteps=remove_timesteps, weights_shape=weights_shape, **kw)

    def reset_union_remover(self):
        self.union_neuron_remover.reset_time_layer()
        self.union_neuron_remover.invalidate()

    def handle_multiple_concepts(self, concepts, device='cuda', column_perms=None):
        """union[t][l] = OR_c M_c[t][l] on packed bits."""
        self.reset_union_remover()
        for i in range(self.T):
            for j in range(self.n_layers):
                perm = None if column_perms is None else column_perms[j]
                acc = None
                for c in concepts:
                    self.removers[c].reset_time_layer()
                    b = self.removers[c].mask_bits(i, j, device, perm)
                    acc = b.clone() if acc is None else ops.mask_union(acc, b, out=acc)
                self.union_neuron_remover.set_mask_bits(i, j, acc)

    def remove_concepts(self, model, prompt, concepts):
        """Returns (original output, output with removal, [single-concept outputs] | None).
        (The reference pastes PIL images side by side; composing images is outside the hot path.)"""
        if len(concepts) == 0:
            return model(prompt).images[0], None, None
        device = getattr(model, 'device', 'cuda')
        singles = []
        if len(concepts) > 1:
            self.handle_multiple_concepts(concepts, device)
            self.union_neuron_remover.reset_time_layer()
            out_removal, _ = self.union_neuron_remover.observe_activation(model, prompt)
            for c in concepts:
                self.removers[c].reset_time_layer()
                im_, _ = self.removers[c].observe_activation(model, prompt)
                singles.append(im_)
        else:
            self.removers[concepts[0]].reset_time_layer()
            out_removal, _ = self.removers[concepts[0]].observe_activation(model, prompt)
        torch.manual_seed(self.seed)
        np.random.seed(self.seed)
        out_pre = model(prompt).images[0]
        return out_pre, out_removal, (singles if singles else None)
