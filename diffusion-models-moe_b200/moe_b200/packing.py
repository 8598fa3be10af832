"""One-time weight pre-pack for a MoEfied FFN.

The reference keeps experts as a one-hot `patterns[E, h]` matrix over arbitrarily ordered
neurons (moefication/helper.py:48-62).  The GEGLU FFN is invariant under a permutation of its
inner dimension applied consistently to the value rows and gate rows of W1 (and b1) and to the
columns of W2, so we sort neurons by expert once: afterwards expert e owns the contiguous
neurons [e*es, (e+1)*es), the score matmul against `patterns` becomes a segment sum in the
up-projection epilogue, and the neuron mask becomes E bits per token.  Expert ids are unchanged.
"""
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch


@dataclass
class ExpertLayout:
    n_experts: int
    expert_size: int
    perm: torch.Tensor      # int64 [h]: packed position j holds original neuron perm[j]
    inv_perm: torch.Tensor  # int64 [h]: original neuron n sits at packed position inv_perm[n]

    @property
    def hidden(self) -> int:
        return self.n_experts * self.expert_size

    @property
    def is_identity(self) -> bool:
        cached = self.__dict__.get("_identity")
        if cached is None:
            cached = bool(torch.equal(self.perm, torch.arange(self.hidden)))
            self.__dict__["_identity"] = cached
        return cached

    @staticmethod
    def from_labels(labels: Sequence[int]) -> "ExpertLayout":
        lab = np.asarray(labels, dtype=np.int64)
        n_experts = int(lab.max()) + 1                       # helper.py:51 `cluster_num = max(labels)+1`
        sizes = np.bincount(lab, minlength=n_experts)
        if not np.all(sizes == sizes[0]):
            raise ValueError(
                "unbalanced experts are not supported by the CUDA path (the reference's balanced k-means, "
                f"moefication/moe_utils.py:97-107, always yields equal sizes); got sizes {sorted(set(sizes.tolist()))}")
        perm = np.argsort(lab, kind="stable")
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        return ExpertLayout(n_experts, int(sizes[0]), torch.from_numpy(perm), torch.from_numpy(inv))

    @staticmethod
    def contiguous(n_experts: int, expert_size: int) -> "ExpertLayout":
        ar = torch.arange(n_experts * expert_size)
        return ExpertLayout(n_experts, expert_size, ar, ar.clone())

    def patterns(self, dtype=torch.float32, device=None, packed: bool = True) -> torch.Tensor:
        """one-hot [E, h] in packed (block) or original neuron order -- API compatibility with
        `module.patterns` (helper.py:55-59); the kernels never read it."""
        h = self.hidden
        p = torch.zeros(self.n_experts, h, dtype=dtype)
        owner = torch.arange(h) // self.expert_size
        p[owner, torch.arange(h)] = 1
        if not packed:
            p = p[:, self.inv_perm]
        return p.to(device) if device is not None else p


@dataclass
class PackedFFN:
    layout: ExpertLayout
    w1p: torch.Tensor            # bf16 [2h, d]
    b1p: Optional[torch.Tensor]  # f32 [2h]
    w2p: Optional[torch.Tensor]  # bf16 [d, h]
    b2: Optional[torch.Tensor]   # f32 [d]


def pack_ffn(layout: ExpertLayout, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: Optional[torch.Tensor] = None,
             b2: Optional[torch.Tensor] = None, device=None) -> PackedFFN:
    """Permute + cast the FFN parameters into the kernels' layout (bf16 weights, f32 biases)."""
    h = layout.hidden
    assert w1.shape[0] == 2 * h, (w1.shape, h)
    dev = device if device is not None else w1.device
    if torch.equal(layout.perm, torch.arange(h)):
        # already in packed order: alias bf16 parameters instead of copying them
        w1s, b1s, w2s = w1.detach(), (None if b1 is None else b1.detach()), (None if w2 is None else w2.detach())
    else:
        perm = layout.perm.to(w1.device)
        rows = torch.cat([perm, perm + h])
        w1s = w1.detach()[rows]
        b1s = None if b1 is None else b1.detach()[rows]
        w2s = None if w2 is None else w2.detach()[:, perm.to(w2.device)]
    w1p = w1s.to(device=dev, dtype=torch.bfloat16).contiguous()
    b1p = None if b1s is None else b1s.to(device=dev, dtype=torch.float32).contiguous()
    w2p = None if w2s is None else w2s.to(device=dev, dtype=torch.bfloat16).contiguous()
    b2f = None if b2 is None else b2.detach().to(device=dev, dtype=torch.float32).contiguous()
    return PackedFFN(layout, w1p, b1p, w2p, b2f)


def bits_from_expert_list(experts: Sequence[int], n_experts: int) -> torch.Tensor:
    """expert id list (remove_skilled_experts.py:18 JSON) -> int32 [ceil(E/32)] bit words (CPU tensor)."""
    words = np.zeros((n_experts + 31) // 32, dtype=np.uint32)
    for e in experts:
        e = int(e)
        if not 0 <= e < n_experts:
            raise ValueError(f"expert id {e} out of range [0, {n_experts})")
        words[e >> 5] |= np.uint32(1 << (e & 31))
    return torch.from_numpy(words.view(np.int32).copy())


def bits_to_sets(bits: torch.Tensor, n_experts: int):
    """int32 [T, W] -> list of python sets (test helper; CPU)."""
    arr = bits.detach().cpu().numpy().view(np.uint32)
    out = []
    for row in arr:
        s = set()
        for w, word in enumerate(row):
            word = int(word)
            while word:
                low = word & -word
                s.add(32 * w + low.bit_length() - 1)
                word ^= low
        out.append(s)
    return out
