"""Wanda scoring and union-over-timesteps baking on the device.

    score_masks(w2, norms_base, norms_adj, ratio)      reference modularity/wanda.py:140-173
    union_over_time(mask_bits, select_ratio)           reference benchmarks/save_union_over_time.py:189-211
    bake(linear, union_bits)                           reference benchmarks/save_union_over_time.py:219-227

Masks are bit words over the row-major [d, h] weight (the layout `moe_mask_weights` consumes); `to_csr` converts one
back to the scipy CSR matrix the reference pickles (`timestep_{t}_layer_{l}.pkl`)."""
import numpy as np
import torch

from moe_b200 import ops


def _weight_in_original_order(w2, column_perm):
    """`w2` is a weight [d, h] or the ff.net.2 module itself.  A module that `modify_ffn` packed in place carries
    `_moe_column_perm` (packed column j holds original neuron perm[j]); its weight is translated back so that
    the masks come out in the ORIGINAL neuron order, like the norms of the Wanda receiver and the reference's pickles."""
    if isinstance(w2, torch.nn.Module):
        column_perm = getattr(w2, '_moe_column_perm', None) if column_perm is None else column_perm
        w2 = w2.weight
    w = w2.detach()
    if column_perm is not None:
        inv = torch.empty_like(column_perm)
        inv[column_perm] = torch.arange(column_perm.numel())
        w = w[:, inv.to(w.device)]
    return w


def score_masks(w2, norms_base, norms_adj, ratio: float, column_perm=None) -> torch.Tensor:
    """w2 [d, h] (any float dtype, on the GPU) or the ff.net.2 module; norms_* = sequence over timesteps of f32 [h]
    column norms in original neuron order (Wanda receiver `get_column_norms()[t][l]`).  Returns int32 [T, d*h/32]
    mask bits over the row-major [d, h] weight in ORIGINAL column order."""
    w2 = _weight_in_original_order(w2, column_perm)
    d, h = w2.shape
    k = int(ratio * h)
    w = w2.to(torch.bfloat16).contiguous()
    out = torch.empty(len(norms_adj), d * h // 32, dtype=torch.int32, device=w.device)
    for t, (nb, na) in enumerate(zip(norms_base, norms_adj)):
        ops.wanda_score_mask(w, nb.to(w.device, torch.float32).contiguous(), na.to(w.device, torch.float32).contiguous(), k,
                             out=out[t])
    return out


def union_over_time(mask_bits: torch.Tensor, select_ratio: float) -> torch.Tensor:
    """mask_bits int32 [T, n_words] -> int32 [n_words]: kept where set at more than select_ratio * T timesteps."""
    return ops.mask_vote(mask_bits.contiguous(), select_ratio * mask_bits.shape[0])


@torch.no_grad()
def bake(linear: torch.nn.Module, union_bits: torch.Tensor) -> None:
    """ff.net.2 weight *= (1 - mask), in place (the baked "union-timesteps" checkpoint of the reference).
    `union_bits` are in original column order; a Linear packed in place gets them through its column permutation."""
    w = linear.weight
    perm = getattr(linear, '_moe_column_perm', None)
    if perm is not None:
        d, h = w.shape
        dense = torch.from_numpy(to_dense(union_bits, d, h).astype(np.uint8))[:, perm]
        union_bits = ops.mask_pack(dense.to(w.device).contiguous())
    masked = ops.mask_weights(w.detach().to(torch.bfloat16).contiguous(), union_bits)
    w.copy_(masked.to(w.dtype))      # (not through .data: bumps the version counter, packed copies are refreshed)


def to_dense(bits: torch.Tensor, d: int, h: int) -> np.ndarray:
    words = bits.detach().cpu().numpy().view(np.uint32)
    return np.unpackbits(words.view(np.uint8), bitorder="little")[: d * h].reshape(d, h).astype(int)


def to_csr(bits: torch.Tensor, d: int, h: int):
    import scipy.sparse
    return scipy.sparse.csr_matrix(to_dense(bits, d, h))
