"""Offline expert construction (reference moefication/moe_utils.py:66-107, moefy_sd_model.py:19-43):
split the neurons of every GEGLU FFN into equally sized experts by clustering the L2-normalised rows of the GATE half of
the up-projection weight (`proj.weight[h:2h]`, moe_utils.py:68-72), and save one label list per FFN in the reference's
format (`torch.save(list[int])` under `{res_path}/param_split/<ffn>.proj.weight`, moe_utils.py:54-61), which
`helper.modify_ffn_to_experts` reads back.

The reference calls `k_means_constrained.KMeansConstrained(size_min = size_max = expert_size, random_state=0)` (Lloyd
iterations whose assignment step is a min-cost flow).  That package is not vendored and not installable here, and its
result depends on its own RNG stream and flow solver, so the LABELS are not reproducible bit for bit -- parity for this
row is "same contract": every expert has exactly `expert_size` neurons, the split is deterministic for a seed, and the
clustering objective (within-cluster inertia on the unit sphere) is optimised by the same alternation.  The assignment
step here is the regret-ordered greedy used for balanced k-means: neurons are visited in decreasing order of
(second-best - best) distance and take the nearest expert that still has room.  This is a once-per-model offline step;
the distance matrices are torch GEMMs on whatever device the weight lives on."""
import os
from typing import List, Optional

import numpy as np
import torch


def _balanced_assign(dist: np.ndarray, size: int) -> np.ndarray:
    """dist [n, E] -> labels [n] with exactly `size` points per cluster (n == E * size)."""
    n, E = dist.shape
    part = np.partition(dist, 1, axis=1) if E > 1 else np.concatenate([dist, dist], 1)
    regret = part[:, 1] - part[:, 0]
    order = np.argsort(-regret, kind="stable")
    room = np.full(E, size, dtype=np.int64)
    labels = np.empty(n, dtype=np.int64)
    pref = np.argsort(dist, axis=1, kind="stable")
    for i in order:
        for e in pref[i]:
            if room[e] > 0:
                labels[i] = e
                room[e] -= 1
                break
    return labels


@torch.no_grad()
def balanced_kmeans(rows: torch.Tensor, expert_size: int, seed: int = 0, max_iter: int = 30) -> List[int]:
    """rows [n, d] (any float dtype / device) -> list of n labels, every label exactly `expert_size` times."""
    n = rows.shape[0]
    if n % expert_size != 0:
        raise ValueError(f"{n} neurons are not divisible by the expert size {expert_size} (moe_utils.py:78)")
    E = n // expert_size
    x = torch.nn.functional.normalize(rows.detach().float(), dim=1)          # sklearn.preprocessing.normalize
    rs = np.random.RandomState(seed)
    centres = x[torch.from_numpy(rs.permutation(n)[:E]).to(x.device)].clone()
    labels = None
    for _ in range(max_iter):
        dist = (2.0 - 2.0 * (x @ centres.t())).cpu().numpy()                  # squared distance of unit vectors
        new = _balanced_assign(dist, expert_size)
        if labels is not None and np.array_equal(new, labels):
            break
        labels = new
        idx = torch.from_numpy(labels).to(x.device)
        sums = torch.zeros_like(centres).index_add_(0, idx, x)
        centres = torch.nn.functional.normalize(sums, dim=1)
    return [int(v) for v in labels]


def inertia(rows: torch.Tensor, labels) -> float:
    """Within-cluster sum of squared distances on the unit sphere (the k-means objective)."""
    x = torch.nn.functional.normalize(rows.detach().float().cpu(), dim=1)
    lab = torch.as_tensor(labels, dtype=torch.long)
    E = int(lab.max()) + 1
    centres = torch.zeros(E, x.shape[1]).index_add_(0, lab, x)
    centres = centres / torch.bincount(lab, minlength=E).clamp(min=1).unsqueeze(1)
    return float(((x - centres[lab]) ** 2).sum())


def split_ffn_weight(proj_weight: torch.Tensor, expert_size: int, seed: int = 0) -> List[int]:
    """`proj.weight` [2h, d] of a GEGLU -> expert labels of its h neurons (clusters the GATE half, moe_utils.py:68-72)."""
    h = proj_weight.shape[0] // 2
    return balanced_kmeans(proj_weight[h:], expert_size, seed)


def save_labels(labels: List[int], res_path: str, ffn_weight_name: str) -> str:
    """`{res_path}/param_split/<ffn>.proj.weight` = torch.save(list of labels) (moe_utils.py:54-61)."""
    folder = os.path.join(res_path, "param_split")
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, ffn_weight_name)
    torch.save([int(v) for v in labels], path)
    return path


def moefy_sd_model(model, res_path: str, expert_size: int = 20, seed: int = 0, geglu_type: Optional[type] = None):
    """moefy_sd_model.main (moefy_sd_model.py:19-43): one label file per GEGLU FFN of `model.unet`.  Returns
    {ffn weight name: labels}."""
    if geglu_type is None:
        from moe_b200.sd_modules import GEGLU as geglu_type
    out = {}
    for name, module in model.unet.named_modules():
        if 'ff.net' in name and isinstance(module, geglu_type):
            ffn_name = name + '.proj.weight'
            labels = split_ffn_weight(module.proj.weight, expert_size, seed)
            save_labels(labels, res_path, ffn_name)
            out[ffn_name] = labels
    return out
