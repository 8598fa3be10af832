"""CPU oracle for the MoEfied GEGLU feed-forward hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as
the *checker* / the CPU arm.  Nothing under `diffusion-models-moe_b200/` imports it; the
product path fails loudly when the CUDA library is missing.

What it is: a plain fp32 CPU restatement (torch-CPU for the floating-point tensor algebra --
the reference itself is torch -- and numpy for the integer counters) of what the reference's
forward hooks compute.  Every function cites the reference file:line it follows
(paths relative to the reference root).

Pinning: `oracle/gen_golden.py` imports the reference's own `neuron_receivers` UNMODIFIED
(through `oracle/shim/`, a stand-in for the uninstallable `diffusers`), runs its hook_fns on
seeded synthetic inputs, asserts this restatement reproduces them bit-for-bit
(`torch.equal`), and writes the reference outputs to `tests/golden/*.npz`.
`tests/test_oracle_golden.py` re-checks the restatement against those committed vectors
without needing /root/reference.  The reference ships no numeric tests of its own for this
path, so the reference code executed on synthetic inputs is the pin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ACT_GELU = 0  # exact erf GELU: upstream diffusers GEGLU.gelu == F.gelu(approximate='none')
ACT_RELU = 1  # sparsity/relufy_model.py:28-40 swaps module.gelu for relu


# --------------------------------------------------------------------------------------
# MoE set-up  (moefication/helper.py:48-62)
# --------------------------------------------------------------------------------------
def patterns_from_labels(labels: Sequence[int], dtype=torch.float32) -> torch.Tensor:
    """patterns[e, n] = 1 iff neuron n belongs to expert e; E = max(label)+1.
    (moefication/helper.py:50-59)"""
    lab = np.asarray(labels)
    n_experts = int(lab.max()) + 1
    rows = [lab == e for e in range(n_experts)]
    return torch.tensor(np.stack(rows).astype(np.float32)).to(dtype)


def topk_from_ratio(n_experts: int, ratio: float) -> int:
    """k = int(E * ratio) with Python float semantics (moefication/helper.py:61)."""
    return int(n_experts * ratio)


def balanced_labels(hidden: int, expert_size: int, seed: int = 0) -> np.ndarray:
    """Synthetic stand-in for the balanced k-means label file
    (moefication/moe_utils.py:97-107 writes one label per gate neuron, every expert exactly
    `expert_size` neurons): a seeded random permutation of the balanced assignment."""
    assert hidden % expert_size == 0  # moefication/moe_utils.py:78
    base = np.repeat(np.arange(hidden // expert_size), expert_size)
    return np.random.RandomState(seed).permutation(base)


# --------------------------------------------------------------------------------------
# The GEGLU pieces  (neuron_receivers/moefy.py:10-27, upstream GEGLU / FeedForward)
# --------------------------------------------------------------------------------------
def activation(gate: torch.Tensor, act: int) -> torch.Tensor:
    if act == ACT_GELU:
        return F.gelu(gate)
    if act == ACT_RELU:
        return F.relu(gate)
    raise ValueError(f"unknown activation {act}")


def geglu_up(x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], act: int = ACT_GELU
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """value, act(gate) of the up-projection.  The value half is rows [0,h) of W1 and the gate
    half rows [h,2h) (moefy.py:12 `.chunk(2, dim=-1)`; moefication/moe_utils.py:68-72)."""
    y = F.linear(x, w1, b1)
    value, gate = y.chunk(2, dim=-1)
    return value, activation(gate, act)


def expert_scores(gate_act: torch.Tensor, patterns: torch.Tensor) -> torch.Tensor:
    """score[t, e] = sum of activated gate over the neurons of expert e, for all B*S tokens
    (moefy.py:18-20)."""
    flat = gate_act.clone().view(-1, gate_act.shape[-1])
    return torch.matmul(flat, patterns.transpose(0, 1))


def route_topk(score: torch.Tensor, k: int) -> torch.Tensor:
    """indices of the k largest experts per token (moefy.py:21).  Order within a token's list is
    irrelevant downstream; compare as sets."""
    return torch.topk(score, k=k, dim=-1)[1]


def neuron_mask(labels: torch.Tensor, patterns: torch.Tensor) -> torch.Tensor:
    """union of the selected experts' one-hot rows (moefy.py:22)."""
    return F.embedding(labels, patterns).sum(-2)


def moefy_forward(x, w1, b1, patterns, k, act=ACT_GELU):
    """MOEFy.hook_fn (moefy.py:10-27).  Returns (H, labels[B,S,k], gate_masked, score)."""
    value, gate = geglu_up(x, w1, b1, act)
    bsz, seq_len, _ = gate.shape
    score = expert_scores(gate, patterns)
    labels = route_topk(score, k).view(bsz, seq_len, k)
    mask = neuron_mask(labels, patterns)
    gate = gate.clone()
    gate[mask == 0] = 0
    return value * gate, labels, gate, score


def down_proj(h: torch.Tensor, w2: torch.Tensor, b2: Optional[torch.Tensor]) -> torch.Tensor:
    """ff.net.2 of the upstream FeedForward: Linear(h, d); Dropout(p=0) in between is identity."""
    return F.linear(h, w2, b2)


# --------------------------------------------------------------------------------------
# Variants
# --------------------------------------------------------------------------------------
REMOVE_EXPERTS_LAST_TIMESTEP = 20  # hard-coded `self.timestep < 20`, remove_skilled_experts.py:32
REMOVED_NEURON_GATE = -0.17       # hard-coded, remove_skilled_neurons.py:39


def remove_experts_forward(x, w1, b1, patterns, k, removed: Sequence[int], timestep: int,
                           act=ACT_GELU):
    """RemoveExperts.hook_fn (remove_skilled_experts.py:24-55): listed experts get their pattern
    rows zeroed iff the list is non-empty and timestep < 20, so they score exactly 0, still
    compete in the top-k, and contribute no neurons."""
    pat = patterns.clone()
    if len(removed) > 0 and timestep < REMOVE_EXPERTS_LAST_TIMESTEP:
        pat[list(removed), :] = 0
    value, gate = geglu_up(x, w1, b1, act)
    bsz, seq_len, _ = gate.shape
    score = expert_scores(gate, pat)
    labels = route_topk(score, k).view(bsz, seq_len, k)
    mask = neuron_mask(labels, pat)
    gate = gate.clone()
    gate[mask == 0] = 0
    return value * gate, labels, gate, score


def remove_neurons_forward(x, w1, b1, neuron_flags: Sequence[float], act=ACT_GELU):
    """RemoveNeurons.hook_fn, GEGLU branch (remove_skilled_neurons.py:30-42): no routing; flagged
    neurons get gate := -0.17 AFTER the activation, at every timestep."""
    value, gate = geglu_up(x, w1, b1, act)
    gate = gate.clone()
    if len(neuron_flags) > 0:
        idx = torch.where(torch.tensor(neuron_flags) == 1)[0]
        gate[:, :, idx] = REMOVED_NEURON_GATE
    return value * gate, gate


def wanda_down_proj(h, w2, b2, weight_mask):
    """WandaRemoveNeuronsFast.linear_hook_fn (remove_wanda_neurons_fast.py:69-83):
    y = h (W2 * (1 - M))^T + b2 with M in {0,1}^{d x h}."""
    m = torch.as_tensor(np.asarray(weight_mask)).to(w2.dtype)
    return F.linear(h, w2 * (1 - m), b2)


def mask_union(*masks) -> np.ndarray:
    """MultiConceptRemoverWanda.handle_multiple_concepts (multi_concept_remover.py:43-53):
    elementwise OR folded over the concepts, cast to int."""
    out = np.zeros(np.asarray(masks[0]).shape)
    for m in masks:
        out = np.logical_or(out, np.asarray(m)).astype(int)
    return out


# --------------------------------------------------------------------------------------
# Counters and statistics
# --------------------------------------------------------------------------------------
def frequency_update(counter: np.ndarray, labels: torch.Tensor, seq_len: int) -> None:
    """FrequencyMeasure.hook_fn counter (frequency_measure.py:53-57): only batch row 0; each
    (token, selected expert) adds 1/seq_len in float64."""
    rows = labels[0, :, :].detach().cpu().numpy()
    for i in range(rows.shape[0]):
        counter[rows[i, :]] += (1.0 / seq_len)


def selection_counts(labels: torch.Tensor, n_experts: int, rows: slice = slice(0, 1)) -> np.ndarray:
    """Integer form of the same counter (what the CUDA histogram produces):
    count[e] = #(token, slot) pairs in batch rows `rows` that selected e."""
    flat = labels[rows].reshape(-1).cpu().numpy()
    return np.bincount(flat, minlength=n_experts).astype(np.int64)


def expert_predictivity(score: torch.Tensor) -> np.ndarray:
    """ExpertPredictivity.hook_fn (expert_activation.py:56-58): per-expert max of the score over
    ALL B*S tokens."""
    return torch.max(score, dim=0)[0].detach().cpu().numpy()


def neuron_predictivity(gate_act: torch.Tensor) -> np.ndarray:
    """NeuronPredictivity.hook_fn (predictivity.py:49): per-neuron max of act(gate) over tokens."""
    return torch.max(gate_act.reshape(-1, gate_act.shape[-1]), dim=0)[0].detach().cpu().numpy()


def get_experts_labels(x, w1, b1, patterns, k, bounding_box=None, act=ACT_GELU):
    """GetExperts.hook_fn (get_experts.py:50-83): top-k (torch.topk order: descending score) of the expert score
    AVERAGED over tokens -- all B*S tokens, or the bounding-box positions of every batch row -- and the unmasked
    GEGLU output.  Returns (H, labels list[int], mean score [E])."""
    v, g = geglu_up(x, w1, b1, act)
    gate = g if g.dim() == 3 else g.unsqueeze(0)
    h = gate.shape[-1]
    if bounding_box is not None:
        gate = gate[:, bounding_box, :]
    mean = torch.matmul(gate.reshape(-1, h), patterns.transpose(0, 1)).mean(0)
    labels = torch.topk(mean, k=k, dim=-1)[1]
    return v * g, labels.reshape(-1).tolist(), mean


def add_experts_forward(x, w1, b1, patterns, k, expert_idx: Sequence[int], std: Sequence[float], act=ACT_GELU):
    """AddExperts.hook_fn (add_skilled_experts.py:37-62): the scores of the listed experts are raised by
    5 * std[e] before a top-int(0.8 k) selection; masking as in MOEFy.  Returns (H, labels, gate_masked, score)."""
    v, g = geglu_up(x, w1, b1, act)
    lead, h = g.shape[:-1], g.shape[-1]
    score = torch.matmul(g.reshape(-1, h), patterns.transpose(0, 1))
    idx = list(expert_idx)
    score[:, idx] = score[:, idx] + 5.0 * torch.tensor(std).to(g.dtype)[idx]
    kk = int(0.8 * k)
    labels = torch.topk(score, k=kk, dim=-1)[1].view(*lead, kk)
    mask = torch.nn.functional.embedding(labels, patterns).sum(-2)
    g = g.clone()
    g[mask == 0] = 0
    return v * g, labels, g, score


def wanda_column_sumsq(h_out: torch.Tensor) -> torch.Tensor:
    """Wanda.hook_fn (wanda_receiver.py:47-53) + ColumnNormCalculator.add_rows (utils.py:330-337): rows of the
    GEGLU output are L2-normalised, the column norms accumulate as sqrt(old^2 + new^2); this returns the squared
    column norms of ONE call (the quantity that adds up)."""
    rows = torch.nn.functional.normalize(h_out.reshape(-1, h_out.shape[-1]), p=2, dim=1)
    return torch.norm(rows, dim=0) ** 2


def wanda_score_mask(w2_abs: torch.Tensor, norm_base: torch.Tensor, norm_adj: torch.Tensor, ratio: float) -> np.ndarray:
    """modularity/wanda.py:143-165: metric = |W2| * column norm (base / adj prompts); a weight is a skilled-neuron
    weight iff its adj metric exceeds its base metric AND it is among the int(ratio * h) largest adj metrics of its
    output row.  Returns the dense {0,1} int mask [d, h]."""
    metric_base = w2_abs * norm_base
    metric_adj = w2_abs * norm_adj
    k = int(ratio * metric_adj.shape[1])
    _, order = torch.sort(metric_adj, dim=1, descending=True)
    top = torch.zeros_like(w2_abs)
    top.scatter_(1, order[:, :k], 1)
    return ((metric_adj > metric_base) * top).numpy().astype(int)


def union_over_time(masks: Sequence[np.ndarray], select_ratio: float) -> np.ndarray:
    """benchmarks/save_union_over_time.py:192-209: a weight stays masked iff it is masked at more than
    select_ratio * T of the T timesteps."""
    total = np.zeros_like(np.asarray(masks[0]), dtype=np.int64)
    for m in masks:
        total += np.asarray(m)
    return (total > (select_ratio * len(masks))).astype(int)


class TimeLayerClock:
    """(timestep, layer) state machine advanced once per hook call
    (predictivity.py:25-30; frequency_measure.py:24-29 hard-codes n_layers-1 == 15)."""

    def __init__(self, n_layers: int):
        self.n_layers = n_layers
        self.timestep = 0
        self.layer = 0

    def tick(self) -> None:
        if self.layer == self.n_layers - 1:
            self.layer = 0
            self.timestep += 1
        else:
            self.layer += 1

    def reset(self) -> None:
        self.timestep = 0
        self.layer = 0


class Welford:
    """Running mean (utils.py:233-251 `Average`) and Welford variance (utils.py:253-273
    `StandardDev`) of a stream of vectors, as StatMeter keeps per (t, layer) (utils.py:276-303)."""

    def __init__(self):
        self.n = 0
        self.sum = 0
        self.avg = 0
        self.mean = 0
        self.m2 = 0

    def update(self, x: np.ndarray) -> None:
        self.sum = self.sum + x
        self.n += 1
        self.avg = self.sum / self.n
        delta = x - self.mean
        self.mean = self.mean + delta / self.n
        self.m2 = self.m2 + delta * (x - self.mean)

    def stddev(self):
        if self.n < 2:
            return float("nan")
        return (self.m2 / (self.n - 1)) ** 0.5


def average_counters(per_image: List[Dict[int, Dict[int, np.ndarray]]], layer_names: List[str],
                     timesteps: int) -> Dict[int, Dict[str, List[float]]]:
    """freq_expert_select.main accumulation (freq_expert_select.py:43-64): per-image counters
    divided by the number of images, keyed by sorted FFN weight name."""
    n_img = len(per_image)
    out = {t: {nm: [0] * len(per_image[0][t][i]) for i, nm in enumerate(layer_names)}
           for t in range(timesteps)}
    for counter in per_image:
        for t in range(timesteps):
            for i, nm in enumerate(layer_names):
                for e in range(len(out[t][nm])):
                    out[t][nm][e] += counter[t][i][e] / n_img
    return out


# --------------------------------------------------------------------------------------
# Helpers shared by tests / bench (synthetic inputs of SURVEY.md section 8d)
# --------------------------------------------------------------------------------------
def synthetic_layer(d: int, h: int, tokens_shape: Tuple[int, int], expert_size: int,
                    seed: int = 0, layernorm_input: bool = True):
    """nn.Linear default init (seed), x = randn (seed+1) optionally LayerNorm-ed (the FFN input
    is norm3(x) upstream), balanced random labels (seed)."""
    g = torch.Generator().manual_seed(seed)
    bound1 = 1.0 / math.sqrt(d)
    w1 = (torch.rand(2 * h, d, generator=g) * 2 - 1) * bound1
    b1 = (torch.rand(2 * h, generator=g) * 2 - 1) * bound1
    bound2 = 1.0 / math.sqrt(h)
    w2 = (torch.rand(d, h, generator=g) * 2 - 1) * bound2
    b2 = (torch.rand(d, generator=g) * 2 - 1) * bound2
    gx = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(*tokens_shape, d, generator=gx)
    if layernorm_input:
        x = F.layer_norm(x, (d,))
    labels = balanced_labels(h, expert_size, seed)
    return dict(x=x, w1=w1, b1=b1, w2=w2, b2=b2, labels=labels)


def topk_margin(score: torch.Tensor, k: int) -> torch.Tensor:
    """k-th minus (k+1)-th largest score per token: the router margin of BASELINE.md section 5."""
    if k >= score.shape[-1]:
        return torch.full((score.shape[0],), float("inf"))
    top = torch.topk(score, k=k + 1, dim=-1)[0]
    return top[:, k - 1] - top[:, k]


def labels_to_bitmask(labels: torch.Tensor, n_experts: int) -> np.ndarray:
    """[T, k] indices -> [T, ceil(E/32)] uint32 words, bit (e % 32) of word (e // 32)."""
    lab = labels.reshape(-1, labels.shape[-1]).cpu().numpy()
    words = (n_experts + 31) // 32
    out = np.zeros((lab.shape[0], words), dtype=np.uint32)
    for j in range(lab.shape[1]):
        e = lab[:, j]
        np.bitwise_or.at(out, (np.arange(lab.shape[0]), e // 32), (1 << (e % 32)).astype(np.uint32))
    return out
