"""CPU: the oracle restatement against the committed golden vectors, which were produced by the
reference's own receivers (oracle/gen_golden.py).  Needs neither /root/reference nor a GPU."""
import os

import numpy as np
import pytest
import torch

import moe_ffn_oracle as O


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def T(a):
    return torch.from_numpy(np.asarray(a))


FULL_CASES = ["moefy_small_gelu", "moefy_small_relu", "moefy_es20", "moefy_k_equals_E", "moefy_ragged"]


@pytest.mark.parametrize("name", FULL_CASES)
def test_moefy_full_cases_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    pat = O.patterns_from_labels(g["labels"])
    k = O.topk_from_ratio(pat.shape[0], float(g["ratio"]))
    assert k == int(g["k"]) and pat.shape[0] == int(g["E"])
    H, labels, gate, score = O.moefy_forward(T(g["x"]), T(g["w1"]), T(g["b1"]), pat, k, int(g["act"]))
    assert torch.equal(H, T(g["H"]))
    assert torch.equal(gate, T(g["gate"]))
    assert torch.equal(score, T(g["score"]))
    assert np.array_equal(O.labels_to_bitmask(labels, pat.shape[0]), g["bitmask"])
    assert np.array_equal(O.selection_counts(labels, pat.shape[0]), g["counts_row0"])
    assert np.array_equal(O.selection_counts(labels, pat.shape[0], slice(None)), g["counts_all"])
    assert torch.equal(O.down_proj(H, T(g["w2"]), T(g["b2"])), T(g["y"]))
    # the reference's own sign assertion (moefy.py:50-51): gates >= 0 iff ReLU
    assert bool(torch.all(gate >= 0)) == (int(g["act"]) == O.ACT_RELU)


@pytest.mark.parametrize("name", ["config1_E64", "config1_E20"])
def test_config1_digests(golden_dir, name):
    """BASELINE config 1 (d=320, h=1280, 4096 tokens) regenerated from seeds and checked against
    the digests of the reference run."""
    g = load(golden_dir, name)
    layer = O.synthetic_layer(int(g["d"]), int(g["h"]), tuple(int(v) for v in g["shape"]), int(g["es"]), int(g["seed"]))
    assert abs(float(layer["x"].double().sum()) - float(g["x_sum"])) < 1e-9      # RNG stream unchanged
    pat = O.patterns_from_labels(layer["labels"])
    H, labels, _, score = O.moefy_forward(layer["x"], layer["w1"], layer["b1"], pat, int(g["k"]), int(g["act"]))
    assert np.array_equal(O.labels_to_bitmask(labels, pat.shape[0]), g["bitmask"])
    assert np.array_equal(O.selection_counts(labels, pat.shape[0]), g["counts_row0"])
    assert np.array_equal(H.reshape(-1)[::97].numpy(), g["H_sample"])
    assert np.array_equal(O.expert_predictivity(score), g["score_colmax"])
    assert int(g["counts_row0"].sum()) == 4096 * int(g["k"])


def test_frequency_counter(golden_dir):
    g = load(golden_dir, "frequency_small")
    n_layers, S, E = int(g["n_layers"]), int(g["S"]), int(g["E"])
    pat = O.patterns_from_labels(g["labels"])
    clock = O.TimeLayerClock(n_layers)
    counter = np.zeros((2, n_layers, E))
    for x in g["xs"]:
        _, labels, _, _ = O.moefy_forward(T(x), T(g["w1"]), T(g["b1"]), pat, int(g["k"]))
        O.frequency_update(counter[clock.timestep, clock.layer], labels, S)
        clock.tick()
    assert (clock.timestep, clock.layer) == (1, 2)       # 18 calls wrap the 16-layer sweep once
    assert np.array_equal(counter, g["label_counter"])
    assert np.array_equal(np.rint(counter * S).astype(np.int64), g["int_counts"])
    assert int(g["int_counts"][0, 0].sum()) == S * int(g["k"])   # row 0 only: S tokens x k slots


def test_expert_predictivity_and_welford(golden_dir):
    g = load(golden_dir, "expert_predictivity_small")
    pat = O.patterns_from_labels(g["labels"])
    w = O.Welford()
    for x, want in zip(g["xs"], g["max_gate"]):
        _, gate = O.geglu_up(T(x), T(g["w1"]), T(g["b1"]))
        got = O.expert_predictivity(O.expert_scores(gate, pat))
        assert np.array_equal(got, want)
        w.update(got)
    assert np.array_equal(w.avg, g["avg"]) and np.array_equal(w.stddev(), g["std"])


@pytest.mark.parametrize("name", ["remove_experts_small", "remove_experts_crowded"])
def test_remove_experts(golden_dir, name):
    g = load(golden_dir, name)
    pat = O.patterns_from_labels(g["labels"])
    removed = [int(v) for v in g["removed"]]
    for (t, l) in [(0, 0), (0, 1), (19, 0), (20, 0)]:
        lst = removed if l == 0 else []
        H, labels, _, score = O.remove_experts_forward(T(g["x"]), T(g["w1"]), T(g["b1"]), pat, int(g["k"]), lst, t)
        assert torch.equal(H, T(g[f"H_t{t}_l{l}"]))
        assert np.array_equal(O.labels_to_bitmask(labels, pat.shape[0]), g[f"bitmask_t{t}_l{l}"])
        if lst and t < 20:   # removed experts score exactly 0
            assert torch.all(score[:, lst] == 0)
    # t = 20 behaves like no removal at all
    assert np.array_equal(g["H_t20_l0"], g["H_t0_l1"])


def test_remove_neurons(golden_dir):
    g = load(golden_dir, "remove_neurons_small")
    flags = g["flags"].tolist()
    H, gate = O.remove_neurons_forward(T(g["x"]), T(g["w1"]), T(g["b1"]), flags)
    assert torch.equal(H, T(g["H_removed"]))
    idx = np.nonzero(g["flags"])[0]
    assert torch.all(gate[:, :, idx] == O.REMOVED_NEURON_GATE)
    H2, _ = O.remove_neurons_forward(T(g["x"]), T(g["w1"]), T(g["b1"]), [])
    assert torch.equal(H2, T(g["H_plain"]))


def test_wanda_and_union(golden_dir):
    g = load(golden_dir, "wanda_small")
    hid, w2, b2 = T(g["hid"]), T(g["w2"]), T(g["b2"])
    assert torch.equal(O.wanda_down_proj(hid, w2, b2, g["mask_a"]), T(g["y_a"]))
    union = O.mask_union(g["mask_a"], g["mask_b"], g["mask_c"])
    assert np.array_equal(union, g["union"])
    assert torch.equal(O.wanda_down_proj(hid, w2, b2, union), T(g["y_union"]))
    assert torch.equal(O.down_proj(hid, w2, b2), T(g["y_stock"]))


def test_wanda_csv_fixture_digest(golden_dir):
    """Digest of the reference's real mask fixture weights_320_1280.csv."""
    g = load(golden_dir, "wanda_csv_320_1280")
    assert g["packed"].shape == (5, 320, 160)
    masks = np.unpackbits(g["packed"], axis=2, bitorder="little")
    assert np.allclose(masks.reshape(5, -1).mean(1), g["density"])
    assert 0.02 < g["density"].min() and g["density"].max() < 0.03     # SURVEY: density 2.2-2.8 %
    union = O.mask_union(*[masks[i] for i in range(5)])
    assert np.array_equal(np.packbits(union.astype(np.uint8), axis=1, bitorder="little"), g["union_packed"])


def test_topk_ratio_table():
    """k = int(E * ratio) with Python float semantics."""
    assert [O.topk_from_ratio(64, r / 10) for r in range(1, 11)] == [6, 12, 19, 25, 32, 38, 44, 51, 57, 64]
    assert [O.topk_from_ratio(256, r / 10) for r in range(1, 11)] == [25, 51, 76, 102, 128, 153, 179, 204, 230, 256]
    assert [O.topk_from_ratio(20, r / 10) for r in range(1, 11)] == [2, 4, 6, 8, 10, 12, 14, 16, 18, 20]


# ------------------------------------------------------------------ remaining receivers
def test_get_experts(golden_dir):
    g = load(golden_dir, "get_experts_small")
    pat = O.patterns_from_labels(g["labels"])
    k = O.topk_from_ratio(pat.shape[0], float(g["ratio"]))
    for tag, box in (("all", None), ("bb", g["bb"].tolist())):
        H, labels, mean = O.get_experts_labels(T(g["x"]), T(g["w1"]), T(g["b1"]), pat, k, box)
        assert labels == g[f"labels_{tag}"].tolist()
        assert np.array_equal(mean.numpy(), g[f"mean_{tag}"])
    assert torch.equal(H, T(g["H"]))


def test_add_experts(golden_dir):
    g = load(golden_dir, "add_experts_small")
    pat = O.patterns_from_labels(g["labels"])
    k = O.topk_from_ratio(pat.shape[0], float(g["ratio"]))
    H, labels, gate, score = O.add_experts_forward(T(g["x"]), T(g["w1"]), T(g["b1"]), pat, k, g["experts"].tolist(),
                                                   g["std"].tolist())
    assert torch.equal(H, T(g["H"])) and np.array_equal(score.numpy(), g["score"])
    assert np.array_equal(O.labels_to_bitmask(labels.reshape(-1, labels.shape[-1]), pat.shape[0]), g["bitmask"])
    assert labels.shape[-1] == int(0.8 * k)


def test_wanda_receiver_column_norms(golden_dir):
    g = load(golden_dir, "wanda_receiver_small")
    ssq = torch.zeros(int(g["h"]))
    for x in g["xs"]:
        v, gate = O.geglu_up(T(x), T(g["w1"]), T(g["b1"]), O.ACT_RELU)
        ssq += O.wanda_column_sumsq(v * gate)
    assert np.allclose(torch.sqrt(ssq).numpy(), g["column_norms"], rtol=1e-5, atol=1e-7)
    v, gate = O.geglu_up(T(g["xs"][0]), T(g["w1"]), T(g["b1"]), O.ACT_RELU)
    assert torch.equal(gate, T(g["gate0"])) and torch.equal(v * gate, T(g["H0"]))     # SparsityMeasure


def test_wanda_scoring_and_union_over_time(golden_dir):
    """modularity/wanda.py:143-165 and save_union_over_time.py:189-211 (fixtures produced by executing those lines)."""
    g = load(golden_dir, "wanda_scoring_small")
    d, h, Tn = int(g["d"]), int(g["h"]), int(g["T"])
    want = np.unpackbits(g["masks"], axis=-1)[..., :h].astype(int)
    masks = []
    for t in range(Tn):
        m = O.wanda_score_mask(T(g["w2"]).abs(), T(g["norm_base"][t]), T(g["norm_adj"][t]), float(g["ratio"]))
        assert np.array_equal(m, want[t])
        assert m.sum(1).max() <= int(float(g["ratio"]) * h)
        masks.append(m)
    union = O.union_over_time(masks, float(g["select_ratio"]))
    assert np.array_equal(union, np.unpackbits(g["union"], axis=-1)[..., :h].astype(int))
