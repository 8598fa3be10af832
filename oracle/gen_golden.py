"""Generate tests/golden/*.npz by running the REFERENCE'S OWN receiver code, unmodified.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

How: `oracle/shim/` provides a stand-in for the uninstallable `diffusers` package (module
surface only) and for the head of the reference's `utils.py`; with
sys.path = [shim, /root/reference, /root/reference/moefication] the reference's
`neuron_receivers` and `helper` import verbatim.  Each case below builds a synthetic GEGLU
module, attaches experts with the reference's `helper.modify_ffn`, calls the reference
`hook_fn`, and (1) asserts that `oracle/moe_ffn_oracle.py` reproduces the reference output
bit-for-bit, (2) stores the reference output.  Large cases store digests (bitmasks, counts,
strided samples) instead of full tensors to keep the fixtures small.
"""
import io
import json
import os
import pickle
import sys
import tempfile
import types
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MOE_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "shim"), REF, os.path.join(REF, "moefication"), HERE]

with contextlib.redirect_stdout(io.StringIO()):
    import neuron_receivers as ref_nr          # the reference, unmodified
    import helper as ref_helper                # /root/reference/moefication/helper.py
from diffusers.models.activations import GEGLU, LoRACompatibleLinear  # the shim
import moe_ffn_oracle as O

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def make_module(layer, ratio, act):
    """Synthetic GEGLU carrying the layer's weights, MoEfied by the reference's helper.modify_ffn."""
    h2, d = layer["w1"].shape
    m = GEGLU(d, h2 // 2)
    with torch.no_grad():
        m.proj.weight.copy_(layer["w1"])
        m.proj.bias.copy_(layer["b1"])
    if act == O.ACT_RELU:
        m.gelu = torch.nn.functional.relu      # sparsity/relufy_model.py:35
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "labels")
        torch.save([int(v) for v in layer["labels"]], p)
        quiet(ref_helper.modify_ffn, m, p, ratio)
    return m


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def sample(t, stride=97):
    return t.reshape(-1)[::stride].numpy().copy()


@torch.no_grad()
def case_moefy(name, d, h, shape, es, ratio, act, seed, full):
    layer = O.synthetic_layer(d, h, shape, es, seed)
    mod = make_module(layer, ratio, act)
    rec = quiet(ref_nr.MOEFy, seed)
    H = rec.hook_fn(mod, (layer["x"],), None)
    gate = rec.gates[0]
    # restatement must be bit-identical to the reference
    pat = O.patterns_from_labels(layer["labels"])
    k = O.topk_from_ratio(pat.shape[0], ratio)
    assert torch.equal(pat, mod.patterns) and k == mod.k
    Ho, labels, gate_o, score = O.moefy_forward(layer["x"], layer["w1"], layer["b1"], pat, k, act)
    assert torch.equal(Ho, H), name
    assert torch.equal(gate_o, gate), name
    y = O.down_proj(H, layer["w2"], layer["b2"])
    E = pat.shape[0]
    common = dict(d=d, h=h, shape=np.array(shape), es=es, ratio=ratio, act=act, seed=seed, k=k, E=E,
                  bitmask=O.labels_to_bitmask(labels, E),
                  counts_row0=O.selection_counts(labels, E),
                  counts_all=O.selection_counts(labels, E, slice(None)),
                  score_colmax=O.expert_predictivity(score),
                  margin=O.topk_margin(score, k).numpy(),
                  x_sum=np.float64(layer["x"].double().sum()), w1_sum=np.float64(layer["w1"].double().sum()))
    if full:
        save(name, x=layer["x"].numpy(), w1=layer["w1"].numpy(), b1=layer["b1"].numpy(),
             w2=layer["w2"].numpy(), b2=layer["b2"].numpy(), labels=layer["labels"],
             H=H.numpy(), gate=gate.numpy(), score=score.numpy(), y=y.numpy(), **common)
    else:
        save(name, H_sample=sample(H), y_sample=sample(y), score_sample=sample(score),
             H_abs_sum=np.float64(H.double().abs().sum()), y_abs_sum=np.float64(y.double().abs().sum()),
             **common)


@torch.no_grad()
def case_frequency(name, d, h, B, S, es, ratio, seed, n_layers=16, n_calls=18):
    """FrequencyMeasure over n_calls hook calls (wraps layer 15 -> timestep+1)."""
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, ratio, O.ACT_GELU)
    E = mod.patterns.shape[0]
    names = [f"l{i:02d}" for i in range(n_layers)]
    rec = quiet(ref_nr.FrequencyMeasure, seed, 2, n_layers, {n: E for n in names}, names)
    clock = O.TimeLayerClock(n_layers)
    oc = {t: {l: np.zeros(E) for l in range(n_layers)} for t in range(2)}
    xs = []
    for c in range(n_calls):
        x = torch.nn.functional.layer_norm(
            torch.randn(B, S, d, generator=torch.Generator().manual_seed(1000 + c)), (d,))
        xs.append(x.numpy())
        H = rec.hook_fn(mod, (x,), None)
        Ho, labels, _, _ = O.moefy_forward(x, layer["w1"], layer["b1"], mod.patterns, mod.k)
        assert torch.equal(H, Ho)
        O.frequency_update(oc[clock.timestep][clock.layer], labels, S)
        clock.tick()
    assert (rec.timestep, rec.layer) == (clock.timestep, clock.layer) == (1, 2)
    ref_counter = np.stack([np.stack([rec.label_counter[t][l] for l in range(n_layers)]) for t in range(2)])
    ora_counter = np.stack([np.stack([oc[t][l] for l in range(n_layers)]) for t in range(2)])
    assert np.array_equal(ref_counter, ora_counter)
    save(name, d=d, h=h, B=B, S=S, es=es, ratio=ratio, seed=seed, n_layers=n_layers, n_calls=n_calls,
         k=mod.k, E=E, xs=np.stack(xs), w1=layer["w1"].numpy(), b1=layer["b1"].numpy(),
         labels=layer["labels"], label_counter=ref_counter,
         int_counts=np.rint(ref_counter * S).astype(np.int64))


@torch.no_grad()
def case_expert_predictivity(name, d, h, B, S, es, ratio, seed, n_prompts=3):
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, ratio, O.ACT_GELU)
    rec = quiet(ref_nr.ExpertPredictivity, seed, 1, 16)
    w = O.Welford()
    xs, outs = [], []
    for p in range(n_prompts):
        x = torch.nn.functional.layer_norm(
            torch.randn(B, S, d, generator=torch.Generator().manual_seed(2000 + p)), (d,))
        xs.append(x.numpy())
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (x,), None)
        v, g = O.geglu_up(x, layer["w1"], layer["b1"])
        assert torch.equal(H, v * g)                      # unmasked output (expert_activation.py:62)
        mx = O.expert_predictivity(O.expert_scores(g, mod.patterns))
        assert np.array_equal(mx, rec.max_gate[0][0])
        w.update(mx)
        outs.append(mx)
    avg = rec.predictivity.results["time_steps"][0][0]["avg"].avg
    std = rec.predictivity.results["time_steps"][0][0]["std"].stddev()
    assert np.array_equal(avg, w.avg) and np.array_equal(std, w.stddev())
    save(name, d=d, h=h, B=B, S=S, es=es, ratio=ratio, seed=seed, xs=np.stack(xs), w1=layer["w1"].numpy(),
         b1=layer["b1"].numpy(), labels=layer["labels"], max_gate=np.stack(outs), avg=avg, std=std)


@torch.no_grad()
def case_remove_experts(name, d, h, B, S, es, ratio, seed, removed, T=22, n_layers=2, with_down=False):
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, ratio, O.ACT_GELU)
    with tempfile.TemporaryDirectory() as td:
        for t in range(T):
            for l in range(n_layers):
                lst = removed if l == 0 else []           # layer 1: empty list branch
                json.dump(lst, open(os.path.join(td, f"timestep_{t}_layer_{l}.json"), "w"))
        rec = quiet(ref_nr.RemoveExperts, seed, td, T, n_layers)
    outs = {}
    for (t, l) in [(0, 0), (0, 1), (19, 0), (20, 0)]:      # <20 removes, >=20 does not
        rec.timestep, rec.layer = t, l
        H = rec.hook_fn(mod, (layer["x"],), None)
        lst = removed if l == 0 else []
        Ho, labels, gate, score = O.remove_experts_forward(layer["x"], layer["w1"], layer["b1"], mod.patterns,
                                                           mod.k, lst, t)
        assert torch.equal(H, Ho), (t, l)
        outs[f"H_t{t}_l{l}"] = H.numpy()
        outs[f"bitmask_t{t}_l{l}"] = O.labels_to_bitmask(labels, mod.patterns.shape[0])
        outs[f"score_t{t}_l{l}"] = score.numpy()
        if with_down:    # what the stock ff.net.2 makes of the hook output (upstream FeedForward: Dropout(0) -> Linear)
            outs[f"y_t{t}_l{l}"] = O.down_proj(H, layer["w2"], layer["b2"]).numpy()
    if with_down:
        outs["w2"], outs["b2"] = layer["w2"].numpy(), layer["b2"].numpy()
    rec.timestep, rec.layer = 0, n_layers - 1
    rec.update_time_layer()
    assert (rec.timestep, rec.layer) == (1, 0)
    save(name, d=d, h=h, B=B, S=S, es=es, ratio=ratio, seed=seed, removed=np.array(removed), k=mod.k,
         x=layer["x"].numpy(), w1=layer["w1"].numpy(), b1=layer["b1"].numpy(), labels=layer["labels"], **outs)


@torch.no_grad()
def case_remove_neurons(name, d, h, B, S, es, seed, frac=0.05):
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, 1.0, O.ACT_GELU)
    flags = (np.random.RandomState(seed + 7).rand(h) < frac).astype(float).tolist()
    with tempfile.TemporaryDirectory() as td:
        json.dump(flags, open(os.path.join(td, "predictivity_0_0.json"), "w"))
        json.dump([], open(os.path.join(td, "predictivity_0_1.json"), "w"))
        rec = quiet(ref_nr.RemoveNeurons, seed, td, 1, 2)
    H0 = rec.hook_fn(mod, (layer["x"],), None)
    H1 = rec.hook_fn(mod, (layer["x"],), None)            # empty list -> plain GEGLU
    Ho0, _ = O.remove_neurons_forward(layer["x"], layer["w1"], layer["b1"], flags)
    Ho1, _ = O.remove_neurons_forward(layer["x"], layer["w1"], layer["b1"], [])
    assert torch.equal(H0, Ho0) and torch.equal(H1, Ho1)
    save(name, d=d, h=h, B=B, S=S, seed=seed, flags=np.array(flags), x=layer["x"].numpy(),
         w1=layer["w1"].numpy(), b1=layer["b1"].numpy(), H_removed=H0.numpy(), H_plain=H1.numpy())


def wanda_like_mask(rs, d, h, ratio):
    """per-output-row top-(ratio*h) columns AND Bernoulli(1/2) -- mirrors the structure produced by
    modularity/wanda.py:151-165 (row-wise top-ratio selection intersected with a comparison)."""
    m = np.zeros((d, h), dtype=np.int64)
    kk = max(1, int(ratio * h))
    for r in range(d):
        cols = rs.choice(h, kk, replace=False)
        m[r, cols[rs.rand(kk) < 0.5]] = 1
    return m


@torch.no_grad()
def case_wanda(name, d, h, B, S, seed):
    import scipy.sparse as sp
    layer = O.synthetic_layer(d, h, (B, S), 16, seed)
    rs = np.random.RandomState(seed + 11)
    masks = {c: wanda_like_mask(rs, d, h, 0.05) for c in ("a", "b", "c")}
    hid = torch.randn(B, S, h, generator=torch.Generator().manual_seed(seed + 3))
    lin = LoRACompatibleLinear(h, d)
    lin.weight.copy_(layer["w2"]); lin.bias.copy_(layer["b2"])
    stock = lin(hid)
    removers = {}
    for c, m in masks.items():
        with tempfile.TemporaryDirectory() as td:
            with open(os.path.join(td, "timestep_0_layer_0.pkl"), "wb") as f:
                pickle.dump(sp.csr_matrix(m), f)                       # modularity/wanda.py:168-173
            removers[c] = quiet(ref_nr.WandaRemoveNeuronsFast, seed, td, 1, 1)
    y = {c: removers[c].linear_hook_fn(lin, (hid,), stock) for c in masks}
    for c in masks:
        assert torch.equal(y[c], O.wanda_down_proj(hid, layer["w2"], layer["b2"], masks[c]))
        removers[c].reset_time_layer()
    # union through the reference's handle_multiple_concepts (its ctor is bit-rotted, SURVEY A.3 item 6,
    # so the method is driven with a stand-in `self` holding real reference removers)
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "timestep_0_layer_0.pkl"), "wb") as f:
            pickle.dump(sp.csr_matrix(masks["a"]), f)
        union = quiet(ref_nr.WandaRemoveNeuronsFast, seed, td, 1, 1)
    fake = types.SimpleNamespace(removers=removers, union_neuron_remover=union)
    # numpy>=2 / torch>=2 return a Tensor from np.logical_or(ndarray, Tensor), which breaks the
    # reference's `.astype(int)` (multi_concept_remover.py:53); hand it ndarrays as its era's numpy did.
    for c in removers:
        removers[c].expert_indices[0][0] = removers[c].expert_indices[0][0].numpy()
    quiet(ref_nr.MultiConceptRemoverWanda.reset_union_remover, fake)
    quiet(ref_nr.MultiConceptRemoverWanda.handle_multiple_concepts, fake, ["a", "b", "c"])
    u = np.asarray(union.expert_indices[0][0])
    assert np.array_equal(u, O.mask_union(*masks.values()))
    union.expert_indices[0][0] = torch.tensor(u)
    union.reset_time_layer()
    yu = union.linear_hook_fn(lin, (hid,), stock)
    assert torch.equal(yu, O.wanda_down_proj(hid, layer["w2"], layer["b2"], u))
    save(name, d=d, h=h, B=B, S=S, seed=seed, hid=hid.numpy(), w2=layer["w2"].numpy(), b2=layer["b2"].numpy(),
         mask_a=masks["a"].astype(np.uint8), mask_b=masks["b"].astype(np.uint8),
         mask_c=masks["c"].astype(np.uint8), union=u.astype(np.uint8),
         y_a=y["a"].numpy(), y_union=yu.numpy(), y_stock=stock.numpy())


def case_wanda_csv_fixture(name):
    """Digest of the reference's real Wanda-mask fixture weights_320_1280.csv (five flattened
    320x1280 binary masks) + their union: pins the bit-packing / union kernels on real data."""
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, "weights_320_1280.csv"))
    cols = list(df.columns)
    masks = [df[c].to_numpy().astype(np.uint8).reshape(320, 1280) for c in cols]
    union = O.mask_union(*masks).astype(np.uint8)
    save(name, columns=np.array(cols), packed=np.stack([np.packbits(m, axis=1, bitorder="little") for m in masks]),
         union_packed=np.packbits(union, axis=1, bitorder="little"),
         density=np.array([m.mean() for m in masks]), union_density=np.float64(union.mean()))


@torch.no_grad()
def case_get_experts(name, d, h, B, S, es, ratio, seed):
    """GetExperts (get_experts.py): top-k of the token-averaged expert score, with and without a bounding box."""
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, ratio, O.ACT_GELU)
    E = mod.patterns.shape[0]
    rec = quiet(ref_nr.GetExperts, seed, 1, 16, {"l": E}, ["l"] * 16)
    out = {}
    bb = sorted(np.random.RandomState(seed).choice(S, S // 3, replace=False).tolist())
    for tag, box in (("all", None), ("bb", bb)):
        mod.bounding_box = box
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (layer["x"],), None)
        Ho, labels, mean = O.get_experts_labels(layer["x"], layer["w1"], layer["b1"], mod.patterns, mod.k, box)
        assert torch.equal(H, Ho) and labels == rec.label_counter[0][0]
        out[f"labels_{tag}"] = np.array(labels)
        out[f"mean_{tag}"] = mean.numpy()
    save(name, d=d, h=h, B=B, S=S, es=es, ratio=ratio, seed=seed, x=layer["x"].numpy(), w1=layer["w1"].numpy(),
         b1=layer["b1"].numpy(), labels=layer["labels"], bb=np.array(bb), H=H.numpy(), **out)


@torch.no_grad()
def case_add_experts(name, d, h, B, S, es, ratio, seed):
    """AddExperts (add_skilled_experts.py): boosted scores for the listed experts, top-int(0.8 k)."""
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, ratio, O.ACT_GELU)
    E = mod.patterns.shape[0]
    rs = np.random.RandomState(seed)
    experts = sorted(rs.choice(E, 2, replace=False).tolist())
    std = rs.uniform(0.5, 2.0, E).tolist()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "adj", "skilled", "experts")
        os.makedirs(path)
        json.dump({"time_steps": {"0": {"0": {"std": std}}}}, open(os.path.join(td, "adj", "predictivity_base_expert.json"), "w"))
        json.dump(experts, open(os.path.join(path, "timestep_0_layer_0.json"), "w"))
        rec = quiet(ref_nr.AddExperts, seed, path, 1, 1)
    H = rec.hook_fn(mod, (layer["x"],), None)
    Ho, labels, gate, score = O.add_experts_forward(layer["x"], layer["w1"], layer["b1"], mod.patterns, mod.k, experts, std)
    assert torch.equal(H, Ho) and torch.equal(rec.gates[0], gate)
    kk = labels.shape[-1]
    save(name, d=d, h=h, B=B, S=S, es=es, ratio=ratio, seed=seed, x=layer["x"].numpy(), w1=layer["w1"].numpy(),
         b1=layer["b1"].numpy(), labels=layer["labels"], experts=np.array(experts), std=np.array(std), H=H.numpy(),
         bitmask=O.labels_to_bitmask(labels.reshape(-1, kk), E), score=score.numpy(), margin=O.topk_margin(score, kk).numpy())


@torch.no_grad()
def case_wanda_receiver(name, d, h, B, S, es, seed, n_prompts=2):
    """Wanda (wanda_receiver.py): incremental column norms of the row-normalised GEGLU output; SparsityMeasure
    (sparsity_measure.py): captured activated gate, unmasked output."""
    layer = O.synthetic_layer(d, h, (B, S), es, seed)
    mod = make_module(layer, 1.0, O.ACT_RELU)
    rec = quiet(ref_nr.Wanda, seed, 1, 1)
    ssq = torch.zeros(h)
    xs = []
    for p in range(n_prompts):
        x = torch.nn.functional.layer_norm(torch.randn(B, S, d, generator=torch.Generator().manual_seed(3000 + p)), (d,))
        xs.append(x.numpy())
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (x,), None)
        v, g = O.geglu_up(x, layer["w1"], layer["b1"], O.ACT_RELU)
        assert torch.equal(H, v * g)
        ssq += O.wanda_column_sumsq(H)
    norms = rec.predictivity.get_column_norms()[0][0]
    assert torch.allclose(norms, torch.sqrt(ssq), rtol=1e-5, atol=1e-7)
    sp = quiet(ref_nr.SparsityMeasure, seed)
    Hs = sp.hook_fn(mod, (torch.from_numpy(xs[0]),), None)
    v, g = O.geglu_up(torch.from_numpy(xs[0]), layer["w1"], layer["b1"], O.ACT_RELU)
    assert torch.equal(Hs, v * g) and torch.equal(sp.gates[0], g)
    save(name, d=d, h=h, B=B, S=S, es=es, seed=seed, xs=np.stack(xs), w1=layer["w1"].numpy(), b1=layer["b1"].numpy(),
         labels=layer["labels"], column_norms=norms.numpy(), gate0=g.numpy(), H0=Hs.numpy())


def reference_lines(relpath, first, last):
    """Source lines [first, last] (1-based) of a reference script, dedented: the scripts below are `main()`s with
    file I/O around the arithmetic, so the arithmetic itself is executed from their own text."""
    import textwrap
    with open(os.path.join(REF, relpath)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


@torch.no_grad()
def case_wanda_scoring(name, d, h, ratio, seed, T=6, select_ratio=0.5):
    """Wanda scoring (modularity/wanda.py:143-165) and union over timesteps (save_union_over_time.py:189-211), both
    run from the reference's own source text on synthetic weights / norms."""
    import scipy.sparse
    rs = np.random.RandomState(seed)
    w2 = torch.from_numpy(rs.standard_normal((d, h)).astype(np.float32)).to(torch.bfloat16).float()
    masks, nb, na = [], [], []
    for t in range(T):
        norm_base = torch.from_numpy(np.abs(rs.standard_normal(h)).astype(np.float32))
        norm_adj = norm_base * torch.from_numpy(rs.uniform(0.5, 1.5, h).astype(np.float32))
        norm_adj[rs.choice(h, h // 16, replace=False)] = 0.0          # dead neurons (ReLU-fied model)
        ns = dict(torch=torch, np=np, scipy=scipy, gate_weights={"l": w2.abs()}, layer_names=["l"], l=0, t=0,
                  act_norms_base={0: {0: norm_base}}, act_norms_adj={0: {0: norm_adj}}, sparsity_ratio=ratio, print=lambda *a: None)
        exec(reference_lines("modularity/wanda.py", 143, 165), ns)
        ref_mask = ns["binary_mask"].numpy().astype(int)
        assert np.array_equal(O.wanda_score_mask(w2.abs(), norm_base, norm_adj, ratio), ref_mask)
        masks.append(ref_mask); nb.append(norm_base.numpy()); na.append(norm_adj.numpy())
    with tempfile.TemporaryDirectory() as td:
        for t, m in enumerate(masks):
            with open(os.path.join(td, f"timestep_{t}_layer_0.pkl"), "wb") as f:
                pickle.dump(scipy.sparse.csr_matrix(m), f)
        ns = dict(np=np, scipy=scipy, pickle=pickle, os=os, path=td, n_layers=1, timesteps=T, weights_shape=[(d, h)],
                  select_ratio=select_ratio, layer_names=["l"], print=lambda *a: None, masks={})
        exec(reference_lines("benchmarks/save_union_over_time.py", 189, 211), ns)
    union = np.asarray(ns["masks"]["l"]).astype(int)
    assert np.array_equal(O.union_over_time(masks, select_ratio), union)
    save(name, d=d, h=h, ratio=ratio, T=T, select_ratio=select_ratio, w2=w2.numpy(), norm_base=np.stack(nb), norm_adj=np.stack(na),
         masks=np.packbits(np.stack(masks).astype(np.uint8), axis=-1), union=np.packbits(union.astype(np.uint8), axis=-1))


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    # small, fully stored cases (inputs + full outputs)
    case_moefy("moefy_small_gelu", 32, 128, (2, 48), 16, 0.3, O.ACT_GELU, 0, full=True)
    case_moefy("moefy_small_relu", 32, 128, (2, 48), 16, 0.3, O.ACT_RELU, 1, full=True)
    case_moefy("moefy_es20", 64, 320, (2, 40), 20, 0.3, O.ACT_GELU, 2, full=True)       # 16 experts x 20
    case_moefy("moefy_k_equals_E", 32, 128, (1, 33), 16, 1.0, O.ACT_GELU, 3, full=True)  # identity masking
    case_moefy("moefy_ragged", 32, 128, (3, 1), 16, 0.5, O.ACT_GELU, 4, full=True)       # one token per row
    # config-1 sized (digest only): reference-faithful and BASELINE-literal expert geometry
    case_moefy("config1_E64", 320, 1280, (1, 4096), 20, 0.3, O.ACT_GELU, 0, full=False)
    case_moefy("config1_E20", 320, 1280, (1, 4096), 64, 0.3, O.ACT_GELU, 0, full=False)
    case_moefy("config1_E64_relu_b2", 320, 1280, (2, 4096), 20, 0.3, O.ACT_RELU, 5, full=False)
    case_frequency("frequency_small", 32, 128, 2, 48, 16, 0.3, 0)
    case_expert_predictivity("expert_predictivity_small", 32, 128, 2, 48, 16, 0.3, 0)
    case_remove_experts("remove_experts_small", 32, 128, 2, 48, 16, 0.6, 0, removed=[1, 5])
    case_remove_experts("remove_experts_crowded", 32, 128, 2, 48, 16, 1.0, 1, removed=[0, 2, 3, 7])
    case_remove_neurons("remove_neurons_small", 32, 128, 2, 48, 16, 0)
    case_wanda("wanda_small", 32, 128, 2, 24, 0)
    case_wanda_csv_fixture("wanda_csv_320_1280")
    # the remaining receivers on the same primitives
    case_get_experts("get_experts_small", 32, 128, 2, 48, 16, 0.3, 0)
    case_add_experts("add_experts_small", 32, 128, 2, 48, 16, 0.6, 1)
    case_wanda_receiver("wanda_receiver_small", 32, 128, 2, 24, 16, 2)
    # Wanda scoring and the union over timesteps
    case_wanda_scoring("wanda_scoring_small", 64, 256, 0.05, 3)
    # round 2: geometries the fused layer kernel covers (d, h multiples of 64), for the hook-API tests of moe_ffn_fused
    case_moefy("moefy_d64_gelu", 64, 256, (2, 96), 16, 0.3, O.ACT_GELU, 6, full=True)
    case_moefy("moefy_d64_relu", 64, 256, (2, 96), 16, 0.3, O.ACT_RELU, 7, full=True)
    case_moefy("moefy_d128_es64", 128, 512, (2, 40), 64, 0.5, O.ACT_GELU, 8, full=True)     # BASELINE-literal expert size
    case_remove_experts("remove_experts_d64", 64, 256, 2, 96, 16, 0.5, 9, removed=[1, 5, 9], with_down=True)


if __name__ == "__main__":
    main()
