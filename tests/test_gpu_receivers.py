"""GPU: the reference-facing receiver / hook API (neuron_receivers.*, moefication.helper) driven the
way the reference drives it, checked against the golden vectors produced by the reference's own
receivers and against the oracle."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

import moe_ffn_oracle as O
import neuron_receivers as nr
from moefication import helper
from moe_b200.sd_modules import GEGLU, FeedForward, FFNStackUNet, SyntheticFFNPipeline, sd_ffn_shapes
from gpu_util import DEV, rel_err, r16, OUT_REL_TOL, label_sets

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():   # the reference hooks run inside the pipeline's no_grad
        yield


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def T(a):
    return torch.from_numpy(np.asarray(a))


def make_ff(g, ratio=None, with_down=True):
    """FeedForward carrying the fixture weights (bf16, on the GPU), MoEfied through helper.modify_ffn."""
    h2, d = g["w1"].shape
    ff = FeedForward(d, inner_dim=h2 // 2)
    with torch.no_grad():
        ff.net[0].proj.weight.copy_(T(g["w1"]))
        ff.net[0].proj.bias.copy_(T(g["b1"]))
        if with_down and "w2" in g:
            ff.net[2].weight.copy_(T(g["w2"]))
            ff.net[2].bias.copy_(T(g["b2"]))
    ff = ff.to(DEV, torch.bfloat16)
    if ratio is not None:
        helper.modify_ffn(ff.net[0], [int(v) for v in g["labels"]], ratio, down=ff.net[2])
    return ff


def unpack_cols(t, module):
    """hook outputs are in the module's (packed) neuron order; golden tensors in the original order."""
    return t.float().cpu()[..., module._moe_state.layout.inv_perm]


@pytest.mark.parametrize("name", ["moefy_small_gelu", "moefy_small_relu", "moefy_es20", "moefy_k_equals_E"])
def test_moefy_hook(lib, golden_dir, name):
    g = load(golden_dir, name)
    ff = make_ff(g, float(g["ratio"]))
    mod = ff.net[0]
    if int(g["act"]) == O.ACT_RELU:
        mod.gelu = torch.nn.functional.relu                  # reference sparsity/relufy_model.py:35
    assert mod.k == int(g["k"]) and mod.patterns.shape[0] == int(g["E"])
    rec = nr.MOEFy(seed=0)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    H = rec.hook_fn(mod, (x,), None)
    assert H.shape == tuple(g["H"].shape) and H.dtype == x.dtype
    wide = np.repeat((g["margin"] > 0.25)[:, None], 1, 1).reshape(-1)
    n = wide.shape[0]
    Hc, Hr = unpack_cols(H, mod).reshape(n, -1), T(g["H"]).reshape(n, -1)
    assert rel_err(Hc[wide], Hr[wide]) < OUT_REL_TOL
    # the hook output feeds the stock (column-permuted) ff.net.2: full FFN within 1e-2 of the reference
    y = ff.net[2](H).float().cpu().reshape(n, -1)
    assert rel_err(y[wide], T(g["y"]).reshape(n, -1)[wide]) < 1.5e-2      # + cuBLAS bf16 down-projection
    # captured gate (reference moefy.py:25) is the masked activation, on the host
    assert len(rec.gates) == 1 and rec.gates[0].device.type == "cpu" and rec.gates[0].shape == H.shape
    gc = rec.gates[0].float()[..., mod._moe_state.layout.inv_perm].reshape(n, -1)
    assert rel_err(gc[wide], T(g["gate"]).reshape(n, -1)[wide]) < OUT_REL_TOL
    assert bool(torch.all(rec.gates[0] >= 0)) == (int(g["act"]) == O.ACT_RELU)


def test_frequency_measure_hook(lib, golden_dir):
    g = load(golden_dir, "frequency_small")
    n_layers, S, E, k = int(g["n_layers"]), int(g["S"]), int(g["E"]), int(g["k"])
    ff = make_ff(g, float(g["ratio"]), with_down=False)
    mod = ff.net[0]
    names = [f"l{i:02d}" for i in range(n_layers)]
    rec = nr.FrequencyMeasure(0, 2, n_layers, {n: E for n in names}, names)
    pat = O.patterns_from_labels(g["labels"])
    want = np.zeros((2, n_layers, E), dtype=np.int64)
    unsafe = np.zeros((2, n_layers), dtype=np.int64)
    clock = O.TimeLayerClock(n_layers)
    for x in g["xs"]:
        rec.hook_fn(mod, (T(x).to(DEV, torch.bfloat16),), None)
        _, labels, _, score = O.moefy_forward(r16(T(x)), r16(T(g["w1"])), T(g["b1"]), pat, k)
        want[clock.timestep, clock.layer] += O.selection_counts(labels, E)
        unsafe[clock.timestep, clock.layer] += int((O.topk_margin(score, k)[:S] <= 2e-3).sum())
        clock.tick()
    assert (rec.timestep, rec.layer) == (1, 2)
    got = rec.int_counts().cpu().numpy()
    assert got.sum(-1).tolist() == want.sum(-1).tolist()            # S*k selections per visited cell
    # bit-exact wherever the router margin allows; each unsafe token can move at most one count pair
    assert (np.abs(got - want).sum(-1) <= 2 * unsafe).all()
    assert np.array_equal(got[unsafe == 0], want[unsafe == 0])
    # vs the reference's own (fp32-input) counters: same totals, small L1 drift from bf16 inputs only
    assert np.abs(got - g["int_counts"]).sum() <= 0.02 * g["int_counts"].sum()
    lc = rec.label_counter
    assert np.allclose(lc[0][3], got[0, 3] / S) and abs(lc[0][0].sum() - k) < 1e-9
    rec.reset()
    assert int(rec.int_counts().sum()) == 0 and (rec.timestep, rec.layer) == (0, 0)


def test_expert_predictivity_hook(lib, golden_dir):
    g = load(golden_dir, "expert_predictivity_small")
    ff = make_ff(g, float(g["ratio"]), with_down=False)
    mod = ff.net[0]
    rec = nr.ExpertPredictivity(0, 1, 16)
    for x, want in zip(g["xs"], g["max_gate"]):
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (T(x).to(DEV, torch.bfloat16),), None)
        assert np.allclose(rec.max_gate[0][0], want, atol=0.05)
        v, gate = O.geglu_up(r16(T(x)), r16(T(g["w1"])), T(g["b1"]))
        assert rel_err(unpack_cols(H, mod), v * gate) < OUT_REL_TOL          # output is NOT masked
    cell = rec.predictivity.results["time_steps"][0][0]
    assert np.allclose(cell["avg"].avg, g["avg"], atol=0.05) and np.allclose(cell["std"].stddev(), g["std"], atol=0.05)


def test_remove_experts_hook(lib, golden_dir, tmp_path):
    g = load(golden_dir, "remove_experts_small")
    removed = [int(v) for v in g["removed"]]
    for t in range(22):
        for l in range(2):
            json.dump(removed if l == 0 else [], open(tmp_path / f"timestep_{t}_layer_{l}.json", "w"))
    ff = make_ff(g, float(g["ratio"]), with_down=False)
    mod = ff.net[0]
    rec = nr.RemoveExperts(0, str(tmp_path), 22, 2, capture_gates=False)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    pat = O.patterns_from_labels(g["labels"])
    for (t, l) in [(0, 0), (0, 1), (19, 0), (20, 0)]:
        rec.timestep, rec.layer = t, l
        H = rec.hook_fn(mod, (x,), None)
        lst = removed if l == 0 else []
        orc = O.remove_experts_forward(r16(T(g["x"])), r16(T(g["w1"])), T(g["b1"]), pat, mod.k, lst, t)
        safe = (O.topk_margin(orc[3], mod.k) > 2e-3).numpy()
        n = safe.shape[0]
        Hc = unpack_cols(H, mod).reshape(n, -1)
        assert rel_err(Hc[safe], orc[0].reshape(n, -1)[safe]) < OUT_REL_TOL
        # vs the reference's own fp32-input output: tokens whose margin in the REFERENCE run is wide enough
        # that bf16 input rounding cannot change the selected set
        wide = (O.topk_margin(T(g[f"score_t{t}_l{l}"]), mod.k) > 0.25).numpy() if mod.k < pat.shape[0] else safe
        assert wide.mean() > 0.3
        assert rel_err(Hc[wide], T(g[f"H_t{t}_l{l}"]).reshape(n, -1)[wide]) < OUT_REL_TOL
        if lst and t < 20:
            assert torch.all(Hc[:, (pat[lst].sum(0) > 0)] == 0)
    rec.timestep, rec.layer = 0, 1
    rec.update_time_layer()
    assert (rec.timestep, rec.layer) == (1, 0)


def test_remove_neurons_hook(lib, golden_dir, tmp_path):
    g = load(golden_dir, "remove_neurons_small")
    json.dump(g["flags"].tolist(), open(tmp_path / "predictivity_0_0.json", "w"))
    json.dump([], open(tmp_path / "predictivity_0_1.json", "w"))
    gg = dict(w1=g["w1"], b1=g["b1"])
    ff = make_ff(gg, None, with_down=False)          # NOT MoEfied: identity layout is attached lazily
    mod = ff.net[0]
    rec = nr.RemoveNeurons(0, str(tmp_path), 1, 2)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    H0 = rec.hook_fn(mod, (x,), None)
    H1 = rec.hook_fn(mod, (x,), None)
    assert rel_err(H0.float().cpu(), T(g["H_removed"])) < OUT_REL_TOL
    assert rel_err(H1.float().cpu(), T(g["H_plain"])) < OUT_REL_TOL
    idx = np.nonzero(g["flags"])[0]
    assert torch.all(rec.gates[0][..., idx].float() == torch.tensor(-0.17).bfloat16().float())
    assert (rec.timestep, rec.layer) == (1, 0)


def test_wanda_and_multi_concept_hooks(lib, golden_dir, tmp_path):
    import scipy.sparse as sp
    g = load(golden_dir, "wanda_small")
    root = str(tmp_path) + "/seed_%s_%s"
    for c in "abc":
        p = tmp_path / f"seed_0_{c}" / "skilled_neuron_wanda" / "0.05"
        os.makedirs(p)
        with open(p / "timestep_0_layer_0.pkl", "wb") as f:
            pickle.dump(sp.csr_matrix(g[f"mask_{c}"].astype(np.int64)), f)
    mc = nr.MultiConceptRemoverWanda(root, 0, 1, 1, concepts_to_remove=["a", "b", "c"],
                                     wanda_thr={"a": 0.05, "b": 0.05, "c": 0.05})
    lin = FeedForward(32).net[2]
    with torch.no_grad():
        lin.weight.copy_(T(g["w2"])); lin.bias.copy_(T(g["b2"]))
    lin = lin.to(DEV, torch.bfloat16)
    hid = T(g["hid"]).to(DEV, torch.bfloat16)
    ya = mc.removers["a"].linear_hook_fn(lin, (hid,), None)
    assert ya.shape == tuple(g["y_a"].shape) and rel_err(ya.float().cpu(), T(g["y_a"])) < OUT_REL_TOL
    mc.handle_multiple_concepts(["a", "b", "c"], device=DEV)
    un = mc.union_neuron_remover
    assert np.array_equal(un.mask_bits(0, 0, DEV).cpu().numpy().view(np.uint8),
                          np.packbits(g["union"].reshape(-1), bitorder="little"))
    un.reset_time_layer()
    yu = un.linear_hook_fn(lin, (hid,), None)
    assert rel_err(yu.float().cpu(), T(g["y_union"])) < OUT_REL_TOL
    assert (un.timestep, un.layer) == (1, 0)


def test_observe_activation_on_sd15_ffn_stack(lib):
    """The reference call sequence (freq_expert_select.py:29-64) on the SD-1.5 FFN stack with
    diffusers-identical module names: modify_ffn_to_experts -> FrequencyMeasure.observe_activation."""
    torch.manual_seed(0)
    unet = FFNStackUNet(latent_hw=16)                      # 16x16 latents: 256/64/16/4 tokens per layer
    pipe = SyntheticFFNPipeline(unet, num_inference_steps=3, device=DEV)
    labels = {n + ".proj.weight": O.balanced_labels(h, 20, seed=i) for i, (n, d, h, s) in enumerate(sd_ffn_shapes(16))}

    class Args:
        res_path = ""
        moefication = {"topk_experts": 0.3}
    states0 = [torch.randn(2, s, d, device=DEV, dtype=torch.bfloat16) for (_, d, _, s) in sd_ffn_shapes(16)]
    dense = [y.float() for y in unet(states0)]
    pipe, names, n_exp = helper.modify_ffn_to_experts(pipe, Args(), labels_by_name=labels)
    assert list(n_exp.values()) == [64, 64, 128, 128, 256, 256, 256, 256, 256, 256, 128, 128, 128, 64, 64, 64]
    after = [y.float() for y in unet(states0)]           # packing must not change the (unhooked) model
    for a, b in zip(dense, after):
        assert rel_err(b, a) < 2e-2
    rec = nr.FrequencyMeasure(0, 3, 16, n_exp, names)
    rec.reset()
    out, gates = rec.observe_activation(pipe, "a photo of a cat")
    assert gates == [] and len(out) == 16 and (rec.timestep, rec.layer) == (3, 0)
    counts = rec.int_counts().cpu().numpy()
    for t in range(3):
        for l, (n, d, h, s) in enumerate(sd_ffn_shapes(16)):
            E = h // 20
            assert counts[t, l, :E].sum() == s * int(E * 0.3) and counts[t, l, E:].sum() == 0
    # hooks and forward stubs are gone; a second prompt accumulates on top after reset_time_layer
    assert all(len(m._forward_hooks) == 0 and "forward" not in m.__dict__ for m in unet.modules())
    avg = helper.average_expert_counters([rec.label_counter], names, 3)
    assert abs(sum(avg[0][names[0]]) - int(64 * 0.3)) < 1e-9
    # MOEFy over the whole pipeline: masked FFNs change the result but keep it finite
    m = nr.MOEFy(0, capture_gates=False)
    out2, _ = m.observe_activation(pipe, "a photo of a cat")
    assert all(torch.isfinite(t.float()).all() for t in out2)


# ------------------------------------------------------------------ remaining receivers
def test_get_experts_hook(lib, golden_dir):
    """GetExperts (get_experts.py:50-83): top-k of the token-averaged score, all tokens and bounding-box tokens."""
    g = load(golden_dir, "get_experts_small")
    ff = make_ff(g, float(g["ratio"]), with_down=False)
    mod = ff.net[0]
    E = mod.patterns.shape[0]
    rec = nr.GetExperts(0, 1, 16, {"l": E}, ["l"] * 16)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    for tag, box in (("all", None), ("bb", g["bb"].tolist())):
        mod.bounding_box = box
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (x,), None)
        mean = rec.mean_score[0][0]
        want_mean = g[f"mean_{tag}"]
        assert np.allclose(mean, want_mean, atol=0.05)                    # bf16 inputs vs the fp32 reference
        # exact mean of the kernel's own scores (oracle on bf16-rounded inputs), labels = top-k of that vector
        pat = O.patterns_from_labels(g["labels"])
        _, labels, m16 = O.get_experts_labels(r16(T(g["x"])), r16(T(g["w1"])), T(g["b1"]), pat, mod.k, box)
        assert np.allclose(mean, m16.numpy(), atol=1e-3)                  # per-token scores agree to 5e-4 (gpu_util.SCORE_ATOL)
        srt = np.sort(m16.numpy())[::-1]
        if srt[mod.k - 1] - srt[mod.k] > 5e-3:                            # clear margin: identical label SET
            assert set(rec.label_counter[0][0]) == set(labels)
        assert rec.label_counter[0][0] == torch.topk(torch.from_numpy(mean), mod.k)[1].tolist()
    assert rel_err(unpack_cols(H, mod), T(g["H"])) < OUT_REL_TOL          # unmasked output
    mod.bounding_box = None


def test_add_experts_hook(lib, golden_dir, tmp_path):
    """AddExperts (add_skilled_experts.py:37-62): boosted experts are always selected, k' = int(0.8 k)."""
    g = load(golden_dir, "add_experts_small")
    experts, std = g["experts"].tolist(), g["std"].tolist()
    path = tmp_path / "adj" / "skilled" / "experts"
    os.makedirs(path)
    json.dump({"time_steps": {"0": {"0": {"std": std}}}}, open(tmp_path / "adj" / "predictivity_base_expert.json", "w"))
    json.dump(experts, open(path / "timestep_0_layer_0.json", "w"))
    ff = make_ff(g, float(g["ratio"]), with_down=False)
    mod = ff.net[0]
    rec = nr.AddExperts(0, str(path), 1, 1)
    H = rec.hook_fn(mod, (T(g["x"]).to(DEV, torch.bfloat16),), None)
    pat = O.patterns_from_labels(g["labels"])
    Ho, labels, gate, score = O.add_experts_forward(r16(T(g["x"])), r16(T(g["w1"])), T(g["b1"]), pat, mod.k, experts, std)
    kk = labels.shape[-1]
    safe = (O.topk_margin(score, kk) > 2e-3).numpy()
    Hc = unpack_cols(H, mod).reshape(-1, H.shape[-1])
    Hor = Ho.reshape(-1, H.shape[-1])
    assert safe.mean() > 0.85 and rel_err(Hc[safe], Hor[safe]) < OUT_REL_TOL
    # every token keeps the boosted experts' neurons (their score was raised by >= 2.5)
    live = torch.stack([(Hc[:, np.nonzero(pat[e].numpy())[0]] != 0).any(1) for e in range(pat.shape[0])], 1)
    for e in experts:
        assert bool(live[:, e].all())
    assert int(live.sum(1).max()) <= kk
    assert rec.gates and rel_err(unpack_cols(rec.gates[0].to(DEV), mod).reshape(-1, H.shape[-1])[safe],
                                 gate.reshape(-1, H.shape[-1])[safe]) < OUT_REL_TOL


def test_wanda_receiver_and_sparsity_measure(lib, golden_dir):
    """Wanda (wanda_receiver.py:37-57): column norms of the row-normalised output accumulated over two prompts;
    SparsityMeasure (sparsity_measure.py:13-18): captured activated gate, unmasked output."""
    g = load(golden_dir, "wanda_receiver_small")
    ff = make_ff(g, 1.0, with_down=False)
    mod = ff.net[0]
    mod.gelu = torch.nn.functional.relu
    rec = nr.Wanda(0, 1, 1)
    for x in g["xs"]:
        rec.reset_time_layer()
        H = rec.hook_fn(mod, (T(x).to(DEV, torch.bfloat16),), None)
    norms = rec.predictivity.get_column_norms()[0][0].numpy()
    norms = norms if mod._moe_state.weights_permuted_in_model is False else norms[mod._moe_state.layout.inv_perm.numpy()]
    assert np.allclose(norms, g["column_norms"], rtol=2e-2, atol=2e-3)     # bf16 H vs fp32 reference
    # exact restatement on the kernel's own (bf16) H of the last call
    last = O.wanda_column_sumsq(H.float().cpu())
    one = nr.Wanda(0, 1, 1)
    one.hook_fn(mod, (T(g["xs"][-1]).to(DEV, torch.bfloat16),), None)
    assert torch.allclose(one.predictivity.sumsq[(0, 0)].cpu(), last, rtol=1e-4, atol=1e-6)
    sp = nr.SparsityMeasure(0)
    Hs = sp.hook_fn(mod, (T(g["xs"][0]).to(DEV, torch.bfloat16),), None)
    assert rel_err(unpack_cols(Hs, mod), T(g["H0"])) < OUT_REL_TOL
    assert rel_err(unpack_cols(sp.gates[0].to(DEV), mod), T(g["gate0"])) < OUT_REL_TOL
    assert bool(torch.all(sp.gates[0] >= 0)) and 0.3 < sp.zero_fraction() < 0.7      # ReLU: about half exact zeros


def test_wanda_scoring_and_union_kernels_bit_exact(lib, golden_dir):
    """moe_wanda_score_mask / moe_mask_vote / bake against the fixture made from the reference's own source lines
    (modularity/wanda.py:143-165, save_union_over_time.py:189-227): integer masks, bit-exact."""
    from moefication import wanda_scoring as ws
    g = load(golden_dir, "wanda_scoring_small")
    d, h, Tn = int(g["d"]), int(g["h"]), int(g["T"])
    w2 = T(g["w2"]).to(DEV)
    bits = ws.score_masks(w2, [T(v) for v in g["norm_base"]], [T(v) for v in g["norm_adj"]], float(g["ratio"]))
    want = np.unpackbits(g["masks"], axis=-1)[..., :h].astype(int)
    for t in range(Tn):
        assert np.array_equal(ws.to_dense(bits[t], d, h), want[t]), t
    union = ws.union_over_time(bits, float(g["select_ratio"]))
    want_u = np.unpackbits(g["union"], axis=-1)[..., :h].astype(int)
    assert np.array_equal(ws.to_dense(union, d, h), want_u)
    lin = torch.nn.Linear(h, d).to(DEV, torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w2)
    ws.bake(lin, union)
    assert torch.equal(lin.weight.float().cpu(), T(g["w2"]) * torch.from_numpy(1 - want_u).float())
    assert ws.to_csr(union, d, h).nnz == int(want_u.sum())
    # ties on the k-th value go to the lowest column: a row of equal metrics keeps exactly the first k columns
    flat = torch.ones(2, 64, dtype=torch.bfloat16, device=DEV)
    nb, na = torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
    tb = ws.to_dense(ws.score_masks(flat, [nb], [na], 0.25)[0], 2, 64)
    assert tb[:, :16].all() and not tb[:, 16:].any()
