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
    gc = rec.gates[0].float().reshape(n, -1)                 # captured gates are in the ORIGINAL neuron order
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
    assert rec.gates and rel_err(rec.gates[0].float().reshape(-1, H.shape[-1])[safe],
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
    norms = rec.predictivity.get_column_norms()[0][0].numpy()        # original neuron order
    assert np.allclose(norms, g["column_norms"], rtol=2e-2, atol=2e-3)     # bf16 H vs fp32 reference
    # exact restatement on the kernel's own (bf16) H of the last call
    last = O.wanda_column_sumsq(H.float().cpu())
    one = nr.Wanda(0, 1, 1)
    one.hook_fn(mod, (T(g["xs"][-1]).to(DEV, torch.bfloat16),), None)
    assert torch.allclose(one.predictivity.sumsq[(0, 0)].cpu(), last, rtol=1e-4, atol=1e-6)   # (device cells: packed order)
    sp = nr.SparsityMeasure(0)
    Hs = sp.hook_fn(mod, (T(g["xs"][0]).to(DEV, torch.bfloat16),), None)
    assert rel_err(unpack_cols(Hs, mod), T(g["H0"])) < OUT_REL_TOL
    assert rel_err(sp.gates[0].float(), T(g["gate0"])) < OUT_REL_TOL
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


# ------------------------------------------------------------------ round 2: the fused kernel BEHIND the hook API
class _OneBlockUNet(torch.nn.Module):
    """`transformer_blocks.0.ff.net.{0,2}`: the names the receivers filter on, around ONE FeedForward."""

    def __init__(self, ff):
        super().__init__()
        blk = torch.nn.Module()
        blk.ff = ff
        self.transformer_blocks = torch.nn.ModuleList([blk])

    def forward(self, x):
        return self.transformer_blocks[0].ff(x)


class _OnePipe:
    """Pipeline stand-in: `pipe(x_list)` runs the FFN once per entry (= layer calls of consecutive UNet steps)."""

    class Out:
        def __init__(self, images):
            self.images = images

    def __init__(self, ff):
        self.unet = _OneBlockUNet(ff)

    def __call__(self, xs, **kw):
        return _OnePipe.Out([[self.unet(x) for x in xs]])


def _kernel_names(fn):
    """CUDA kernel names launched by fn() (CUPTI through torch.profiler)."""
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    return [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]


@pytest.mark.parametrize("name", ["moefy_d64_gelu", "moefy_d64_relu", "moefy_es20", "moefy_d128_es64"])
def test_hooked_ffn_is_one_fused_launch_and_matches_reference(lib, golden_dir, name):
    """through `observe_activation` a routed receiver issues exactly ONE user kernel per layer call
    (moe_ffn_fused) -- no K1/K2/K3 triple, no cuBLAS down-projection -- and the FFN output matches the reference's."""
    import moe_b200 as M
    g = load(golden_dir, name)
    ff = make_ff(g, float(g["ratio"]))
    if int(g["act"]) == O.ACT_RELU:
        ff.net[0].gelu = torch.nn.functional.relu
    pipe = _OnePipe(ff)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    rec = nr.MOEFy(seed=0, capture_gates=False)
    rec.observe_activation(pipe, [x])                       # warm-up: workspace allocation, kernel attributes
    M.reset_launch_count()
    out, gates = rec.observe_activation(pipe, [x, x])
    torch.cuda.synchronize()
    assert M.launch_count() == 2 and gates == []            # one launch per layer call
    y = out[0].float().cpu()
    wide = (g["margin"] > 0.25).reshape(-1)
    n = wide.shape[0]
    assert y.shape == tuple(g["y"].shape)
    assert rel_err(y.reshape(n, -1)[wide], T(g["y"]).reshape(n, -1)[wide]) < OUT_REL_TOL
    # the patch is gone: ff.net.2 is the stock Linear again
    assert "forward" not in ff.net[2].__dict__ and "forward" not in ff.net[0].__dict__
    assert not ff.net[0]._moe_state.fused_down
    # kernel list of one hooked call: exactly one of ours, nothing from cuBLAS / CUTLASS
    try:
        names = _kernel_names(lambda: rec.observe_activation(pipe, [x]))
    except Exception as e:      # noqa: BLE001 -- CUPTI unavailable (e.g. another profiler attached)
        pytest.skip(f"torch.profiler could not collect CUDA kernels: {e}")
    if not names:
        pytest.skip("torch.profiler returned no CUDA kernels (CUPTI unavailable)")
    ours = [k for k in names if "ffn_fused_kernel" in k]
    lib_gemm = [k for k in names if any(s in k.lower() for s in ("gemm", "cutlass", "cublas", "xmma", "nvjet"))]
    assert len(ours) == 1 and lib_gemm == [], names


def test_hooked_ffn_with_gate_capture_stays_native(lib, golden_dir):
    """capture_gates (the reference default) needs the activated gate as a tensor: K1 -> K2 -> K2(gate) -> K3, all
    ours; the gates come back on the host in original neuron order."""
    import moe_b200 as M
    g = load(golden_dir, "moefy_d64_gelu")
    ff = make_ff(g, float(g["ratio"]))
    pipe = _OnePipe(ff)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    rec = nr.MOEFy(seed=0)
    rec.observe_activation(pipe, [x])
    M.reset_launch_count()
    out, gates = rec.observe_activation(pipe, [x])
    assert M.launch_count() == 4
    wide = (g["margin"] > 0.25).reshape(-1)
    n = wide.shape[0]
    assert rel_err(out[0].float().cpu().reshape(n, -1)[wide], T(g["y"]).reshape(n, -1)[wide]) < OUT_REL_TOL
    assert len(gates) == 1 and gates[0].device.type == "cpu" and gates[0].is_pinned()
    assert rel_err(gates[0].float().reshape(n, -1)[wide], T(g["gate"]).reshape(n, -1)[wide]) < OUT_REL_TOL
    # fused and split paths agree on the FFN output
    rec2 = nr.MOEFy(seed=0, capture_gates=False)
    out2, _ = rec2.observe_activation(pipe, [x])
    assert rel_err(out2[0].float(), out[0].float()) < 2e-3


def test_frequency_and_removal_through_fused_hooks(lib, golden_dir, tmp_path):
    """FrequencyMeasure / RemoveExperts through observe_activation: the fused launch carries the histogram and the
    removed-expert bits; counts equal the unfused path's bit for bit, the FFN output matches the reference's
    RemoveExperts.hook_fn followed by the stock down-projection (golden y_t*_l0)."""
    import moe_b200 as M
    g = load(golden_dir, "remove_experts_d64")
    removed = [int(v) for v in g["removed"]]
    times = [0, 19, 20]                                       # < 20 removes, >= 20 does not
    for t in range(21):
        json.dump(removed, open(tmp_path / f"timestep_{t}_layer_0.json", "w"))
    ff = make_ff(g, float(g["ratio"]))
    mod = ff.net[0]
    E = mod.patterns.shape[0]
    pipe = _OnePipe(ff)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    ntok = x.shape[0] * x.shape[1]
    counts, outs = {}, {}
    for fuse in (True, False):
        hist = torch.zeros(21, 1, E, dtype=torch.int64, device=DEV)
        rec = nr.RemoveExperts(0, str(tmp_path), 21, 1, capture_gates=False, hist=hist, count_rows='all', fuse_down_proj=fuse)
        rec.observe_activation(pipe, [x])                      # warm-up
        hist.zero_()
        M.reset_launch_count()
        rec.reset_time_layer()
        out, _ = rec.observe_activation(pipe, [x] * 21)
        assert (rec.timestep, rec.layer) == (21, 0)
        assert M.launch_count() == (21 if fuse else 42)        # fused: one launch; unfused: K1 + K2 (+ cuBLAS)
        counts[fuse] = hist.cpu().numpy()
        outs[fuse] = [o.float().cpu() for o in out]
    assert np.array_equal(counts[True], counts[False])
    assert counts[True].sum(-1).reshape(-1).tolist() == [ntok * mod.k] * 21
    for t in times:
        score = T(g[f"score_t{t}_l0"])
        wide = (O.topk_margin(score, mod.k) > 0.25).numpy()
        assert wide.mean() > 0.3
        yr = T(g[f"y_t{t}_l0"]).reshape(ntok, -1)
        for fuse in (True, False):
            assert rel_err(outs[fuse][t].reshape(ntok, -1)[wide], yr[wide]) < (OUT_REL_TOL if fuse else 1.5e-2), (t, fuse)
        # oracle on the bf16-rounded inputs, every token whose margin exceeds 2e-3
        pat = O.patterns_from_labels(g["labels"])
        H, _, _, sc = O.remove_experts_forward(r16(T(g["x"])), r16(T(g["w1"])), T(g["b1"]), pat, mod.k, removed, t)
        safe = (O.topk_margin(sc, mod.k) > 2e-3).numpy()
        yo = O.down_proj(r16(H), r16(T(g["w2"])), T(g["b2"])).reshape(ntok, -1)
        assert rel_err(outs[True][t].reshape(ntok, -1)[safe], yo[safe]) < OUT_REL_TOL
    names = ["l0"]
    fm = nr.FrequencyMeasure(0, 3, 1, {"l0": E}, names)
    fm.observe_activation(pipe, [x, x, x])
    fm2 = nr.FrequencyMeasure(0, 3, 1, {"l0": E}, names, fuse_down_proj=False)
    fm2.observe_activation(pipe, [x, x, x])
    assert torch.equal(fm.int_counts(), fm2.int_counts()) and int(fm.int_counts()[0].sum()) == x.shape[1] * mod.k


def test_unrouted_receivers_use_the_native_down_projection(lib, golden_dir):
    """ExpertPredictivity / NeuronPredictivity through observe_activation: K1 (+ statistic) + native K3 -> Y."""
    import moe_b200 as M
    g = load(golden_dir, "moefy_d64_gelu")
    ff = make_ff(g, float(g["ratio"]))
    pipe = _OnePipe(ff)
    x = T(g["x"]).to(DEV, torch.bfloat16)
    dense = ff(x).float()
    rec = nr.ExpertPredictivity(0, 1, 1)
    out, _ = rec.observe_activation(pipe, [x])
    assert out[0].shape == dense.shape and rel_err(out[0].float(), dense) < 1.5e-2       # unmasked FFN
    rec = nr.NeuronPredictivity(0, 1, 1)
    M.reset_launch_count()
    out, _ = rec.observe_activation(pipe, [x])
    assert M.launch_count() == 3                                   # K1, colmax, K3
    assert rel_err(out[0].float(), dense) < 1.5e-2


def test_neuron_artefacts_round_trip_on_a_permuted_model(lib, tmp_path):
    """per-neuron results must come back in ORIGINAL neuron order on a model MoEfied with
    permute_model_weights=True, so that NeuronPredictivity -> JSON -> RemoveNeurons ablates the right neurons."""
    torch.manual_seed(0)
    d, h, S = 64, 256, 96
    layer = O.synthetic_layer(d, h, (2, S), 16, seed=11)
    g = dict(w1=layer["w1"].numpy(), b1=layer["b1"].numpy(), w2=layer["w2"].numpy(), b2=layer["b2"].numpy(),
             labels=layer["labels"])
    ff = make_ff(g, 0.5)
    assert ff.net[0]._moe_state.weights_permuted_in_model and not ff.net[0]._moe_state.layout.is_identity
    pipe = _OnePipe(ff)
    x = layer["x"].to(DEV, torch.bfloat16)
    x16, w1, w2 = r16(layer["x"]), r16(layer["w1"]), r16(layer["w2"])
    rec = nr.NeuronPredictivity(0, 1, 1)
    rec.observe_activation(pipe, [x])
    _, gate = O.geglu_up(x16, w1, layer["b1"])
    want_max = O.neuron_predictivity(gate)
    got_max = np.asarray(rec.max_gate[0][0])
    assert np.allclose(got_max, want_max, atol=2e-2, rtol=2e-2)             # ORIGINAL order
    # the "skilled" neurons: the 24 with the largest max activation (stand-in for skilled_neuron_ap.py:171-177)
    flags = np.zeros(h)
    flags[np.argsort(got_max)[-24:]] = 1.0
    json.dump(flags.tolist(), open(tmp_path / "predictivity_0_0.json", "w"))
    rn = nr.RemoveNeurons(0, str(tmp_path), 1, 1)
    out, gates = rn.observe_activation(pipe, [x])
    Ho = O.remove_neurons_forward(x16, w1, layer["b1"], flags.tolist())
    Ho = Ho[0] if isinstance(Ho, tuple) else Ho
    yo = O.down_proj(r16(Ho), w2, layer["b2"])
    assert rel_err(out[0].float().cpu(), yo) < OUT_REL_TOL
    idx = np.nonzero(flags)[0]
    assert torch.all(gates[0][..., idx].float() == torch.tensor(-0.17).bfloat16().float())       # original order


def test_wanda_artefacts_round_trip_on_a_permuted_model(lib, tmp_path):
    """Wanda norms -> score_masks -> CSR pickle -> WandaRemoveNeuronsFast on a permuted model
    masks the same weights the reference would (all artefacts in original column order)."""
    from moefication import wanda_scoring as ws
    torch.manual_seed(1)
    d, h, S = 64, 256, 96
    layer = O.synthetic_layer(d, h, (2, S), 16, seed=12)
    adj = O.synthetic_layer(d, h, (2, S), 16, seed=13)
    g = dict(w1=layer["w1"].numpy(), b1=layer["b1"].numpy(), w2=layer["w2"].numpy(), b2=layer["b2"].numpy(),
             labels=layer["labels"])
    ff = make_ff(g, 0.5)
    pipe = _OnePipe(ff)
    w1, w2 = r16(layer["w1"]), r16(layer["w2"])
    norms = {}
    for tag, xin in (("base", layer["x"]), ("adj", adj["x"] * 1.5)):
        rec = nr.Wanda(0, 1, 1)
        rec.observe_activation(pipe, [xin.to(DEV, torch.bfloat16)])
        norms[tag] = rec.predictivity.get_column_norms()[0][0]
        v, gt = O.geglu_up(r16(xin), w1, layer["b1"])
        want = torch.sqrt(O.wanda_column_sumsq((v * gt).reshape(-1, h)))
        assert torch.allclose(norms[tag], want, rtol=3e-2, atol=3e-3)              # ORIGINAL order
    bits = ws.score_masks(ff.net[2], [norms["base"].to(DEV)], [norms["adj"].to(DEV)], 0.05)
    dense = ws.to_dense(bits[0], d, h)
    # the oracle's mask from the ORIGINAL weight and the receiver's norms: identical bits where the row top-k is not tied
    want_mask = O.wanda_score_mask(w2.abs(), norms["base"], norms["adj"], 0.05)
    assert (dense != want_mask).mean() < 2e-3
    with open(tmp_path / "timestep_0_layer_0.pkl", "wb") as f:
        pickle.dump(ws.to_csr(bits[0], d, h), f)
    wr = nr.WandaRemoveNeuronsFast(0, str(tmp_path), 1, 1)
    x = layer["x"].to(DEV, torch.bfloat16)
    out, _ = wr.observe_activation(pipe, [x])
    v, gt = O.geglu_up(r16(layer["x"]), w1, layer["b1"])
    yo = O.wanda_down_proj(r16(v * gt), w2, layer["b2"], torch.from_numpy(dense))
    assert rel_err(out[0][0].float().cpu(), yo) < 1.5e-2        # stock (cuBLAS) GEGLU + native masked down-projection
    y_unmasked = O.down_proj(r16(v * gt), w2, layer["b2"])
    assert rel_err(out[0][0].float().cpu(), y_unmasked) > rel_err(out[0][0].float().cpu(), yo) * 2


def test_wanda_mask_modes_agree_and_cache_makes_it_one_launch(lib, golden_dir, tmp_path):
    """WandaRemoveNeuronsFast: 'cache' (resident masked weights), 'copy' (mask_weights + K3) and 'fused' (mask applied
    in shared memory) give the same bits; from the second prompt on 'cache' is ONE launch per layer call."""
    import moe_b200 as M
    import scipy.sparse as sp
    g = load(golden_dir, "wanda_csv_320_1280")
    d, h = 320, 1280
    rs = np.random.RandomState(0)
    mask = (rs.rand(d, h) < 0.03).astype(np.int64)
    for t in range(2):
        with open(tmp_path / f"timestep_{t}_layer_0.pkl", "wb") as f:
            pickle.dump(sp.csr_matrix(mask if t == 0 else 1 - mask), f)
    ff = FeedForward(d).to(DEV, torch.bfloat16)
    pipe = _OnePipe(ff)
    x = torch.randn(2, 200, d, device=DEV, dtype=torch.bfloat16)
    outs = {}
    for mode in ("cache", "copy", "fused"):
        wr = nr.WandaRemoveNeuronsFast(0, str(tmp_path), 2, 1, mask_mode=mode)
        out, _ = wr.observe_activation(pipe, [x, x])
        outs[mode] = [o.clone() for o in out[0]]
        wr.reset_time_layer()
        M.reset_launch_count()
        out2, _ = wr.observe_activation(pipe, [x, x])
        torch.cuda.synchronize()
        assert M.launch_count() == {"cache": 2, "copy": 4, "fused": 2}[mode]
        assert all(torch.equal(a, b) for a, b in zip(outs[mode], out2[0]))
    for mode in ("copy", "fused"):
        assert all(torch.equal(a, b) for a, b in zip(outs["cache"], outs[mode]))
    Hd = ff.net[0](x).float().cpu()
    ref0 = torch.nn.functional.linear(r16(Hd), ff.net[2].weight.float().cpu() * torch.from_numpy(1 - mask).float(), ff.net[2].bias.float().cpu())
    assert rel_err(outs["cache"][0].float().cpu(), ref0) < 1.5e-2
    # an in-place weight update invalidates the resident copy
    wr = nr.WandaRemoveNeuronsFast(0, str(tmp_path), 2, 1)
    a, _ = wr.observe_activation(pipe, [x])
    with torch.no_grad():
        ff.net[2].weight.mul_(2.0)
    wr.reset_time_layer()
    b, _ = wr.observe_activation(pipe, [x])
    assert rel_err(b[0][0].float() - ff.net[2].bias.float(), 2 * (a[0][0].float() - ff.net[2].bias.float())) < 2e-2
