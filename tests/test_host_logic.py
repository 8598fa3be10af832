"""CPU: host-side logic of the drop-in layer (packing, helper, receivers' state machines and file
loading, statistics) -- nothing here launches a kernel."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

import moe_ffn_oracle as O
from moe_b200.packing import ExpertLayout, pack_ffn, bits_from_expert_list, bits_to_sets
from moe_b200.sd_modules import FFNStackUNet, GEGLU, FeedForward, sd_ffn_shapes
from moe_b200.stats import StatMeter
from moe_b200.ffn import attach_state, find_down_proj, activation_code, default_expert_size
from moe_b200 import ops
import neuron_receivers as nr
from moefication import helper


def test_expert_layout_and_patterns_match_reference_helper():
    labels = O.balanced_labels(160, 20, seed=3)
    lay = ExpertLayout.from_labels(labels)
    assert (lay.n_experts, lay.expert_size) == (8, 20)
    # original-order patterns == what helper.modify_ffn builds (oracle restates helper.py:50-59)
    assert torch.equal(lay.patterns(packed=False), O.patterns_from_labels(labels))
    # packed position j holds original neuron perm[j]; experts become contiguous
    assert np.array_equal(labels[lay.perm.numpy()], np.repeat(np.arange(8), 20))
    assert torch.equal(lay.perm[lay.inv_perm], torch.arange(160))
    with pytest.raises(ValueError, match="unbalanced"):
        ExpertLayout.from_labels([0, 0, 0, 1])


def test_packing_is_function_preserving():
    """Permuting W1 rows / b1 / W2 columns by expert leaves the FFN output unchanged and turns the
    reference's pattern matmul into contiguous segment sums with the same expert ids."""
    layer = O.synthetic_layer(32, 128, (2, 24), 16, seed=5)
    lay = ExpertLayout.from_labels(layer["labels"])
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"])
    w1p, b1p, w2p = p.w1p.float(), p.b1p, p.w2p.float()
    x = layer["x"]
    rt = lambda t: t.bfloat16().float()   # weights are cast to bf16 by pack_ffn
    v, g = O.geglu_up(x, rt(layer["w1"]), layer["b1"])
    vp, gp = O.geglu_up(x, w1p, b1p)
    assert torch.equal(vp, v[..., lay.perm]) and torch.equal(gp, g[..., lay.perm])
    score_ref = O.expert_scores(g, O.patterns_from_labels(layer["labels"]))
    score_seg = gp.reshape(-1, lay.n_experts, lay.expert_size).sum(-1)
    assert torch.allclose(score_ref, score_seg, atol=1e-5)
    y = O.down_proj(v * g, rt(layer["w2"]), layer["b2"])
    yp = O.down_proj(vp * gp, w2p, p.b2)
    assert torch.allclose(y, yp, atol=1e-5)


def test_bits_from_expert_list():
    b = bits_from_expert_list([0, 5, 31, 32, 63], 64)
    assert b.dtype == torch.int32 and b.shape == (2,)
    assert bits_to_sets(b.view(1, 2), 64) == [{0, 5, 31, 32, 63}]
    with pytest.raises(ValueError):
        bits_from_expert_list([64], 64)


def test_statmeter_matches_oracle_welford(tmp_path):
    rs = np.random.RandomState(0)
    sm, w = StatMeter(2, 3), O.Welford()
    for _ in range(5):
        v = rs.randn(7)
        sm.update(v, 1, 2)
        w.update(v)
    cell = sm.results["time_steps"][1][2]
    assert np.array_equal(cell["avg"].avg, w.avg) and np.array_equal(cell["std"].stddev(), w.stddev())
    sm.save(tmp_path / "s.json")
    saved = json.load(open(tmp_path / "s.json"))
    assert np.allclose(saved["time_steps"]["1"]["2"]["avg"], w.avg)


def test_unet_module_names_match_diffusers_surface():
    unet = FFNStackUNet()
    names = [n for n, m in unet.named_modules() if isinstance(m, GEGLU) and "ff.net" in n]
    assert len(names) == 16
    assert sorted(names) == sorted(n for n, _, _, _ in sd_ffn_shapes())
    assert "mid_block.attentions.0.transformer_blocks.0.ff.net.0" in names
    assert "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0" in names
    # sorted weight names == firing order down -> mid -> up (reference helper.py:76-77)
    firing = [n + ".proj.weight" for n, _, _, _ in sd_ffn_shapes()]
    assert sorted(firing) == firing
    assert sum(s for _, _, _, s in sd_ffn_shapes(64)) == 5 * 4096 + 5 * 1024 + 5 * 256 + 64
    assert isinstance(find_down_proj(unet, names[0]), torch.nn.Linear)


class _Args:
    res_path = ""
    moefication = {"topk_experts": 0.3}


class _Pipe:
    def __init__(self, unet):
        self.unet = unet


def _tiny_unet():
    """A 2-FFN module tree with diffusers-style names (CPU, fp32)."""
    import torch.nn as nn

    class Blk(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.ff = FeedForward(d)

    class U(nn.Module):
        def __init__(self):
            super().__init__()
            self.transformer_blocks = nn.ModuleList([Blk(32), Blk(40)])
    return U()


def test_modify_ffn_to_experts_permutes_model_consistently(tmp_path):
    torch.manual_seed(0)
    unet = _tiny_unet()
    pipe = _Pipe(unet)
    x0, x1 = torch.randn(2, 5, 32), torch.randn(2, 5, 40)
    y_before = [unet.transformer_blocks[0].ff(x0), unet.transformer_blocks[1].ff(x1)]
    os.makedirs(tmp_path / "param_split")
    labels = {}
    for name, m in unet.named_modules():
        if isinstance(m, GEGLU):
            h = m.proj.weight.shape[0] // 2
            lab = O.balanced_labels(h, 16 if h == 128 else 20, seed=h)
            labels[name + ".proj.weight"] = lab
            torch.save([int(v) for v in lab], tmp_path / "param_split" / (name + ".proj.weight"))
    args = _Args()
    args.res_path = str(tmp_path)
    _, layer_names, n_exp = helper.modify_ffn_to_experts(pipe, args)
    assert layer_names == sorted(labels) and list(n_exp.values()) == [8, 8]
    g0 = unet.transformer_blocks[0].ff.net[0]
    assert g0.k == int(8 * 0.3) == 2 and g0.patterns.shape == (8, 128)
    # packed patterns are block one-hot; the module still computes the same function
    assert torch.equal(g0.patterns, torch.eye(8).repeat_interleave(16, dim=1))
    y_after = [unet.transformer_blocks[0].ff(x0), unet.transformer_blocks[1].ff(x1)]
    for a, b in zip(y_before, y_after):
        assert torch.allclose(a, b, atol=1e-5)
    st = g0._moe_state
    assert st.weights_permuted_in_model and st.w1p.dtype == torch.bfloat16 and st.b1p.dtype == torch.float32
    assert st.w2p.shape == (32, 128) and st.k == 2
    # oracle MoE forward on the ORIGINAL labels == oracle on packed weights with block patterns
    assert unet.transformer_blocks[0].ff.net[2]._moe_column_perm is not None


def test_activation_code_detects_relufied_module():
    m = GEGLU(8, 16)
    assert activation_code(m) == ops.ACT_GELU
    m.gelu = torch.nn.functional.relu          # reference sparsity/relufy_model.py:35
    assert activation_code(m) == ops.ACT_RELU
    m.gelu = lambda g: torch.tanh(g)
    with pytest.raises(ValueError, match="neither"):
        activation_code(m)
    assert default_expert_size(1280) == 64 and default_expert_size(40) == 8


def test_time_layer_state_machines():
    names = [f"l{i}" for i in range(16)]
    fm = nr.FrequencyMeasure(0, 3, 16, {n: 8 for n in names}, names, device="cpu")
    ep = nr.ExpertPredictivity(0, 3, 16)
    clock = O.TimeLayerClock(16)
    for _ in range(35):
        fm.update_time_layer(); ep.update_time_layer(); clock.tick()
        assert (fm.timestep, fm.layer) == (ep.timestep, ep.layer) == (clock.timestep, clock.layer)
    assert (fm.timestep, fm.layer) == (2, 3)
    fm.reset_time_layer()
    assert (fm.timestep, fm.layer) == (0, 0)
    fm.reset()
    lc = fm.label_counter
    assert set(lc) == {0, 1, 2} and lc[0][15].shape == (8,) and lc[0][15].dtype == np.float64
    assert fm.int_counts().shape == (3, 16, 8) and fm.int_counts().dtype == torch.int64


def test_removal_receivers_load_reference_file_formats(tmp_path):
    import scipy.sparse as sp
    ed, nd, wd = tmp_path / "e", tmp_path / "n", tmp_path / "w"
    for p in (ed, nd, wd):
        os.makedirs(p)
    for t in range(2):
        for l in range(2):
            json.dump([1, 3] if l == 0 else [], open(ed / f"timestep_{t}_layer_{l}.json", "w"))
            json.dump([0.0, 1.0] * 8, open(nd / f"predictivity_{t}_{l}.json", "w"))
            with open(wd / f"timestep_{t}_layer_{l}.pkl", "wb") as f:
                pickle.dump(sp.csr_matrix(np.eye(4, 32, dtype=np.int64)), f)
    re_ = nr.RemoveExperts(0, str(ed), 2, 2)
    assert re_.expert_indices[1][0] == [1, 3] and re_.expert_indices[0][1] == []
    assert re_._removed_bits(8, "cpu").tolist() == [0b1010]
    re_.layer = 1
    assert re_._removed_bits(8, "cpu") is None          # empty list
    re_.layer, re_.timestep = 0, 20
    assert re_.timestep >= 20 and re_.expert_indices.get(20) is None
    rn = nr.RemoveNeurons(0, str(nd), 2, 2)
    assert rn.expert_indices[0][0][:4] == [0.0, 1.0, 0.0, 1.0]
    for _ in range(3):
        rn.update_time_layer()
    assert (rn.timestep, rn.layer) == (1, 1)
    wr = nr.WandaRemoveNeuronsFast(0, str(wd), 2, 2, remove_timesteps=None, weights_shape=None)
    assert wr.expert_indices[1][1].shape == (4, 32)
    mc = nr.MultiConceptRemoverWanda(str(tmp_path) + "/%s_%s", 0, 2, 2, concepts_to_remove=[], wanda_thr={})
    assert mc.removers == {}


def test_removed_bits_respects_timestep_rule(tmp_path):
    os.makedirs(tmp_path / "e")
    for t in range(22):
        json.dump([2], open(tmp_path / "e" / f"timestep_{t}_layer_0.json", "w"))
    r = nr.RemoveExperts(0, str(tmp_path / "e"), 22, 1)
    r.timestep = 19
    assert r._removed_bits(8, "cpu").tolist() == [4]
    r.timestep = 20                                  # hard-coded `timestep < 20` (remove_skilled_experts.py:32)
    assert r._removed_bits(8, "cpu") is None


def test_hook_lifecycle_stubs_and_restores_stock_forward():
    unet = _tiny_unet()
    pipe = _Pipe(unet)
    rec = nr.MOEFy(0)
    hooks = rec.register_hooks(pipe)
    mods = [m for _, m in unet.named_modules() if isinstance(m, GEGLU)]
    assert len(hooks) == 2 and all("forward" in m.__dict__ for m in mods)
    assert all(m.bounding_box is None for m in mods)
    rec.remove_hooks(hooks)
    assert all("forward" not in m.__dict__ for m in mods) and all(len(m._forward_hooks) == 0 for m in mods)
    # stock forward is back
    assert unet.transformer_blocks[0].ff(torch.randn(1, 3, 32)).shape == (1, 3, 32)
    wr = nr.WandaRemoveNeuronsFast(0, None, 1, 2)
    assert [n for n, _ in wr._select_modules(pipe)] == ["transformer_blocks.0.ff.net.2", "transformer_blocks.1.ff.net.2"]


def test_average_expert_counters_matches_reference_accumulation():
    rs = np.random.RandomState(1)
    names = ["a", "b"]
    per_image = [{t: {i: rs.rand(4) for i in range(2)} for t in range(3)} for _ in range(5)]
    got = helper.average_expert_counters(per_image, names, 3)
    want = O.average_counters(per_image, names, 3)
    for t in range(3):
        for n in names:
            assert got[t][n] == want[t][n]


def _moe_reference_output(ff_dense, labels, ratio, x):
    """The oracle's MoEfied FFN on the ORIGINAL (unpermuted) weights of a FeedForward."""
    g, lin = ff_dense.net[0], ff_dense.net[2]
    pat = O.patterns_from_labels(labels)
    H, lab, _, _ = O.moefy_forward(x, g.proj.weight.detach(), g.proj.bias.detach(), pat, int(pat.shape[0] * ratio))
    return O.down_proj(H, lin.weight.detach(), lin.bias.detach()), lab


def test_modify_ffn_without_down_projection_leaves_the_model_alone():
    """`helper.modify_ffn(ffn, path, k)` -- the reference's exact signature, no down-projection handed
    over -- must not permute W1 in place (ff.net.2 would then meet packed H with unpermuted columns)."""
    torch.manual_seed(1)
    ff = FeedForward(32)
    w1_before = ff.net[0].proj.weight.detach().clone()
    labels = O.balanced_labels(128, 16, seed=2)
    st = helper.modify_ffn(ff.net[0], [int(v) for v in labels], 0.5)
    assert not st.weights_permuted_in_model and st.applied_perm is None
    assert torch.equal(ff.net[0].proj.weight, w1_before)
    # patterns refer to the module's (original) neuron order, exactly the reference's one-hot matrix
    assert torch.equal(ff.net[0].patterns, O.patterns_from_labels(labels))
    # the kernels' packed copy has expert e's neurons at [16 e, 16 e + 16)
    lay = st.layout
    assert torch.equal(st.w1p[:128].float(), w1_before[:128][lay.perm].bfloat16().float())


def test_modify_ffn_twice_is_idempotent_on_the_model():
    """re-attaching (new top-k / new labels) must not permute already-packed weights again."""
    torch.manual_seed(2)
    ff = FeedForward(32)
    x = torch.randn(2, 7, 32)
    y_dense = ff(x)
    w1_orig = ff.net[0].proj.weight.detach().clone()
    w2_orig = ff.net[2].weight.detach().clone()
    lab_a = [int(v) for v in O.balanced_labels(128, 16, seed=3)]
    lab_b = [int(v) for v in O.balanced_labels(128, 16, seed=4)]
    helper.modify_ffn(ff.net[0], lab_a, 0.5, down=ff.net[2])
    helper.modify_ffn(ff.net[0], lab_a, 0.25, down=ff.net[2])         # same labels, new ratio
    st = ff.net[0]._moe_state
    lay = st.layout
    assert ff.net[0].k == 2 and st.k == 2 and st.weights_permuted_in_model
    assert torch.equal(ff.net[0].proj.weight[:128], w1_orig[:128][lay.perm])      # packed ONCE from the original
    assert torch.equal(ff.net[2].weight, w2_orig[:, lay.perm])
    assert torch.allclose(ff(x), y_dense, atol=1e-5)
    helper.modify_ffn(ff.net[0], lab_b, 0.5, down=ff.net[2])          # different labels
    lay_b = ff.net[0]._moe_state.layout
    assert torch.equal(ff.net[0].proj.weight[:128], w1_orig[:128][lay_b.perm])
    assert torch.equal(ff.net[2]._moe_column_perm, lay_b.perm)
    assert torch.allclose(ff(x), y_dense, atol=1e-5)
    from moe_b200.ffn import undo_model_permutation
    undo_model_permutation(ff.net[0])
    assert torch.equal(ff.net[0].proj.weight, w1_orig) and torch.equal(ff.net[2].weight, w2_orig)


def test_packed_copies_follow_parameter_updates():
    """fp32 parameters are copied to bf16 for the kernels; an in-place update must not leave a stale copy."""
    from moe_b200.ffn import get_state
    torch.manual_seed(3)
    ff = FeedForward(32)
    helper.modify_ffn(ff.net[0], [int(v) for v in O.balanced_labels(128, 16, seed=5)], 0.5, down=ff.net[2])
    st = get_state(ff.net[0])
    before = st.w2p.clone()
    with torch.no_grad():
        ff.net[2].weight.mul_(0.5)
    st2 = get_state(ff.net[0])
    assert st2 is st and torch.equal(st.w2p.float(), (before.float() * 0.5).bfloat16().float())
    # activation probe is cached on the identity of module.gelu and redone when it is swapped
    assert st.act == ops.ACT_GELU
    ff.net[0].gelu = torch.nn.functional.relu
    assert get_state(ff.net[0]).act == ops.ACT_RELU


def test_fused_down_projection_patch_is_undone():
    """While hooked, ff.net.2 of a CUDA model is a pass-through; on a CPU model (no kernels) nothing is patched."""
    unet = _tiny_unet()
    rec = nr.MOEFy(0, capture_gates=False)
    hooks = rec.register_hooks(_Pipe(unet))
    downs = [unet.transformer_blocks[i].ff.net[2] for i in range(2)]
    assert all("forward" not in d.__dict__ for d in downs)           # CPU weights: the reference flow is kept
    rec.remove_hooks(hooks)
    assert rec._fused_states == [] and all(not m._moe_state.fused_down for m in unet.modules() if isinstance(m, GEGLU))


def test_balanced_kmeans_split_meets_the_reference_contract(tmp_path):
    """SURVEY 8f row 4 (moe_utils.py:97-107): equally sized experts from the gate half of W1, saved in the reference's
    label-file format and consumed by helper.modify_ffn_to_experts."""
    from moefication import moe_utils
    torch.manual_seed(0)
    unet = _tiny_unet()
    # plant structure: 8 groups of 16 gate rows around 8 directions, shuffled -> the split must find the groups
    g0 = unet.transformer_blocks[0].ff.net[0]
    dirs = torch.nn.functional.normalize(torch.randn(8, 32), dim=1)
    perm = torch.randperm(128)
    with torch.no_grad():
        g0.proj.weight[128:][perm] = (dirs.repeat_interleave(16, 0) + 0.05 * torch.randn(128, 32))
    labels = moe_utils.split_ffn_weight(g0.proj.weight, 16, seed=0)
    assert len(labels) == 128 and np.bincount(labels).tolist() == [16] * 8
    assert labels == moe_utils.split_ffn_weight(g0.proj.weight, 16, seed=0)                 # deterministic
    planted = torch.empty(128, dtype=torch.long)
    planted[perm] = torch.arange(8).repeat_interleave(16)
    for e in range(8):                                          # every expert is exactly one planted group
        assert len(set(planted[[i for i, l in enumerate(labels) if l == e]].tolist())) == 1
    rnd = [i // 16 for i in range(128)]
    assert moe_utils.inertia(g0.proj.weight[128:], labels) < 0.2 * moe_utils.inertia(g0.proj.weight[128:], rnd)
    with pytest.raises(ValueError, match="divisible"):
        moe_utils.balanced_kmeans(torch.randn(30, 4), 8)
    # whole-model driver -> label files -> the reference's loading path
    args = _Args()
    args.res_path = str(tmp_path)
    written = moe_utils.moefy_sd_model(_Pipe(unet), str(tmp_path), expert_size=16 if False else 8)
    assert sorted(written) == sorted(n + ".proj.weight" for n, m in unet.named_modules() if isinstance(m, GEGLU))
    _, names, n_exp = helper.modify_ffn_to_experts(_Pipe(unet), args)
    assert list(n_exp.values()) == [16, 20] and unet.transformer_blocks[0].ff.net[0].expert_size == 8
