"""GPU: the remaining BASELINE.json configurations as parity / property cases (configs[2], [3]-single-rank, [4])."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

import moe_b200 as M
import moe_ffn_oracle as O
import neuron_receivers as nr
from moefication import helper, freq_expert_select
from moe_b200.sd_modules import FFNStackUNet, SyntheticFFNPipeline, sd_ffn_shapes, GEGLU
from gpu_util import DEV

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


class Args:
    res_path = ""
    moefication = {"topk_experts": 0.3}


def moefied_pipeline(latent_hw, steps, ratio=0.3, es=20):
    torch.manual_seed(0)
    unet = FFNStackUNet(latent_hw=latent_hw)
    pipe = SyntheticFFNPipeline(unet, num_inference_steps=steps, device=DEV)
    shapes = sd_ffn_shapes(latent_hw)
    labels = {n + ".proj.weight": O.balanced_labels(h, es, seed=i) for i, (n, d, h, s) in enumerate(shapes)}
    a = Args()
    a.moefication = {"topk_experts": ratio}
    pipe, names, n_exp = helper.modify_ffn_to_experts(pipe, a, labels_by_name=labels)
    return pipe, names, n_exp, shapes


def test_config3_ddim50_batch8_removal_with_counters(lib, tmp_path):
    """BASELINE configs[2]: 50 steps, 8 prompts (UNet batch 16), RemoveExperts masks + per-timestep counters."""
    steps, n_prompts = 50, 8
    pipe, names, n_exp, shapes = moefied_pipeline(64, steps)
    rs = np.random.RandomState(2)
    removed = {}
    for t in range(steps):
        for l, (n, d, h, s) in enumerate(shapes):
            E = h // 20
            lst = rs.choice(E, E // 10, replace=False).tolist()
            removed[(t, l)] = lst
            json.dump(lst, open(tmp_path / f"timestep_{t}_layer_{l}.json", "w"))
    hist = torch.zeros(steps, 16, 256, dtype=torch.int64, device=DEV)
    rec = nr.RemoveExperts(0, str(tmp_path), steps, 16, capture_gates=False, hist=hist, count_rows="all")
    rec.reset_time_layer()
    out, _ = rec.observe_activation(pipe, ["a photo of a cat"] * n_prompts)
    assert (rec.timestep, rec.layer) == (steps, 0)
    counts = hist.cpu().numpy()
    for l, (n, d, h, s) in enumerate(shapes):
        E = h // 20
        k = int(E * 0.3)
        # every (timestep, layer) cell: all 16 batch rows x S tokens select exactly k experts
        assert (counts[:, l, :E].sum(-1) == 2 * n_prompts * s * k).all()
        assert (counts[:, l, E:] == 0).all()
        # removed experts score exactly 0 for t < 20 (GELU scores of live experts are almost always > 0), so they
        # are (almost) never selected there, and selected at the usual rate from t = 20 on
        early = np.mean([counts[t, l, removed[(t, l)]].mean() for t in range(20)])
        late = np.mean([counts[t, l, removed[(t, l)]].mean() for t in range(20, steps)])
        mean_rate = 2 * n_prompts * s * k / E
        assert early < 0.05 * mean_rate and 0.5 * mean_rate < late < 1.5 * mean_rate
    assert all(torch.isfinite(x.float()).all() for x in out)


def test_freq_expert_select_driver_writes_reference_schema(lib, tmp_path):
    """Row a11: counters averaged over prompts, keyed by sorted FFN name, JSON schema of the reference."""
    pipe, names, n_exp, shapes = moefied_pipeline(16, 2)
    prompts = ["p0", "p1", "p2"]
    counter = freq_expert_select.run(pipe, prompts, seed=0, timesteps=2, n_layers=16, ffn_names=names,
                                     num_experts_per_ffn=n_exp, topk=0.3, save_path=str(tmp_path))
    saved = json.load(open(tmp_path / "expert_counter_0.3.json"))
    assert set(saved) == {"0", "1"} and set(saved["0"]) == set(names)
    for name in names:
        E = n_exp[name]
        assert len(saved["0"][name]) == E
        # mean over images of (selecting tokens / S) sums to k per layer
        assert abs(sum(counter[0][name]) - int(E * 0.3)) < 1e-9
    # sharded over 2 "ranks" without a process group == single process (integer sums are associative)
    c0 = freq_expert_select.run(pipe, prompts[0::2], 0, 2, 16, names, n_exp, 0.3)
    c1 = freq_expert_select.run(pipe, prompts[1::2], 0, 2, 16, names, n_exp, 0.3)
    for name in names[:3]:
        merged = (2 * np.array(c0[1][name]) + 1 * np.array(c1[1][name])) / 3
        assert np.allclose(merged, counter[1][name], atol=1e-12)


@pytest.mark.parametrize("ratio", [0.1, 0.2, 0.3, 0.4, 0.5])
def test_config5_sd21_768_topk_sweep(lib, ratio):
    """BASELINE configs[4]: SD-2.1 768x768 geometry (96x96 latents: 9216 / 2304 / 576 / 144 tokens), top-k sweep.
    S is not a power of two, so counters are compared as integers."""
    pipe, names, n_exp, shapes = moefied_pipeline(96, 1, ratio=ratio)
    assert [s for (_, _, _, s) in shapes][:7] == [9216, 9216, 2304, 2304, 576, 576, 144]
    rec = nr.FrequencyMeasure(0, 1, 16, n_exp, names)
    rec.reset()
    out, _ = rec.observe_activation(pipe, "a photo of a dog")
    counts = rec.int_counts().cpu().numpy()
    for l, (n, d, h, s) in enumerate(shapes):
        E = h // 20
        k = int(E * ratio)                              # helper.py:61, SURVEY appendix A.4
        assert pipe.unet.get_submodule(n).k == k
        assert counts[0, l, :E].sum() == s * k           # batch row 0 only
        lc = rec.label_counter[0][l]
        assert np.array_equal(np.rint(lc * s).astype(np.int64), counts[0, l, :E])
    assert all(torch.isfinite(x.float()).all() for x in out)


def _wanda_like(rs, d, h, ratio):
    m = np.zeros((d, h), dtype=np.int64)
    kk = max(1, int(ratio * h))
    for r in range(d):
        cols = rs.choice(h, kk, replace=False)
        m[r, cols[rs.rand(kk) < 0.5]] = 1
    return m


def test_config5_multi_concept_union_masks(lib, golden_dir, tmp_path):
    """BASELINE configs[4], second half: union of per-concept Wanda weight masks on every ff.net.2 of the
    SD-2.1-768 FFN stack; the 320x1280 layers use the reference's real masks (weights_320_1280.csv)."""
    import scipy.sparse as sp
    g = np.load(os.path.join(golden_dir, "wanda_csv_320_1280.npz"), allow_pickle=False)
    real = np.unpackbits(g["packed"], axis=2, bitorder="little")           # [5, 320, 1280]
    torch.manual_seed(0)
    unet = FFNStackUNet(latent_hw=96)
    pipe = SyntheticFFNPipeline(unet, num_inference_steps=1, device=DEV)
    shapes = sd_ffn_shapes(96)
    rs = np.random.RandomState(3)
    masks = {}
    for ci, c in enumerate(["van_gogh", "monet", "nudity"]):
        p = tmp_path / f"seed_0_{c}" / "skilled_neuron_wanda" / "0.02"
        os.makedirs(p)
        for l, (n, d, h, s) in enumerate(shapes):
            m = real[(ci + l) % 5].astype(np.int64) if d == 320 else _wanda_like(rs, d, h, 0.02)
            masks[(c, l)] = m
            with open(p / f"timestep_0_layer_{l}.pkl", "wb") as f:
                pickle.dump(sp.csr_matrix(m), f)
    mc = nr.MultiConceptRemoverWanda(str(tmp_path) + "/seed_%s_%s", 0, 1, 16, concepts_to_remove=["van_gogh", "monet", "nudity"],
                                     wanda_thr={"van_gogh": 0.02, "monet": 0.02, "nudity": 0.02})
    pre, removed, singles = mc.remove_concepts(pipe, "a painting", ["van_gogh", "monet"])
    assert len(singles) == 2 and len(removed) == 16
    # union bits == OR of the concept masks, bit-exact, for every layer
    for l, (n, d, h, s) in enumerate(shapes):
        want = np.logical_or(masks[("van_gogh", l)], masks[("monet", l)]).astype(np.uint8)
        got = mc.union_neuron_remover.mask_bits(0, l, DEV).cpu().numpy().view(np.uint8)
        assert np.array_equal(got, np.packbits(want.reshape(-1), bitorder="little"))
    # masked runs differ from the unmasked one, and the union differs from each single-concept run
    diff = lambda a, b: max(float((x.float() - y.float()).abs().max()) for x, y in zip(a, b))
    assert diff(pre, removed) > 0 and diff(removed, singles[0]) > 0 and diff(removed, singles[1]) > 0
    assert all(torch.isfinite(x.float()).all() for x in removed)
    # the masked down-projection equals the oracle's F.linear(x, W2 * (1 - M)) on a layer sample
    l = 0
    lin = pipe.unet.get_submodule(shapes[l][0][:-1] + "2")
    x = torch.randn(2, 64, 1280, device=DEV, dtype=torch.bfloat16)
    single = mc.removers["monet"]
    single.reset_time_layer()
    y = single.linear_hook_fn(lin, (x,), None).float().cpu()
    ref = O.wanda_down_proj(x.float().cpu(), lin.weight.float().cpu(), lin.bias.float().cpu(), masks[("monet", 0)])
    assert float((y - ref).norm() / ref.norm()) < 1e-2
