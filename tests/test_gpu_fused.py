"""GPU: the fused layer kernel (moe_ffn_fused: K1 -> in-kernel routing -> K3 in one persistent launch) against
the oracle, against the three separate kernels, and for the cross-CTA hand-over protocol (repeat launches on the
same workspace, ragged / tiny / odd row-block counts, split-K shapes)."""
import numpy as np
import pytest
import torch

import moe_b200 as M
import moe_ffn_oracle as O
from gpu_util import DEV, cuda_layer, fused_layer, oracle_layer, check_layer, rel_err, OUT_REL_TOL

pytestmark = pytest.mark.gpu

FUSED_CASES = [
    # d, h, (B, S), es, ratio, act
    (64, 256, (2, 48), 16, 0.3, O.ACT_GELU),        # one row block, one pair busy
    (64, 256, (1, 1), 16, 0.5, O.ACT_RELU),         # single token
    (64, 256, (3, 171), 16, 1.0, O.ACT_GELU),       # k == E (identity mask), ragged, odd number of row blocks
    (128, 640, (1, 300), 20, 0.3, O.ACT_GELU),      # nv = 80 tiles, 3 row blocks
    (320, 1280, (1, 1024), 20, 0.3, O.ACT_GELU),    # SD-1.5 down_blocks.0 geometry, reference experts
    (320, 1280, (1, 1024), 64, 0.3, O.ACT_RELU),    # BASELINE-literal 20 experts of 64 (tpt = 2)
    (640, 2560, (2, 256), 20, 0.3, O.ACT_GELU),     # split-K down-projection
    (1280, 5120, (2, 64), 20, 0.3, O.ACT_GELU),     # mid-block geometry: E = 256 (tpt = 16), one row block
    (1280, 5120, (2, 256), 20, 0.1, O.ACT_GELU),    # d = 1280 at 512 tokens: 4 chunks per block
]


@pytest.mark.parametrize("d,h,shape,es,ratio,act", FUSED_CASES)
def test_fused_layer_matches_oracle(lib, d, h, shape, es, ratio, act):
    layer = O.synthetic_layer(d, h, shape, es, seed=d + es)
    cu = fused_layer(layer, ratio, act)
    orc = oracle_layer(layer, ratio, act)
    stats = check_layer(cu, orc, min_safe_fraction=0.85)
    S = shape[1]
    assert torch.equal(cu["hist"], torch.bincount(cu["idx"][:S].reshape(-1), minlength=cu["E"]))
    assert int(cu["hist"].sum()) == S * cu["k"]
    # labels ascending and consistent with the expert-set words
    assert (cu["idx"][:, 1:] > cu["idx"][:, :-1]).all() or cu["k"] <= 1
    assert [set(r.tolist()) for r in cu["idx"]] == cu["sets"]
    print(stats)


def test_fused_config1_full_size_matches_unfused_and_oracle(lib):
    """BASELINE configs[0] at full size (d 320, 2 x 4096 tokens, 64 experts of 20, k = 19)."""
    layer = O.synthetic_layer(320, 1280, (2, 4096), 20, seed=1)
    fu = fused_layer(layer, 0.3)
    un = cuda_layer(layer, 0.3)
    orc = oracle_layer(layer, 0.3)
    check_layer(fu, orc)
    # fused vs separate kernels: same scores up to summation order, identical routing away from ties
    assert torch.allclose(fu["scores"], un["scores"], atol=1e-4, rtol=1e-5)
    margin = O.topk_margin(un["scores"], un["k"]).numpy()
    agree = np.array([a == b for a, b in zip(fu["sets"], un["sets"])])
    assert agree[margin > 1e-3].all()
    assert rel_err(fu["y"].reshape(-1, 320)[agree], un["y"].reshape(-1, 320)[agree]) < 2e-3
    assert int(fu["hist"].sum()) == 4096 * 19


@pytest.mark.parametrize("d,h,shape,es", [(640, 2560, (2, 1024), 20), (1280, 5120, (2, 256), 20), (320, 1280, (2, 4096), 64),
                                          (1280, 5120, (2, 64), 64)])
def test_fused_full_size_layers_match_oracle(lib, d, h, shape, es):
    """The other SD-1.5 layer shapes of BASELINE configs[1] at FULL size against the oracle (not against the separate
    kernels): d = 640 / 2048 tokens, d = 1280 / 512 tokens (split-K down-projection), and the BASELINE-literal
    64-neuron experts at 8192 tokens and at the mid-block's 128 tokens."""
    layer = O.synthetic_layer(d, h, shape, es, seed=d + shape[1])
    fu = fused_layer(layer, 0.3)
    orc = oracle_layer(layer, 0.3)
    stats = check_layer(fu, orc, min_safe_fraction=0.85)
    assert int(fu["hist"].sum()) == shape[1] * fu["k"]
    print(stats)


def test_fused_kernel_finishes_next_to_a_competing_kernel(lib):
    """the fused kernel's CTA pairs wait on each other, so a kernel of another stream that holds
    SMs delays the pairs that found none -- it must finish (the stragglers start when the competitor's CTAs retire),
    not trap, and give the same bits as an undisturbed launch.  Two fused launches on two streams are ordered by the
    library (they must never be resident together)."""
    layer = O.synthetic_layer(320, 1280, (2, 2048), 20, seed=9)
    ref = fused_layer(layer, 0.3)
    from moe_b200.packing import ExpertLayout, pack_ffn
    lay = ExpertLayout.from_labels(layer["labels"])
    E, es, k = lay.n_experts, lay.expert_size, ref["k"]
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"], device=DEV)
    xt = layer["x"].reshape(-1, 320).to(DEV, torch.bfloat16).contiguous()
    side, side2 = torch.cuda.Stream(), torch.cuda.Stream()
    big = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        with torch.cuda.stream(side):                      # ~2 ms of cuBLAS work occupying the SMs
            for _ in range(3):
                big @ big
        y, _, _, bits, _ = M.ffn_fused(xt, p.w1p, p.b1p, p.w2p, p.b2, E, es, k, want_bits=True)    # current stream
        with torch.cuda.stream(side2):                     # a second fused launch on another stream
            y2, _, _, bits2, _ = M.ffn_fused(xt, p.w1p, p.b1p, p.w2p, p.b2, E, es, k, want_bits=True)
        outs.append((y, bits, y2, bits2))
    torch.cuda.synchronize()
    for y, bits, y2, bits2 in outs:
        assert torch.equal(bits.cpu(), ref["bits"]) and torch.equal(bits2.cpu(), ref["bits"])
        assert torch.equal(y.float().cpu().view(ref["y"].shape), ref["y"]) and torch.equal(y2.float().cpu().view(ref["y"].shape), ref["y"])


def test_fused_removed_experts_and_count_window(lib):
    """RemoveExperts rule inside the fused kernel: removed experts score exactly 0, still compete, own no neurons
    (remove_skilled_experts.py:29-49); the histogram counts the rows of the given window only."""
    layer = O.synthetic_layer(320, 1280, (2, 512), 20, seed=5)
    removed = [0, 7, 13, 31, 32, 63]
    cu = fused_layer(layer, 0.3, removed=removed, count_rows=(100, 700))
    un = cuda_layer(layer, 0.3, removed=removed, count_rows=(100, 700))
    orc = oracle_layer(layer, 0.3, removed=removed, timestep=0)
    want = [set(r.tolist()) for r in orc["labels"].reshape(-1, orc["labels"].shape[-1])]
    safe = (orc["margin"] > 2e-3).numpy()
    n = len(want)
    got = [set(r.tolist()) for r in cu["idx"]]                    # labels include removed experts that took a slot
    agree = np.array([got[i] == want[i] for i in range(n)])
    assert safe.mean() > 0.85 and agree[safe].all()
    for t in range(n):
        assert cu["sets"][t] == got[t] - set(removed)             # active set = selected and not removed
    Hc, Ho = cu["H"].reshape(n, -1), orc["H"].reshape(n, -1)
    assert rel_err(Hc[agree], Ho[agree]) < OUT_REL_TOL
    assert rel_err(cu["y"].reshape(n, -1)[agree], orc["y"].reshape(n, -1)[agree]) < OUT_REL_TOL
    dead = orc["pat"][removed].sum(0) > 0
    assert torch.all(Hc[:, dead] == 0)                            # removed experts' neurons are always masked
    assert torch.equal(cu["hist"], torch.bincount(cu["idx"][100:700].reshape(-1), minlength=64))
    # same routing as the separate router kernel on (nearly) the same scores
    same = np.array([a == b for a, b in zip(cu["sets"], un["sets"])])
    assert same[O.topk_margin(un["scores"], un["k"]).numpy() > 1e-3].all()


def test_fused_removed_experts_with_k_equal_E(lib):
    """ratio 1.0 (moefy_config.yaml:5) with a removal list: every expert is selected, the removed ones still own no
    neurons (remove_skilled_experts.py:29-35 zeroes their pattern rows) -- found by tools/stress_fused.py."""
    layer = O.synthetic_layer(128, 640, (2, 150), 20, seed=9)
    removed = [3, 4, 17, 31]
    cu = fused_layer(layer, 1.0, removed=removed)
    un = cuda_layer(layer, 1.0, removed=removed)
    orc = oracle_layer(layer, 1.0, removed=removed, timestep=0)
    n = 300
    dead = orc["pat"][removed].sum(0) > 0
    assert torch.all(cu["H"].reshape(n, -1)[:, dead] == 0)
    assert torch.equal(cu["H"], un["H"]) and rel_err(cu["y"], un["y"]) < 2e-3
    assert rel_err(cu["H"], orc["H"]) < OUT_REL_TOL and rel_err(cu["y"], orc["y"]) < OUT_REL_TOL
    assert all(s == set(range(32)) - set(removed) for s in cu["sets"])


def test_fused_repeat_launches_share_workspace(lib):
    """The sync counters are left at zero by every launch: back-to-back launches on one workspace, of different
    geometries, give the same results as fresh ones."""
    a = O.synthetic_layer(320, 1280, (2, 700), 20, seed=3)
    b = O.synthetic_layer(640, 2560, (1, 130), 20, seed=4)
    first_a = fused_layer(a, 0.3)
    first_b = fused_layer(b, 0.3)
    again_a = fused_layer(a, 0.3, repeats=5)
    again_b = fused_layer(b, 0.3, repeats=3)
    for x, y in ((first_a, again_a), (first_b, again_b)):
        assert torch.equal(x["scores"], y["scores"])
        assert torch.equal(x["idx"], y["idx"])
        assert torch.equal(x["H"], y["H"])
        assert torch.equal(x["y"], y["y"])            # deterministic, including the split-K reduction order
        assert torch.equal(x["hist"], y["hist"])
    ws = M.fused_workspace(DEV, 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sync_bytes = 4 * (32 + 2048 * 32)      # header + per-block records
    assert int(ws[:sync_bytes].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("env,val", [("MOE_FUSED_ARES", "1"), ("MOE_FUSED_ARES", "2"), ("MOE_FUSED_DIRECT", "1")])
@pytest.mark.parametrize("d,h,shape,es", [(320, 1280, (2, 2500), 20), (128, 640, (1, 9600), 20), (64, 256, (5, 4000), 16),
                                          (640, 2560, (1, 700), 20)])
def test_fused_experimental_schedules_are_bit_identical(lib, monkeypatch, env, val, d, h, shape, es):
    """The two experimental phase-1 schedules (off by default) give the same bits as the default:
    MOE_FUSED_ARES=1 -- contiguous runs of column tiles per CTA pair with the row block's x panels resident in shared
    memory (=2: whole-tile W1 slots, one barrier round trip per tile); MOE_FUSED_DIRECT=1 -- H rows stored straight from the
    epilogue threads, the staging space used as ring slots.
    Both put phase 1 on its own ring and re-carve shared memory for phase 3; ragged last row block included."""
    layer = O.synthetic_layer(d, h, shape, es, seed=11)
    ref = fused_layer(layer, 0.3)
    monkeypatch.setenv(env, val)
    res = fused_layer(layer, 0.3, repeats=2)
    for key in ("scores", "idx", "H", "y", "hist"):
        assert torch.equal(ref[key], res[key]), key


def test_fused_reversed_phase3_order_is_bit_identical(lib, monkeypatch):
    """When H exceeds 96 MB (UNet batch 16 at d = 320) the phase-3 items walk the row blocks from the last-written one
    down (their H tiles are still in L2); MOE_FUSED_REV3 forces the order either way.  Same arithmetic, same bits --
    checked on a small ragged shape (forced) and on one large enough to switch by itself."""
    for d, h, shape, es in [(320, 1280, (3, 1000), 20), (320, 1280, (1, 40100), 20)]:
        layer = O.synthetic_layer(d, h, shape, es, seed=5)
        monkeypatch.setenv("MOE_FUSED_REV3", "0")
        ref = fused_layer(layer, 0.3)
        monkeypatch.setenv("MOE_FUSED_REV3", "1")
        rev = fused_layer(layer, 0.3, repeats=2)
        monkeypatch.delenv("MOE_FUSED_REV3")
        auto = fused_layer(layer, 0.3)
        for key in ("scores", "idx", "H", "y", "hist"):
            assert torch.equal(ref[key], rev[key]), key
            assert torch.equal(ref[key], auto[key]), key


def test_fused_without_masking_is_the_dense_ffn(lib):
    """mask_h=False (the ExpertPredictivity contract, expert_activation.py:62: statistics only, unmasked output): the
    routing outputs are unchanged, H and Y are those of the dense FFN."""
    for d, h, shape, es in [(320, 1280, (2, 777), 20), (1280, 5120, (2, 64), 20), (320, 1280, (1, 500), 64)]:
        layer = O.synthetic_layer(d, h, shape, es, seed=11)
        a = fused_layer(layer, 0.3, mask_h=True)
        b = fused_layer(layer, 0.3, mask_h=False)
        assert torch.equal(a["idx"], b["idx"]) and torch.equal(a["hist"], b["hist"]) and torch.equal(a["bits"], b["bits"])
        assert torch.equal(a["scores"], b["scores"])
        dense = cuda_layer(layer, 1.0)                            # k == E: identity mask through the separate kernels
        assert rel_err(b["H"], dense["H_unmasked"]) < 1e-6
        assert rel_err(b["y"], dense["y"]) < 2e-3
        assert (a["H"] == 0).float().mean() > 0.6                 # ~70 % of the neurons masked at ratio 0.3


def test_fused_randomised_against_separate_kernels(lib):
    """tools/stress_fused.py: 80 random (geometry, token count, ratio, activation, removal list, count window) cases;
    every case runs the fused kernel three times on the shared workspace (bit-identical repeats) and must agree
    with the K1 -> K2 -> K3 path: identical labels / histograms, routing bit-exact on its own scores, same Y."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_fused.py"), "80", "3"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "80 cases, 0 failures" in res.stdout


def test_fused_empty_input_is_a_no_op(lib):
    """Zero tokens (an empty prompt shard): no launch, empty outputs, counters untouched."""
    from moe_b200.packing import ExpertLayout, pack_ffn
    layer = O.synthetic_layer(64, 256, (1, 4), 16, seed=2)
    lay = ExpertLayout.from_labels(layer["labels"])
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"], device=DEV)
    x = torch.empty(0, 64, dtype=torch.bfloat16, device=DEV)
    hist = torch.full((lay.n_experts,), 7, dtype=torch.int64, device=DEV)
    y, H, scores, bits, idx = M.ffn_fused(x, p.w1p, p.b1p, p.w2p, p.b2, lay.n_experts, lay.expert_size, 4, O.ACT_GELU,
                                          want_bits=True, want_idx=True, hist=hist, count_rows=(0, 0))
    torch.cuda.synchronize()
    assert y.shape == (0, 64) and H.shape == (0, 256) and scores.shape == (0, lay.n_experts)
    assert idx.shape[0] == 0 and bits.shape[0] == 0
    assert bool((hist == 7).all())
    # the separate entry points accept the same empty shard
    H2, sc2, _ = M.geglu_up(x, p.w1p, p.b1p, lay.n_experts, lay.expert_size, O.ACT_GELU)
    b2, i2 = M.router_topk(sc2, 4, want_bits=True, want_idx=True, hist=hist, H=H2, expert_size=lay.expert_size, count_rows=(0, 0))
    y2 = M.down_proj(H2, p.w2p, p.b2)
    torch.cuda.synchronize()
    assert y2.shape == (0, 64) and i2.shape[0] == 0 and bool((hist == 7).all())


def test_fused_k_zero_masks_everything(lib):
    """ratio so small that k = int(E * ratio) = 0 (helper.py:61 allows it): no expert is active, H is all zeros and the
    output is the down-projection bias."""
    layer = O.synthetic_layer(320, 1280, (1, 300), 20, seed=5)
    cu = fused_layer(layer, 0.01)
    assert cu["k"] == 0
    assert float(cu["H"].abs().max()) == 0.0
    assert int(cu["hist"].sum()) == 0 and all(len(s_) == 0 for s_ in cu["sets"])
    b2 = layer["b2"].to(torch.bfloat16).float()
    assert torch.equal(cu["y"].reshape(-1, 320), b2.expand(300, 320))


def test_fused_unsupported_geometry_raises(lib):
    x = torch.zeros(8, 40, dtype=torch.bfloat16, device=DEV)
    w1 = torch.zeros(320, 40, dtype=torch.bfloat16, device=DEV)
    w2 = torch.zeros(40, 160, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(M._lib.MoeLibraryError, match="multiples of 64"):
        M.ffn_fused(x, w1, None, w2, None, 20, 8, 5)
