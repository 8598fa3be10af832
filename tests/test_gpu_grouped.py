"""GPU: the north star's remaining router / down-projection forms:
  * the several-tokens-per-warp router against the warp-per-token one and torch.topk,
  * the compacted token -> expert permutation (moe_expert_permutation) against numpy,
  * the grouped / gathered down-projection (moe_down_grouped) against the oracle and the dense-masked K3."""
import numpy as np
import pytest
import torch

import moe_b200 as M
import moe_ffn_oracle as O
from moe_b200 import _lib
from moe_b200.packing import ExpertLayout, pack_ffn, bits_to_sets
from gpu_util import DEV, rel_err, r16, OUT_REL_TOL, MARGIN_TOL

pytestmark = pytest.mark.gpu


def _route(scores, k, E, es, removed=None, bias=None, mask=True):
    H = torch.ones(scores.shape[0], E * es, dtype=torch.bfloat16, device=DEV) if mask else None
    hist = torch.zeros(E, dtype=torch.int64, device=DEV)
    cmax = torch.full((E,), float("-inf"), device=DEV)
    rb = None if not removed else M.bits_from_expert_list(removed, E).to(DEV)
    bits, idx = M.router_topk(scores, k, removed_bits=rb, want_idx=True, hist=hist, colmax_out=cmax, H=H, expert_size=es,
                              count_rows=(3, scores.shape[0] - 2), score_bias=bias)
    torch.cuda.synchronize()
    return bits.cpu(), idx.cpu(), hist.cpu(), cmax.cpu(), None if H is None else H.cpu()


@pytest.mark.parametrize("Tn,E,es,k", [(1000, 64, 20, 19), (777, 20, 64, 6), (513, 128, 20, 38), (300, 256, 20, 76),
                                        (90, 256, 4, 255), (64, 8, 16, 3), (200, 40, 8, 12), (33, 96, 4, 1)])
def test_multi_token_router_equals_warp_per_token_router(lib, monkeypatch, Tn, E, es, k):
    g = torch.Generator().manual_seed(E + k)
    scores = torch.randn(Tn, E, generator=g)
    scores[::4] = torch.round(scores[::4] * 2) / 2          # many exact ties
    scores[1::7] = scores[1::7].abs() * 3 + 1               # shared sign / exponent prefixes
    dsc = scores.to(DEV)
    removed = sorted(set(int(v) for v in torch.randint(0, E, (max(1, E // 10),), generator=g)))
    bias = (torch.rand(E, generator=g) * (torch.rand(E, generator=g) < 0.2)).to(DEV)
    for rm, bs in ((None, None), (removed, None), (None, bias), (removed, bias)):
        monkeypatch.setenv("MOE_ROUTER_LEGACY", "1")
        ref = _route(dsc, k, E, es, rm, bs)
        monkeypatch.setenv("MOE_ROUTER_LEGACY", "0")
        new = _route(dsc, k, E, es, rm, bs)
        for a, b, what in zip(ref, new, ("bits", "idx", "hist", "colmax", "H")):
            assert torch.equal(a, b), (what, rm is not None, bs is not None)
    # and against torch.topk on the plain scores: ascending ids, ties to the lowest id
    new = _route(dsc, k, E, es, mask=False)
    want = torch.topk(scores, k, dim=-1)[1].sort(dim=-1)[0]
    srt = torch.sort(scores, dim=-1, descending=True)[0]
    clear = (srt[:, k - 1] > srt[:, k]) if k < E else torch.ones(Tn, dtype=torch.bool)
    assert torch.equal(new[1].long()[clear], want[clear])
    assert torch.equal(new[3], scores.max(0)[0])


def test_router_above_256_experts_uses_the_warp_per_token_kernel(lib):
    E, k = 320, 50
    scores = torch.randn(70, E, generator=torch.Generator().manual_seed(1))
    bits, idx = M.router_topk(scores.to(DEV), k, want_idx=True)
    assert torch.equal(idx.cpu().long(), torch.topk(scores, k, dim=-1)[1].sort(dim=-1)[0])


def _numpy_permutation(sets, E, k, pad):
    lists = [[t for t, s in enumerate(sets) if e in s] for e in range(E)]
    offsets = [0]
    for lst in lists:
        offsets.append(offsets[-1] + (len(lst) + pad - 1) // pad * pad)
    return lists, offsets


@pytest.mark.parametrize("Tn,E,k,pad", [(1, 8, 2, 128), (300, 20, 6, 128), (5000, 64, 19, 128), (9000, 256, 76, 128),
                                         (4097, 96, 3, 128), (2500, 40, 40, 32), (130, 33, 0, 128)])
def test_expert_permutation_matches_numpy(lib, Tn, E, k, pad):
    g = torch.Generator().manual_seed(Tn + E)
    scores = torch.randn(Tn, E, generator=g).to(DEV)
    removed = [1, E - 1] if E > 8 else None
    rb = None if removed is None else M.bits_from_expert_list(removed, E).to(DEV)
    bits, _ = M.router_topk(scores, k, removed_bits=rb)
    perm = M.expert_permutation(bits, E, k, row_pad=pad)
    torch.cuda.synchronize()
    sets = bits_to_sets(bits, E)
    lists, offsets = _numpy_permutation(sets, E, k, pad)
    assert perm.offsets.cpu().tolist() == offsets
    assert perm.counts.cpu().tolist() == [len(l) for l in lists]
    tokens = perm.tokens.cpu().numpy()
    for e in range(E):
        seg = tokens[offsets[e]:offsets[e + 1]]
        assert seg[:len(lists[e])].tolist() == lists[e], e           # ascending token order, deterministic
        assert (seg[len(lists[e]):] == -1).all()
    sp = perm.slot_pos.cpu().numpy()
    for t in range(0, Tn, max(1, Tn // 200)):
        act = sorted(sets[t])
        for j, e in enumerate(act):
            assert tokens[sp[t, j]] == t and offsets[e] <= sp[t, j] < offsets[e] + len(lists[e])
        assert (sp[t, len(act):] == -1).all()
    if removed is not None:
        assert all(len(lists[e]) == 0 for e in removed)        # removed experts own no tokens


@pytest.mark.parametrize("d,h,shape,ratio,removed", [(64, 256, (2, 96), 0.5, None), (128, 512, (1, 300), 0.3, [1, 5]),
                                                       (320, 1280, (2, 1000), 0.3, None), (320, 1280, (1, 4096), 0.3, [0, 7, 13]),
                                                       (640, 2560, (2, 333), 0.3, None)])
def test_grouped_down_projection_matches_dense_and_oracle(lib, d, h, shape, ratio, removed):
    """es = 64 (BASELINE-literal geometry): router -> permutation -> grouped down-projection on the UNMASKED H equals
    the dense-masked K3 on the masked H (same bf16 inputs, fp32 accumulation) and the oracle within 1e-2."""
    es = 64
    layer = O.synthetic_layer(d, h, shape, es, seed=d + shape[1])
    lay = ExpertLayout.from_labels(layer["labels"])
    E = lay.n_experts
    k = O.topk_from_ratio(E, ratio)
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"], device=DEV)
    xt = layer["x"].reshape(-1, d).to(DEV, torch.bfloat16).contiguous()
    H, scores, _ = M.geglu_up(xt, p.w1p, p.b1p, E, es)
    rb = None if not removed else M.bits_from_expert_list(removed, E).to(DEV)
    bits, _ = M.router_topk(scores, k, removed_bits=rb)
    perm = M.expert_permutation(bits, E, k)
    y_grouped = M.down_grouped(H, perm, p.w2p, p.b2, E, es)
    Hm = H.clone()
    M.router_topk(scores, k, removed_bits=rb, want_bits=False, H=Hm, expert_size=es)
    y_dense = M.down_proj(Hm, p.w2p, p.b2)
    torch.cuda.synchronize()
    assert rel_err(y_grouped.float(), y_dense.float()) < 4e-3           # both bf16-rounded from fp32 sums
    pat = O.patterns_from_labels(layer["labels"])
    x16, w1, w2 = r16(layer["x"]), r16(layer["w1"]), r16(layer["w2"])
    if removed:
        Ho, _, _, sc = O.remove_experts_forward(x16, w1, layer["b1"], pat, k, removed, 0)
    else:
        Ho, _, _, sc = O.moefy_forward(x16, w1, layer["b1"], pat, k)
    safe = (O.topk_margin(sc, k) > MARGIN_TOL).numpy()
    yo = O.down_proj(r16(Ho), w2, layer["b2"]).reshape(len(safe), -1)
    assert safe.mean() > 0.85
    assert rel_err(y_grouped.float().cpu()[safe], yo[safe]) < OUT_REL_TOL


def test_grouped_down_projection_rejects_other_expert_sizes(lib):
    H = torch.zeros(128, 320, dtype=torch.bfloat16, device=DEV)
    bits = torch.zeros(128, 1, dtype=torch.int32, device=DEV)
    perm = M.expert_permutation(bits, 16, 4)
    w2 = torch.zeros(64, 320, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.MoeLibraryError, match="code -2"):
        M.down_grouped(H, perm, w2, None, 16, 20)
    # empty token shard: both entry points are no-ops
    e = M.expert_permutation(torch.zeros(0, 1, dtype=torch.int32, device=DEV), 16, 4)
    assert e.offsets.cpu().tolist() == [0] * 17
    y = M.down_grouped(torch.zeros(0, 1024, dtype=torch.bfloat16, device=DEV), e,
                       torch.zeros(64, 1024, dtype=torch.bfloat16, device=DEV), None, 16, 64)
    assert y.shape == (0, 64)


@pytest.mark.parametrize("Tn,h,d,density", [(96, 128, 32, 0.05), (100, 256, 64, 0.5), (256, 1280, 320, 0.025), (8192, 1280, 320, 0.116),
                                             (2048, 2560, 640, 0.03), (512, 5120, 1280, 0.05), (128, 5120, 1280, 1.0), (33, 192, 48, 0.0)])
def test_masked_down_projection_in_one_launch(lib, Tn, h, d, density):
    """moe_down_proj_masked: the bit mask is applied to the W2 tiles in shared memory -- result
    bit-identical to moe_mask_weights + moe_down_proj, one launch, and within 1e-2 of the fp32 reference
    F.linear(x, W2 * (1 - M), b2) (remove_wanda_neurons_fast.py:72-77)."""
    g = torch.Generator().manual_seed(Tn + h)
    H = torch.randn(Tn, h, generator=g).to(DEV, torch.bfloat16)
    w2 = (torch.randn(d, h, generator=g) / h ** 0.5).to(DEV, torch.bfloat16)
    b2 = torch.randn(d, generator=g).to(DEV)
    dense = (torch.rand(d, h, generator=g) < density).to(torch.uint8)
    dense[0, :] = 1 if density > 0 else 0                      # a fully masked output row
    bits = M.mask_pack(dense.to(DEV).contiguous())
    want = M.down_proj(H, M.mask_weights(w2, bits), b2)
    M.reset_launch_count()
    got = M.down_proj(H, w2, b2, mask_bits=bits)
    torch.cuda.synchronize()
    assert M.launch_count() == 1
    assert torch.equal(got, want)
    ref = torch.nn.functional.linear(H.float().cpu(), w2.float().cpu() * (1 - dense.float()), b2.cpu())
    assert rel_err(got.float().cpu(), ref) < OUT_REL_TOL
    if density > 0:
        assert torch.equal(got[:, 0].float().cpu(), b2[0].cpu().bfloat16().float().expand(Tn))     # only the bias is left


def test_masked_down_projection_needs_h_multiple_of_64(lib):
    H = torch.zeros(8, 96, dtype=torch.bfloat16, device=DEV)
    w2 = torch.zeros(16, 96, dtype=torch.bfloat16, device=DEV)
    bits = torch.zeros(16 * 96 // 32, dtype=torch.int32, device=DEV)
    with pytest.raises(_lib.MoeLibraryError, match="code -2"):
        M.down_proj(H, w2, None, mask_bits=bits)


@pytest.mark.parametrize("dtype,n", [(torch.float32, 2 * 4 * 64 * 64), (torch.bfloat16, 16 * 4 * 64 * 64), (torch.float32, 1003),
                                      (torch.bfloat16, 7)])
def test_cfg_ddim_step_matches_torch(lib, dtype, n):
    """CFG combine + DDIM update in one kernel against the formula in fp32 torch."""
    g = torch.Generator().manual_seed(n)
    eu, ec, x = (torch.randn(n, generator=g).to(DEV, dtype) for _ in range(3))
    guidance, a_t, a_prev = 7.5, 0.37, 0.52
    got = M.cfg_ddim_step(eu, ec, x, guidance, a_t, a_prev)
    fe, fc, fx = eu.float(), ec.float(), x.float()
    eps = fe + guidance * (fc - fe)
    x0 = (fx - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    want = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert got.dtype == dtype and rel_err(got.float(), want) < tol
    M.cfg_ddim_step(eu, ec, x, guidance, a_t, a_prev, out=x)           # in place
    assert torch.equal(x, got)
