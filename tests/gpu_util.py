"""Helpers shared by the -m gpu parity tests: run the CUDA path through the C ABI and the oracle
on the SAME (bf16-rounded) inputs."""
import numpy as np
import torch

import moe_b200 as M
import moe_ffn_oracle as O
from moe_b200.packing import ExpertLayout, pack_ffn, bits_to_sets

DEV = "cuda:0"
SCORE_ATOL = 5e-4        # fp32 accumulate order + erff vs torch-CPU erf, sums of <=64 activations
MARGIN_TOL = 2e-3        # tokens whose router margin exceeds this must select identical expert sets
OUT_REL_TOL = 1e-2       # FFN outputs within 1e-2 relative (bf16 vs fp32 reference)


def r16(t):
    """fp32 tensor rounded through bf16 (what the kernels see)."""
    return t.to(torch.bfloat16).float()


def rel_err(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).norm() / (ref.norm() + 1e-30))


def cuda_layer(layer, ratio, act=O.ACT_GELU, removed=None, flags=None, want_gate=False, count_rows=None):
    """K1 -> K2 -> K3 on the GPU through the C ABI.  Returns a dict of CPU tensors, with the
    inner dimension mapped back to the ORIGINAL neuron order."""
    lay = ExpertLayout.from_labels(layer["labels"])
    E, es = lay.n_experts, lay.expert_size
    k = O.topk_from_ratio(E, ratio)
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"], device=DEV)
    x = layer["x"]
    lead = x.shape[:-1]
    xt = x.reshape(-1, x.shape[-1]).to(DEV, torch.bfloat16).contiguous()
    T = xt.shape[0]
    ovr = None
    if flags is not None and len(flags) > 0:
        ovr = torch.from_numpy((np.asarray(flags) == 1)[lay.perm.numpy()].astype(np.uint8)).to(DEV)
    H, scores, gate = M.geglu_up(xt, p.w1p, p.b1p, E, es, act, neuron_override=ovr, want_gate=want_gate)
    H_unmasked = H.clone()
    rb = None if not removed else M.bits_from_expert_list(removed, E).to(DEV)
    hist = torch.zeros(E, dtype=torch.int64, device=DEV)
    cmax = torch.full((E,), float("-inf"), device=DEV)
    rows = count_rows if count_rows is not None else (0, x.shape[-2])
    bits, idx = M.router_topk(scores, k, removed_bits=rb, want_idx=True, hist=hist, colmax_out=cmax, H=H,
                              expert_size=es, count_rows=rows)
    y = M.down_proj(H, p.w2p, p.b2)
    torch.cuda.synchronize()
    inv = lay.inv_perm
    return dict(H=H.float().cpu()[:, inv].view(*lead, -1), H_unmasked=H_unmasked.float().cpu()[:, inv].view(*lead, -1),
                gate=None if gate is None else gate.float().cpu()[:, inv].view(*lead, -1),
                scores=scores.cpu(), idx=idx.cpu().long(), bits=bits.cpu(), hist=hist.cpu(), colmax=cmax.cpu(),
                y=y.float().cpu().view(*lead, -1), k=k, E=E, es=es, sets=bits_to_sets(bits, E))


def fused_layer(layer, ratio, act=O.ACT_GELU, removed=None, count_rows=None, repeats=1, mask_h=True):
    """The same layer through ONE moe_ffn_fused launch (K1 -> routing -> K3 in a persistent kernel)."""
    lay = ExpertLayout.from_labels(layer["labels"])
    E, es = lay.n_experts, lay.expert_size
    k = O.topk_from_ratio(E, ratio)
    p = pack_ffn(lay, layer["w1"], layer["b1"], layer["w2"], layer["b2"], device=DEV)
    x = layer["x"]
    lead = x.shape[:-1]
    xt = x.reshape(-1, x.shape[-1]).to(DEV, torch.bfloat16).contiguous()
    rb = None if not removed else M.bits_from_expert_list(removed, E).to(DEV)
    rows = count_rows if count_rows is not None else (0, x.shape[-2])
    for _ in range(repeats):
        hist = torch.zeros(E, dtype=torch.int64, device=DEV)
        y, H, scores, bits, idx = M.ffn_fused(xt, p.w1p, p.b1p, p.w2p, p.b2, E, es, k, act, removed_bits=rb,
                                              want_bits=True, want_idx=True, hist=hist, count_rows=rows, mask_h=mask_h)
    torch.cuda.synchronize()
    inv = lay.inv_perm
    return dict(H=H.float().cpu()[:, inv].view(*lead, -1), scores=scores.cpu(), idx=idx.cpu().long(), bits=bits.cpu(),
                hist=hist.cpu(), y=y.float().cpu().view(*lead, -1), k=k, E=E, es=es, sets=bits_to_sets(bits, E),
                colmax=scores.max(0)[0].cpu())


def oracle_layer(layer, ratio, act=O.ACT_GELU, removed=None, timestep=0):
    """The oracle on the bf16-rounded weights / inputs (fp32 arithmetic)."""
    pat = O.patterns_from_labels(layer["labels"])
    k = O.topk_from_ratio(pat.shape[0], ratio)
    x, w1, w2 = r16(layer["x"]), r16(layer["w1"]), r16(layer["w2"])
    if removed:
        H, labels, gate, score = O.remove_experts_forward(x, w1, layer["b1"], pat, k, removed, timestep, act)
    else:
        H, labels, gate, score = O.moefy_forward(x, w1, layer["b1"], pat, k, act)
    y = O.down_proj(r16(H), w2, layer["b2"])
    return dict(H=H, labels=labels, gate=gate, score=score, y=y, k=k, E=pat.shape[0],
                margin=O.topk_margin(score, k), pat=pat)


def label_sets(labels):
    return [set(r.tolist()) for r in labels.reshape(-1, labels.shape[-1])]


def check_layer(cu, orc, min_safe_fraction=0.9):
    """The parity bar: scores close; identical expert sets wherever the margin allows; masked H and
    y within 1e-2 relative on the agreeing tokens and overall."""
    assert torch.allclose(cu["scores"], orc["score"], atol=SCORE_ATOL, rtol=1e-5), \
        float((cu["scores"] - orc["score"]).abs().max())
    want = label_sets(orc["labels"])
    safe = (orc["margin"] > MARGIN_TOL).numpy()
    assert safe.mean() >= min_safe_fraction, safe.mean()
    agree = np.array([cu["sets"][t] == want[t] for t in range(len(want))])
    assert agree[safe].all(), f"{(~agree[safe]).sum()} safe tokens disagree"
    # tokens UNDER the margin are not skipped: where the sets differ, only experts whose oracle score lies within
    # 2 * SCORE_ATOL of the oracle's k-th largest score may be involved (a near-tie resolved the other way), and the
    # kernel still selects exactly k experts
    k = cu["k"]
    if 0 < k < cu["E"]:
        kth = torch.topk(orc["score"], k, dim=-1)[0][:, -1]
        for t in np.nonzero(~agree)[0]:
            diff = cu["sets"][t] ^ want[t]
            assert len(cu["sets"][t]) == k, (t, len(cu["sets"][t]))
            for e in diff:
                assert abs(float(orc["score"][t, e]) - float(kth[t])) <= 2 * SCORE_ATOL + 1e-5 * abs(float(kth[t])), \
                    (t, e, float(orc["score"][t, e]), float(kth[t]))
    # router in isolation: the oracle's top-k on the CUDA kernel's own scores must match bit-exactly
    # wherever the margin exceeds 1e-6
    iso_margin = O.topk_margin(cu["scores"], cu["k"]).numpy()
    iso = label_sets(O.route_topk(cu["scores"], cu["k"])) if cu["k"] < cu["E"] else [set(range(cu["E"]))] * len(want)
    for t in np.nonzero(iso_margin > 1e-6)[0]:
        assert cu["sets"][t] == iso[t], t
    Hc = cu["H"].reshape(len(want), -1)
    Ho = orc["H"].reshape(len(want), -1)
    assert rel_err(Hc[agree], Ho[agree]) < OUT_REL_TOL
    yc, yo = cu["y"].reshape(len(want), -1), orc["y"].reshape(len(want), -1)
    assert rel_err(yc[agree], yo[agree]) < OUT_REL_TOL
    return dict(safe=float(safe.mean()), agree=float(agree.mean()), rel_H=rel_err(Hc[agree], Ho[agree]),
                rel_y=rel_err(yc[agree], yo[agree]), rel_y_all=rel_err(yc, yo))
