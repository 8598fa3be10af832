"""Randomised stress of the fused layer kernel against the three separate kernels (same library, same inputs):
python tools/stress_fused.py [n_cases] [seed].  Every case runs the fused kernel several times back to back on the
shared workspace and compares labels / histograms / outputs with the K1 -> K2 -> K3 path."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import moe_b200 as M
import moe_ffn_oracle as O
dev = "cuda:0"
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
geoms = [(64, 256, 16), (128, 512, 16), (128, 640, 20), (320, 1280, 20), (320, 1280, 64), (640, 2560, 20), (1280, 5120, 20),
         (192, 768, 32), (64, 256, 4), (256, 1024, 8)]
bad = 0
t_start = time.time()
for case in range(n_cases):
    d, h, es = geoms[rs.randint(len(geoms))]
    T = int(rs.choice([1, 2, 31, 64, 127, 128, 129, 255, 256, 257, 300, 511, 512, 640, 1000, 1025, 2048, 3000, 4096, 8192]))
    if d >= 640 and T > 2048:
        T = 2048
    E = h // es
    ratio = float(rs.choice([0.1, 0.3, 0.5, 0.9, 1.0]))
    k = int(E * ratio)
    act = int(rs.randint(2))
    gen = torch.Generator().manual_seed(case)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = (torch.randn(2 * h, generator=gen) * 0.05).to(dev)
    w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = (torch.randn(d, generator=gen) * 0.05).to(dev)
    removed = None
    if rs.rand() < 0.3 and E >= 8:
        removed = M.bits_from_expert_list(sorted(rs.choice(E, max(1, E // 10), replace=False).tolist()), E).to(dev)
    lo = int(rs.randint(0, T)); hi = int(rs.randint(lo, T + 1))
    try:
        outs = []
        for rep in range(3):
            hist = torch.zeros(E, dtype=torch.int64, device=dev)
            y, H, sc, bits, idx = M.ffn_fused(x, w1, b1, w2, b2, E, es, k, act, removed_bits=removed, want_bits=True,
                                              want_idx=True, hist=hist, count_rows=(lo, hi))
            outs.append((y.clone(), H.clone(), sc.clone(), bits.clone(), idx.clone(), hist.clone()))
        torch.cuda.synchronize()
    except M._lib.MoeLibraryError as e:
        if "code -2" in str(e):
            print(f"case {case}: d={d} h={h} es={es} T={T} k={k}: unsupported by the fused kernel ({str(e)[-60:]})")
            continue
        raise
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            if not torch.equal(a, b):
                bad += 1; print(f"case {case}: NOT DETERMINISTIC across repeats"); break
    # separate kernels
    H2, sc2, _ = M.geglu_up(x, w1, b1, E, es, act)
    hist2 = torch.zeros(E, dtype=torch.int64, device=dev)
    bits2, idx2 = M.router_topk(sc2, k, removed_bits=removed, want_idx=True, hist=hist2, H=H2, expert_size=es, count_rows=(lo, hi))
    y2 = M.down_proj(H2, w2, b2)
    torch.cuda.synchronize()
    y, H, sc, bits, idx, hist = outs[0]
    ok_scores = torch.allclose(sc, sc2, atol=2e-4, rtol=1e-5)
    own = O.route_topk(sc.cpu(), k).sort(-1)[0] if 0 < k < E else None          # router in isolation, own scores
    margin = O.topk_margin(sc.cpu() if removed is None else sc.cpu(), k) if 0 < k < E else None
    ok_route = True
    if own is not None and removed is None:
        safe = margin > 1e-6
        ok_route = bool((idx.cpu().long()[safe] == own[safe]).all())
    same = (idx == idx2).all(-1) if k > 0 else torch.ones(T, dtype=torch.bool, device=dev)
    ok_hist = bool(hist.sum() == (hi - lo) * k)
    rel = float(((y[same].float() - y2[same].float()).norm() / (y2[same].float().norm() + 1e-20))) if same.any() else 0.0
    ok = ok_scores and ok_route and ok_hist and rel < 5e-3 and float(same.float().mean()) > 0.9
    if not ok:
        bad += 1
    print(f"case {case}: d={d} h={h} es={es} E={E} k={k} T={T} act={act} removed={removed is not None} rows=[{lo},{hi}) "
          f"scores={ok_scores} route={ok_route} hist={ok_hist} same_tokens={float(same.float().mean()):.4f} rel_y={rel:.2e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
print(f"{n_cases} cases, {bad} failures, {time.time() - t_start:.1f}s")
sys.exit(1 if bad else 0)
