"""K1 (geglu_up, cta_group::2 pairs) with pipeline stages switched off: MOE_DEBUG_MODE bit 1 = TMA loads, 2 = MMAs,
4 = epilogue TMEM read + math (results are garbage in those modes, timing only) -- profiles/r02_phase1_cadence.log."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"; REPS = 6
def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS): fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REPS)
for d, T in [(320, 65536), (640, 16384)]:
    h = 4 * d; es = 20; E = h // es
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = torch.zeros(2 * h, device=dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
    out = []
    for mode, name in [(0, "full"), (1, "noTMA"), (2, "noMMA"), (4, "noEPImath"), (5, "noTMA+noEPI"), (6, "noMMA+noEPI"), (3, "noTMA+noMMA"), (7, "none")]:
        os.environ["MOE_DEBUG_MODE"] = str(mode)
        out.append(f"{name} {timed(lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc)):6.1f}")
    tiles = (T // 256) * (h // 80) / 74
    print(f"K1 d={d} T={T} ({tiles:.1f} tiles per pair): " + " | ".join(out), flush=True)
