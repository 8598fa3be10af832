"""Where does the mainloop time go?  Time K1 / K3 with the TMA loads, the MMAs or the epilogue
disabled (MOE_DEBUG_MODE bits 1 / 2 / 4; results are garbage in those modes, timing only)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M  # noqa: E402
dev = "cuda:0"
REPS = 10

def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REPS)

os.environ["MOE_K1_CLUSTER"] = "1,1"; os.environ["MOE_K3_CLUSTER"] = "1,1"
for d, T in [(320, 65536), (1280, 4096)]:
    h = 4 * d; es = 20; E = h // es
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = torch.zeros(2 * h, device=dev)
    w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = torch.zeros(d, device=dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
    y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    out = []
    for mode, name in [(0, "full"), (1, "noTMA"), (2, "noMMA"), (4, "noEPI"), (3, "noTMA+noMMA"), (5, "noTMA+noEPI"), (6, "noMMA+noEPI"), (7, "none")]:
        os.environ["MOE_DEBUG_MODE"] = str(mode)
        t1 = timed(lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc))
        t3 = timed(lambda: M.down_proj(H, w2, b2, out=y))
        out.append(f"{name}: K1 {t1:6.1f} K3 {t3:6.1f}")
    print(f"d={d} T={T}: " + " | ".join(out))
    out = []
    for st in (2, 3, 4, 6):
        os.environ["MOE_DEBUG_STAGES"] = str(st)
        for mode, name in [(7, "none"), (15, "none+plainarrive"), (0, "full")]:
            os.environ["MOE_DEBUG_MODE"] = str(mode)
            t3 = timed(lambda: M.down_proj(H, w2, b2, out=y))
            out.append(f"st{st} {name}: K3 {t3:6.1f}")
    os.environ.pop("MOE_DEBUG_STAGES")
    print("   " + " | ".join(out))
