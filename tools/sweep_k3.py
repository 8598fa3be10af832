"""K3 tuning sweep: tile width x split-K x pairing for the SD-1.5 layer shapes (graph-timed)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
REPS = 10
def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS): fn()
    torch.cuda.current_stream().wait_stream(s); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REPS)
for d, T in [(320, 8192), (640, 2048), (1280, 512), (1280, 128)]:
    h = 4 * d
    gen = torch.Generator().manual_seed(0)
    H = (torch.randn(T, h, generator=gen) * 0.3).to(dev, torch.bfloat16)
    w2 = ((torch.rand(d, h, generator=gen) * 2 - 1) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = torch.zeros(d, device=dev); y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    os.environ["MOE_DEBUG_PRINT"] = "1"
    M.down_proj(H, w2, b2, out=y); torch.cuda.synchronize()
    os.environ.pop("MOE_DEBUG_PRINT")
    print(f"d={d} T={T}: auto {timed(lambda: M.down_proj(H, w2, b2, out=y)):.1f}us  cuBLAS {timed(lambda: torch.matmul(H, w2.t())):.1f}us")
    for pair in ("0", "1"):
        os.environ["MOE_PAIR"] = pair
        for bn in (64, 80, 128, 160, 256):
            if d % bn and bn != 256: continue
            row = []
            for sp in (1, 2, 3, 4, 6, 8):
                os.environ["MOE_K3_BN"] = str(bn); os.environ["MOE_K3_SPLIT"] = str(sp)
                try:
                    row.append(f"S{sp}:{timed(lambda: M.down_proj(H, w2, b2, out=y)):5.1f}")
                except Exception as e:
                    row.append(f"S{sp}:ERR")
            print(f"   pair={pair} bn={bn:3d}  " + "  ".join(row))
    for k in ("MOE_PAIR", "MOE_K3_BN", "MOE_K3_SPLIT"): os.environ.pop(k, None)
