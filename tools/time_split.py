"""Per-kernel time of the K1 -> K2 -> K3 path per layer shape (CUDA-graph replays over 8 distinct buffer sets).
MOE_ROUTER_LEGACY=1 forces the warp-per-token router."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import moe_b200 as M
dev = "cuda:0"
ES, REP = 20, 8


def timed(fn, iters=20):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / REP * 1e3


for d, T in [(320, 8192), (640, 2048), (1280, 512), (1280, 128)]:
    h = 4 * d; E = h // ES; k = int(E * 0.3)
    gen = torch.Generator().manual_seed(0)
    sets = []
    for r in range(REP):
        sets.append(dict(x=torch.randn(T, d, generator=gen).to(dev, torch.bfloat16),
                         w1=(torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16), b1=torch.zeros(2 * h, device=dev),
                         w2=(torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16), b2=torch.zeros(d, device=dev),
                         H=torch.empty(T, h, dtype=torch.bfloat16, device=dev), sc=torch.randn(T, E, device=dev),
                         y=torch.empty(T, d, dtype=torch.bfloat16, device=dev), hist=torch.zeros(E, dtype=torch.int64, device=dev)))
    k1 = timed(lambda: [M.geglu_up(s["x"], s["w1"], s["b1"], E, ES, out=s["H"], scores_out=s["sc"]) for s in sets])
    k2 = timed(lambda: [M.router_topk(s["sc"], k, want_bits=False, hist=s["hist"], H=s["H"], expert_size=ES, count_rows=(0, T // 2)) for s in sets])
    k3 = timed(lambda: [M.down_proj(s["H"], s["w2"], s["b2"], out=s["y"]) for s in sets])
    print(f"d={d} T={T}: K1 {k1:6.1f} us | K2 {k2:6.1f} us | K3 {k3:6.1f} us | sum {k1 + k2 + k3:6.1f}", flush=True)
