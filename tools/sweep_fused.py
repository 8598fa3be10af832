"""Per-layer-shape time of the fused layer kernel vs the K1 -> K2 -> K3 triple (CUDA-graph replays, CUDA events)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
ES = int(os.environ.get("ES", "20"))
REP = 8      # distinct buffers per graph so that weights / activations are not L2-resident across launches


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
    return e0.elapsed_time(e1) / iters / REP * 1e3   # us per layer call


shapes = [(320, 8192), (640, 2048), (1280, 512), (1280, 128)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for d, T in shapes:
    h = 4 * d; E = h // ES; k = int(E * 0.3)
    gen = torch.Generator().manual_seed(0)
    sets = []
    for r in range(REP):
        x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
        w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
        b1 = torch.zeros(2 * h, device=dev)
        w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
        b2 = torch.zeros(d, device=dev)
        sets.append(dict(x=x, w1=w1, b1=b1, w2=w2, b2=b2, H=torch.empty(T, h, dtype=torch.bfloat16, device=dev),
                         sc=torch.empty(T, E, device=dev), y=torch.empty(T, d, dtype=torch.bfloat16, device=dev),
                         hist=torch.zeros(E, dtype=torch.int64, device=dev)))

    def fused():
        for s in sets:
            M.ffn_fused(s["x"], s["w1"], s["b1"], s["w2"], s["b2"], E, ES, k, hist=s["hist"], count_rows=(0, T // 2),
                        H_out=s["H"], scores_out=s["sc"], out=s["y"])

    def split():
        for s in sets:
            M.geglu_up(s["x"], s["w1"], s["b1"], E, ES, out=s["H"], scores_out=s["sc"])
            M.router_topk(s["sc"], k, want_bits=False, hist=s["hist"], H=s["H"], expert_size=ES, count_rows=(0, T // 2))
            M.down_proj(s["H"], s["w2"], s["b2"], out=s["y"])

    def cublas():
        for s in sets:
            torch.matmul(s["x"], s["w1"].t()); torch.matmul(s["H"], s["w2"].t())

    fl = 6.0 * d * h * T
    if os.environ.get("FUSED_ONLY"):
        tf = timed(fused)
        print(f"d={d} T={T} es={ES}: fused {tf:6.1f}us ({fl/tf/1e6:5.0f} TF)", flush=True)
        continue
    tf, ts, tc_ = timed(fused), timed(split), timed(cublas)
    print(f"d={d} T={T} es={ES}: fused {tf:6.1f}us ({fl/tf/1e6:5.0f} TF) | split {ts:6.1f}us ({fl/ts/1e6:5.0f} TF) | "
          f"cuBLAS GEMMs only {tc_:6.1f}us ({fl/tc_/1e6:5.0f} TF)", flush=True)
