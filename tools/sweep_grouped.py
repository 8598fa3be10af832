"""Grouped / gathered down-projection (router permutation + moe_down_grouped) against the dense-masked form
(router zero-writes + moe_down_proj) on the BASELINE-literal expert geometry (64-neuron experts, top-k 30 %).
CUDA-graph replays over REP distinct buffer sets (working set > L2), CUDA events; microseconds per layer call."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
ES = 64


def timed(fn, rep, iters=20):
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
    return e0.elapsed_time(e1) / iters / rep * 1e3


shapes = [(320, 4096), (320, 8192), (320, 65536), (640, 2048), (640, 16384), (1280, 512), (1280, 4096)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
print("d T | dense: router(mask) + down_proj | grouped: router(bits) + permutation + grouped GEMM + combine   [us per layer call]")
for d, T in shapes:
    h = 4 * d; E = h // ES; k = int(E * 0.3)
    rep = 8 if T <= 8192 else 2
    gen = torch.Generator().manual_seed(0)
    sets = []
    for r in range(rep):
        x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
        w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
        w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
        b2 = torch.zeros(d, device=dev)
        H, sc, _ = M.geglu_up(x, w1, torch.zeros(2 * h, device=dev), E, ES)
        bits, _ = M.router_topk(sc, k)
        sets.append(dict(H=H, Hm=H.clone(), sc=sc, w2=w2, b2=b2, bits=bits, perm=M.expert_permutation(bits, E, k),
                         y=torch.empty(T, d, dtype=torch.bfloat16, device=dev), bits_out=torch.empty_like(bits)))
    torch.cuda.synchronize()

    def dense():
        for s in sets:
            M.router_topk(s["sc"], k, want_bits=False, H=s["Hm"], expert_size=ES)
            M.down_proj(s["Hm"], s["w2"], s["b2"], out=s["y"])

    def dense_gemm():
        for s in sets:
            M.down_proj(s["Hm"], s["w2"], s["b2"], out=s["y"])

    def grouped():
        for s in sets:
            M.router_topk(s["sc"], k, bits_out=s["bits_out"])
            p = M.expert_permutation(s["bits_out"], E, k)
            M.down_grouped(s["H"], p, s["w2"], s["b2"], E, ES, out=s["y"])

    def grouped_gemm():
        for s in sets:
            M.down_grouped(s["H"], s["perm"], s["w2"], s["b2"], E, ES, out=s["y"])

    def perm_only():
        for s in sets:
            M.expert_permutation(s["bits"], E, k)

    td, tdg, tg, tgg, tp = timed(dense, rep), timed(dense_gemm, rep), timed(grouped, rep), timed(grouped_gemm, rep), timed(perm_only, rep)
    fl_dense, fl_need = 2.0 * d * h * T, 2.0 * d * ES * k * T
    print(f"d={d} T={T} E={E} k={k}: dense {td:7.1f} (GEMM alone {tdg:7.1f} = {fl_dense/tdg/1e6:5.0f} TF dense-equivalent) | "
          f"grouped {tg:7.1f} (permutation {tp:6.1f}, GEMM + combine {tgg:7.1f} = {fl_need/tgg/1e6:5.0f} TF necessary)", flush=True)
