"""Wanda weight-mask removal on the down-projection (remove_wanda_neurons_fast.py:69-83): the one-launch form
(moe_down_proj_masked: bit mask applied to the W2 tiles in shared memory) against round 1's two launches
(moe_mask_weights writes a masked copy, moe_down_proj consumes it) and the unmasked K3.  us per layer call."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
REP = 8


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


print("d T density | plain K3 | mask_weights + K3 (2 launches) | masked K3 (1 launch)   [us per layer call]")
for d, T in [(320, 8192), (320, 32768), (640, 2048), (640, 8192), (1280, 512), (1280, 2048), (1280, 128)]:
    h = 4 * d
    for density in (0.025, 0.116):
        gen = torch.Generator().manual_seed(0)
        sets = []
        for r in range(REP):
            H = torch.randn(T, h, generator=gen).to(dev, torch.bfloat16)
            w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
            bits = M.mask_pack((torch.rand(d, h, generator=gen) < density).to(torch.uint8).to(dev))
            sets.append(dict(H=H, w2=w2, b2=torch.zeros(d, device=dev), bits=bits, w2m=torch.empty_like(w2),
                             y=torch.empty(T, d, dtype=torch.bfloat16, device=dev)))

        def plain():
            for s in sets:
                M.down_proj(s["H"], s["w2"], s["b2"], out=s["y"])

        def two():
            for s in sets:
                M.mask_weights(s["w2"], s["bits"], out=s["w2m"])
                M.down_proj(s["H"], s["w2m"], s["b2"], out=s["y"])

        def one():
            for s in sets:
                M.down_proj(s["H"], s["w2"], s["b2"], out=s["y"], mask_bits=s["bits"])

        print(f"d={d} T={T} density={density}: plain {timed(plain):6.1f} | two launches {timed(two):6.1f} | one launch {timed(one):6.1f}", flush=True)
