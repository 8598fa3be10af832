"""CUDA-event timing of the HBM-bound side kernels at the UNet-batch-16 shape of the d = 320 layers (T = 65 536, E = 64,
es = 20, k = 19): router with masking, router select-only, histogram over 0.36 GB of labels, permutation.  REP distinct
buffer sets per graph so that nothing is L2-resident between launches.  Env switches of the library (MOE_ROUTER_LANES ...)
are read per call, so variants are set in-process:   python tools/time_aux.py [lanes ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
T, E, es, k = 65536, 64, 20, 19
REP = 4
gen = torch.Generator(device=dev).manual_seed(7)
bufs = [dict(scores=torch.randn(T, E, generator=gen, device=dev),
             H=torch.empty(T, E * es, dtype=torch.bfloat16, device=dev).normal_(generator=gen),
             hist=torch.zeros(E, dtype=torch.int64, device=dev)) for _ in range(REP)]


def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def masked():
    for b in bufs:
        M.router_topk(b["scores"], k, want_bits=False, want_idx=False, hist=b["hist"], H=b["H"], expert_size=es, count_rows=(0, 4096))


idx = []


def select():
    idx.clear()
    for b in bufs:
        _, ix = M.router_topk(b["scores"], k, want_bits=False, want_idx=True)
        idx.append(ix)


by_mask = T * (4 * E + 2 * es * (E - k)); by_sel = T * (4 * E + 2 * k)
variants = sys.argv[1:] or ["default"]
for v in variants:
    if v == "default":
        os.environ.pop("MOE_ROUTER_LANES", None)
    else:
        os.environ["MOE_ROUTER_LANES"] = v
    um, us = timed(masked) / REP, timed(select) / REP
    print(f"router lanes={v:8s}: masked {um:7.2f} us ({by_mask / um / 1e3:7.1f} GB/s) | select-only {us:7.2f} us ({by_sel / us / 1e3:7.1f} GB/s)")
os.environ.pop("MOE_ROUTER_LANES", None)
select()
big = idx[0].repeat(144, 1)
hist = torch.zeros(E, dtype=torch.int64, device=dev)
uh = timed(lambda: M.hist_accumulate(big, E, hist))
ref = torch.bincount(big.flatten().long(), minlength=E)
hist.zero_(); M.hist_accumulate(big, E, hist); torch.cuda.synchronize()
print(f"hist: {uh:7.2f} us for {big.numel() * 2 / 1e6:.0f} MB = {big.numel() * 2 / uh / 1e3:7.1f} GB/s | exact={bool(torch.equal(hist, ref))}")
bits, _ = M.router_topk(bufs[0]["scores"], k, want_bits=True, want_idx=False)
up = timed(lambda: M.expert_permutation(bits, E, k))
print(f"permutation (count + scatter): {up:7.2f} us for T={T}")
