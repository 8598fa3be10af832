"""A short run of the HBM-bound kernels at the UNet-batch-16 shape of the d = 320 layers (T = 65 536 tokens,
64 experts of 20 neurons, k = 19) for ncu: router (select + histogram + masking), router (select only),
permutation (count + scatter), histogram over 60 MB of labels.  Each kernel is launched 3 times."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
T, E, es, k = 65536, 64, 20, 19
gen = torch.Generator(device=dev).manual_seed(0)
scores = torch.randn(T, E, generator=gen, device=dev)
H = torch.empty(T, E * es, dtype=torch.bfloat16, device=dev).normal_(generator=gen)
hist = torch.zeros(E, dtype=torch.int64, device=dev)
for _ in range(3):
    M.router_topk(scores, k, want_bits=False, hist=hist, H=H, expert_size=es, count_rows=(0, 4096))
for _ in range(3):
    bits, idx = M.router_topk(scores, k, want_bits=True, want_idx=True)
for _ in range(3):
    perm = M.expert_permutation(bits, E, k)
big = idx.repeat(24, 1)
for _ in range(3):
    M.hist_accumulate(big, E, hist)
torch.cuda.synchronize()
print("ok", int(hist.sum()), perm.offsets[-1].item())
