"""A few launches of the fused layer kernel on one SD-1.5 layer shape, for ncu:  python tools/run_fused_once.py d T [es]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
dev = "cuda:0"
d, T = int(sys.argv[1]), int(sys.argv[2])
ES = int(sys.argv[3]) if len(sys.argv) > 3 else 20
h = 4 * d; E = h // ES; k = int(E * 0.3)
gen = torch.Generator().manual_seed(0)
x = torch.nn.functional.layer_norm(torch.randn(T, d, generator=gen), (d,)).to(dev, torch.bfloat16)
w1 = ((torch.rand(2 * h, d, generator=gen) * 2 - 1) / d ** 0.5).to(dev, torch.bfloat16)
b1 = ((torch.rand(2 * h, generator=gen) * 2 - 1) / d ** 0.5).to(dev)
w2 = ((torch.rand(d, h, generator=gen) * 2 - 1) / h ** 0.5).to(dev, torch.bfloat16)
b2 = torch.zeros(d, device=dev)
H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
hist = torch.zeros(E, dtype=torch.int64, device=dev)
for _ in range(12):
    M.ffn_fused(x, w1, b1, w2, b2, E, ES, k, hist=hist, count_rows=(0, T // 2), H_out=H, scores_out=sc, out=y)
torch.cuda.synchronize()
print("ok", float(y.float().abs().sum()), int(hist.sum()))
