"""Run the three hot-path kernels on one layer shape a few times (short command for `ncu --set full`).
    python tools/profile_layer.py [d] [tokens] [es] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 320
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
es = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
h = 4 * d
E = h // es
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.nn.functional.layer_norm(torch.randn(T, d, generator=g), (d,)).to(dev, torch.bfloat16)
w1 = ((torch.rand(2 * h, d, generator=g) * 2 - 1) / d ** 0.5).to(dev, torch.bfloat16)
b1 = ((torch.rand(2 * h, generator=g) * 2 - 1) / d ** 0.5).to(dev)
w2 = ((torch.rand(d, h, generator=g) * 2 - 1) / h ** 0.5).to(dev, torch.bfloat16)
b2 = ((torch.rand(d, generator=g) * 2 - 1) / h ** 0.5).to(dev)
hist = torch.zeros(E, dtype=torch.int64, device=dev)
for _ in range(reps):
    H, sc, _ = M.geglu_up(x, w1, b1, E, es)
    M.router_topk(sc, int(E * 0.3), want_bits=False, hist=hist, H=H, expert_size=es, count_rows=(0, T // 2))
    y = M.down_proj(H, w2, b2)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), int(hist.sum()))
