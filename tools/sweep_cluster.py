"""Per-shape timing of K1 / K2 / K3 (CUDA-graph replays, CUDA events) for every cluster shape.
    python tools/sweep_cluster.py [reference|literal]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M  # noqa: E402

dev = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "reference"
SHAPES = [(320, 8192), (640, 2048), (1280, 512), (1280, 128), (320, 65536), (1280, 4096)]
CLUSTERS = ["1,1", "2,1", "1,2", "2,2", "4,1", "1,4", "4,2", "2,4"]
REPS = 10


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REPS)


for d, T in SHAPES:
    h = 4 * d
    es = 20 if mode == "reference" else h // 20
    E = h // es
    gen = torch.Generator().manual_seed(0)
    x = torch.nn.functional.layer_norm(torch.randn(T, d, generator=gen), (d,)).to(dev, torch.bfloat16)
    w1 = ((torch.rand(2 * h, d, generator=gen) * 2 - 1) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = ((torch.rand(2 * h, generator=gen) * 2 - 1) / d ** 0.5).to(dev)
    w2 = ((torch.rand(d, h, generator=gen) * 2 - 1) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = ((torch.rand(d, generator=gen) * 2 - 1) / h ** 0.5).to(dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev)
    sc = torch.empty(T, E, dtype=torch.float32, device=dev)
    y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    hist = torch.zeros(E, dtype=torch.int64, device=dev)
    k1f, k3f = 4.0 * d * h * T, 2.0 * d * h * T
    line1, line3 = [], []
    for c in CLUSTERS:
        os.environ["MOE_K1_CLUSTER"] = c
        os.environ["MOE_K3_CLUSTER"] = c
        try:
            t1 = timed(lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc))
            line1.append(f"{c}:{t1:6.1f}us({k1f / t1 / 1e6:4.0f}TF)")
        except Exception as e:  # noqa: BLE001
            line1.append(f"{c}:ERR")
        try:
            t3 = timed(lambda: M.down_proj(H, w2, b2, out=y))
            line3.append(f"{c}:{t3:6.1f}us({k3f / t3 / 1e6:4.0f}TF)")
        except Exception as e:  # noqa: BLE001
            line3.append(f"{c}:ERR")
    os.environ.pop("MOE_K1_CLUSTER"); os.environ.pop("MOE_K3_CLUSTER")
    t1 = timed(lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc))
    t3 = timed(lambda: M.down_proj(H, w2, b2, out=y))
    t2 = timed(lambda: M.router_topk(sc, int(E * 0.3), want_bits=False, hist=hist, H=H, expert_size=es, count_rows=(0, T // 2)))
    tb1 = timed(lambda: torch.matmul(x, w1.t()))
    tb3 = timed(lambda: torch.matmul(H, w2.t()))
    print(f"d={d} T={T} es={es}")
    print("  K1 " + " ".join(line1))
    print("  K3 " + " ".join(line3))
    print(f"  auto: K1 {t1:.1f}us K2 {t2:.1f}us K3 {t3:.1f}us | cuBLAS up {tb1:.1f}us ({k1f / tb1 / 1e6:.0f}TF) down {tb3:.1f}us ({k3f / tb3 / 1e6:.0f}TF)")
