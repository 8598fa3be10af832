"""First-contact diagnostics for the CUDA kernels on a real B200: prints error statistics for each
kernel against torch on the same inputs (no asserts), so that one gpurun trip tells us what is wrong."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def report(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = err > (1e-2 * ref.abs() + 2e-2 * denom)
    print(f"  {name}: max_abs_err={err.max().item():.4e} ref_max={denom:.4e} mean_err={err.mean().item():.3e} "
          f"bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(got).sum())}")
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"    bad rows: n={len(rows)} first={rows[:8].tolist()} last={rows[-4:].tolist()}")
        print(f"    bad cols: n={len(cols)} first={cols[:8].tolist()} last={cols[-4:].tolist()}")
        r, c = int(rows[0]), int(cols[0])
        print(f"    got[{r},{c}:{c+6}]={got[r, c:c+6].tolist()}\n    ref[{r},{c}:{c+6}]={ref[r, c:c+6].tolist()}")


def run_down(T, h, d):
    print(f"down_proj T={T} h={h} d={d}")
    H = (torch.randn(T, h, device=dev) * 0.5).bfloat16()
    W = (torch.randn(d, h, device=dev) / h ** 0.5).bfloat16()
    b = torch.randn(d, device=dev)
    try:
        Y = M.down_proj(H, W, b)
        torch.cuda.synchronize()
        report("Y", Y, H.float() @ W.float().t() + b)
    except Exception as e:  # noqa: BLE001
        print("  FAILED:", repr(e))


def run_up(T, d, h, es, act=0):
    E = h // es
    print(f"geglu_up T={T} d={d} h={h} es={es} E={E} act={act}")
    x = torch.randn(T, d, device=dev).bfloat16()
    W = (torch.randn(2 * h, d, device=dev) / d ** 0.5).bfloat16()
    b = torch.randn(2 * h, device=dev) * 0.1
    try:
        Hh, sc, gate = M.geglu_up(x, W, b, E, es, act, want_gate=True)
        torch.cuda.synchronize()
        y = x.float() @ W.float().t() + b
        v, g = y[:, :h], y[:, h:]
        g = torch.nn.functional.gelu(g) if act == 0 else torch.relu(g)
        report("H", Hh, v * g)
        report("gate", gate, g)
        report("scores", sc, g.view(T, E, es).sum(-1))
    except Exception as e:  # noqa: BLE001
        print("  FAILED:", repr(e))


def run_router(T, E, k, es=20):
    print(f"router T={T} E={E} k={k}")
    sc = torch.randn(T, E, device=dev)
    hist = torch.zeros(E, dtype=torch.int64, device=dev)
    cm = torch.full((E,), float("-inf"), device=dev)
    H = torch.ones(T, E * es, device=dev).bfloat16()
    try:
        bits, idx = M.router_topk(sc, k, want_idx=True, hist=hist, colmax_out=cm, H=H, expert_size=es, count_rows=(0, T))
        torch.cuda.synchronize()
        ref = torch.topk(sc, k, dim=-1)[1].sort(dim=-1)[0]
        print("  idx equal:", bool((idx.long() == ref).all()), " hist equal:",
              bool((hist == torch.bincount(ref.flatten(), minlength=E)).all()), " colmax equal:",
              bool((cm == sc.max(0)[0]).all()))
        keep = torch.zeros(T, E, device=dev).scatter_(1, ref, 1.0).repeat_interleave(es, dim=1)
        print("  H mask equal:", bool((H.float() == keep).all()))
    except Exception as e:  # noqa: BLE001
        print("  FAILED:", repr(e))


if __name__ == "__main__":
    which = sys.argv[1:] or ["router", "down", "up"]
    if "router" in which:
        run_router(1000, 64, 19)
        run_router(77, 20, 6, es=64)
        run_router(513, 256, 76)
    if "down" in which:
        run_down(128, 64, 16)
        run_down(256, 128, 64)
        run_down(300, 1280, 320)
        run_down(4096, 1280, 320)
    if "up" in which:
        run_up(128, 64, 128, 16)
        run_up(200, 64, 320, 20)
        run_up(4096, 320, 1280, 20)
        run_up(4096, 320, 1280, 64, act=1)
        run_up(512, 1280, 5120, 20)
    # timing of the config-1 layer pieces
    T, d, h, es = 8192, 320, 1280, 20
    E = h // es
    x = torch.randn(T, d, device=dev).bfloat16()
    W1 = (torch.randn(2 * h, d, device=dev) / d ** 0.5).bfloat16()
    b1 = torch.randn(2 * h, device=dev) * 0.1
    W2 = (torch.randn(d, h, device=dev) / h ** 0.5).bfloat16()
    b2 = torch.randn(d, device=dev)
    try:
        for name, fn in [
            ("geglu_up", lambda: M.geglu_up(x, W1, b1, E, es)),
            ("torch up matmul", lambda: x @ W1.t()),
        ]:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                fn()
            t1.record(); torch.cuda.synchronize()
            print(f"time {name}: {t0.elapsed_time(t1) / 20 * 1e3:.1f} us")
        Hh, sc, _ = M.geglu_up(x, W1, b1, E, es)
        for name, fn in [
            ("router+mask", lambda: M.router_topk(sc, 19, H=Hh, expert_size=es)),
            ("down_proj", lambda: M.down_proj(Hh, W2, b2)),
            ("torch down matmul", lambda: Hh @ W2.t()),
        ]:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                fn()
            t1.record(); torch.cuda.synchronize()
            print(f"time {name}: {t0.elapsed_time(t1) / 20 * 1e3:.1f} us")
    except Exception as e:  # noqa: BLE001
        print("timing FAILED:", repr(e))
