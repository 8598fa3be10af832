"""Per-step cycle accounting of the producer / MMA-issue threads (MOE_DEBUG_MODE bit 16)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
from moe_b200 import _lib
dev = "cuda:0"
lib = _lib.load()
PAIR = os.environ.get("MOE_PAIR", "1")
def counters(n_cta):
    buf = (ctypes.c_ulonglong * 2048)()
    assert lib.moe_debug_counters(buf, 2048) == 0
    allc = torch.tensor(list(buf), dtype=torch.float64).view(256, 8)
    global EPI
    EPI = allc[128:256]      # second bank: epilogue sums of CTAs 0..127
    return allc[:n_cta]
for d, T in [(320, 8192), (640, 2048)]:
    h = 4 * d; es = 20; E = h // es
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = torch.zeros(2 * h, device=dev)
    w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = torch.zeros(d, device=dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
    y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    for mode in (16,):
        os.environ["MOE_DEBUG_MODE"] = str(mode)
        for name, fn in (("K1", lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc)), ("K3", lambda: M.down_proj(H, w2, b2, out=y))):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            c = counters(128)
            c = c[c[:, 2] > 0]          # CTAs that ran a producer loop
            lead = c[c[:, 4] > 0]       # CTAs whose MMA thread ran (pair leaders)
            it = c[:, 2].clamp(min=1)
            m = (c / it[:, None]).mean(0)
            if len(lead):
                ml = (lead / lead[:, 2:3].clamp(min=1)).mean(0)
                m[3:7] = ml[3:7]
            ep = EPI[EPI[:, 3] > 0]
            es = (ep[:, :3] / ep[:, 3:4]).mean(0) if len(ep) and name == "K1" else None
            if es is not None:
                print(f"      K1 epilogue per tile (cycles): wait_acc_full={es[0]:.0f} math={es[1]:.0f} store+barrier={es[2]:.0f} tiles/CTA={ep[:,3].mean():.1f}")
            print(f"pair={PAIR} d={d} T={T} mode={mode:2d} {name}: {e0.elapsed_time(e1)*1e3:7.1f}us iters/CTA={it.mean():.0f} per-iter cycles: "
                  f"P.wait_empty={m[0]:.0f} P.issue={m[1]:.0f} | M.wait_acc={m[3]:.0f} M.wait_full={m[4]:.0f} M.mma_issue={m[5]:.0f} M.commit={m[6]:.0f}")
