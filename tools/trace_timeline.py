"""In-kernel timeline of K1 / K3 (MOE_DEBUG_MODE bit 32): per-CTA %globaltimer stamps, printed relative to the
earliest CTA entry.  Usage: python tools/trace_timeline.py [d T]..."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
os.environ["MOE_DEBUG_MODE"] = str(int(os.environ.get("MOE_DEBUG_MODE", "0")) | 32)
import moe_b200 as M
from moe_b200 import _lib
dev = "cuda:0"
lib = _lib.load()

def trace():
    buf = (ctypes.c_ulonglong * (256 * 64))()
    assert lib.moe_debug_trace(buf, 256 * 64) == 0
    return torch.tensor(list(buf), dtype=torch.int64).view(256, 64)

def clear():
    # stamps are only overwritten by CTAs that run; a launch of a tiny dummy is not needed: compare against entry
    pass

def show(name, tr, ms):
    used = tr[:, 0] > 0
    tr = tr[used]
    t0 = tr[:, 0].min()
    rel = (tr - t0).double() / 1e3      # us
    rel[tr == 0] = float("nan")
    n = tr.shape[0]
    def col(i):
        c = rel[:, i]; c = c[~torch.isnan(c)]
        return (f"{c.min():6.2f}/{c.median():6.2f}/{c.max():6.2f}" if len(c) else "     -")
    print(f"{name}: {ms*1e3:7.1f}us event-timed, {n} CTAs; columns min/median/max over CTAs (us since first CTA entry)")
    print(f"   entry {col(0)} | setup {col(1)} | pdl_wait {col(2)} | first TMA {col(7)} | first MMA {col(3)}")
    print(f"   producer: loop entry {col(60)} | empty-wait passed {col(61)} | before TMA issue {col(62)}")
    for it in range(14):
        b = 8 + 4 * it
        if torch.isnan(rel[:, b]).all() and torch.isnan(rel[:, b + 1]).all():
            break
        print(f"   tile {it:2d}: mma-commit {col(b)} | epi-begin {col(b+1)} | epi-math-done {col(b+2)} | store {col(b+3)}")
    print(f"   epi all done {col(4)} | teardown sync {col(5)} | exit {col(6)}")

shapes = [(320, 8192), (640, 2048), (1280, 512), (1280, 128)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for d, T in shapes:
    h = 4 * d; es = int(os.environ.get("ES", "20")); E = h // es
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = torch.zeros(2 * h, device=dev)
    w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = torch.zeros(d, device=dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
    y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    print(f"==== d={d} T={T} es={es}")
    for name, fn in (("K1", lambda: M.geglu_up(x, w1, b1, E, es, out=H, scores_out=sc)),
                     ("K3", lambda: M.down_proj(H, w2, b2, out=y))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a second launch right behind the first, so that the traced (second) one sees a PDL predecessor
        trace()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        show(name, trace(), e0.elapsed_time(e1))
