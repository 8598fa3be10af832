"""In-kernel timeline of the fused layer kernel.  Needs the trace build:
    MOE_LIB_VARIANT=trace python diffusion-models-moe_b200/moe_b200/build.py
    MOE_LIB_VARIANT=trace python tools/trace_fused.py [d T]..."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
os.environ.setdefault("MOE_LIB_VARIANT", "trace")
import moe_b200 as M
from moe_b200 import _lib
dev = "cuda:0"
lib = _lib.load()
ES = int(os.environ.get("ES", "20"))


def trace():
    buf = (ctypes.c_ulonglong * (256 * 64))()
    assert lib.moe_debug_trace_fused(buf, 256 * 64) == 0, lib.moe_last_error()
    return torch.tensor(list(buf), dtype=torch.int64).view(256, 64)


def show(name, tr, ms):
    tr = tr[tr[:, 0] > 0]
    t0 = tr[:, 0].min()
    rel = (tr - t0).double() / 1e3
    rel[tr == 0] = float("nan")

    def col(i):
        c = rel[:, i]; c = c[~torch.isnan(c)]
        return f"{c.min():6.2f}/{c.median():6.2f}/{c.max():6.2f} ({len(c):3d})" if len(c) else "     -"
    print(f"{name}: {ms*1e3:7.1f}us event-timed, {tr.shape[0]} CTAs; min/median/max over CTAs (n), us since first CTA entry")
    print(f"   entry {col(0)} | setup {col(1)} | pdl_wait {col(2)} | first TMA {col(7)} | first MMA {col(3)}")
    if os.environ.get("MOE_TRACE_P3"):
        # per-item stamps follow the PHASE-3 items: route begin / end (epilogue warps), block ready at the A producer,
        # MMA commit, epilogue begin / done
        for it in range(8):
            b = 8 + 4 * it
            print(f"   p3 item {it}: route begin {col(40+it)} | route end {col(b+3)} | ready seen {col(48+it)} | mma-commit {col(b)} | "
                  f"epi-begin {col(b+1)} | epi-done {col(b+2)}")
        print(f"   A-producer reaches first ready-wait {col(62)} | epi all done {col(4)} | exit {col(6)}")
        return
    for it in range(8):
        b = 8 + 4 * it
        if torch.isnan(rel[:, b]).all() and torch.isnan(rel[:, b + 1]).all():
            break
        print(f"   item {it:2d}: mma-commit {col(b)} | epi-begin {col(b+1)} | epi-done {col(b+2)} | stored+signalled {col(b+3)}")
    print(f"   route: first begin {col(60)} | last end {col(61)} | K3 A-producer: ready-wait begin {col(62)} end {col(63)}")
    for i in range(2):
        print(f"   sync warp, tile {i+2}: hs_full seen {col(40+4*i)} | stores read, buffer released {col(41+4*i)} | "
              f"previous tile's stores complete {col(42+4*i)} | previous tile published {col(43+4*i)}")
    print(f"   split-K epilogue (first phase-3 item): partials stored {col(48)} | counted {col(49)}   (without split-K: A producer knows its first tile | past its first slot wait)")
    def d(i, j):
        c = rel[:, j] - rel[:, i]; c = c[~torch.isnan(c)]
        return f"{c.min():5.2f}/{c.median():5.2f}/{c.max():5.2f}" if len(c) else "-"
    print(f"   routing (last chunk of each CTA): whole item {d(60,61)} | select {d(56,57)} | labels/hist {d(57,58)} | "
          f"zero-writes {d(58,59)} | 59->end {d(59,61)}")
    print(f"   MMA issuer, tile 4 (needs > 5 tiles per pair; us since tile 2 epi-done = its TMEM stage free): reaches tmem_empty wait {d(18,50)} | "
          f"stage free seen {d(18,51)} | first k-block landed {d(18,52)} | last k-block landed {d(18,53)} | commit issued {d(18,24)} | epi-begin {d(18,25)}")
    print(f"   A producer, tile 4 (same origin): reaches first slot wait {d(18,54)} | last slot free seen {d(18,55)} ; tile 3 commit issued {d(18,20)}")
    print(f"   phase 3, first item (pairs with <= 5 phase-1 tiles; absolute us): A producer past the ready-wait {col(54)} | its last slot free {col(55)} | "
          f"MMA thread at tmem wait {col(50)} | past it {col(51)} | first slot landed {col(52)} | last slot landed {col(53)}")
    print(f"   epi all done {col(4)} | teardown sync {col(5)} | exit {col(6)}")


shapes = [(320, 8192), (640, 2048), (1280, 512), (1280, 128)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for d, T in shapes:
    h = 4 * d; E = h // ES; k = int(E * 0.3)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
    w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
    b1 = torch.zeros(2 * h, device=dev)
    w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
    b2 = torch.zeros(d, device=dev)
    H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
    y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    hist = torch.zeros(E, dtype=torch.int64, device=dev)
    fn = lambda: M.ffn_fused(x, w1, b1, w2, b2, E, ES, k, int(os.environ.get("ACT", "0")), hist=hist, count_rows=(0, T // 2), H_out=H, scores_out=sc, out=y, mask_h=bool(int(os.environ.get("MASK_H", "1"))))
    print(f"==== d={d} T={T} es={ES}")
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    trace()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    show("fused", trace(), e0.elapsed_time(e1))
