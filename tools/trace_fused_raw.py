"""Raw %globaltimer stamps of the fused kernel for a few CTAs (trace build: MOE_LIB_VARIANT=trace build.py first):
commit / epilogue begin / done of the first tiles and the MMA / producer warps' stamps of tile 4 (valid for <= 10 items per
pair, i.e. T <= 8192 at d = 320) -- profiles/r02_phase1_cadence.log.    python tools/trace_fused_raw.py [T]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
os.environ.setdefault("MOE_LIB_VARIANT", "trace")
import moe_b200 as M
from moe_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
d, T, ES = 320, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 20
h = 4 * d; E = h // ES; k = int(E * 0.3)
gen = torch.Generator().manual_seed(0)
x = torch.randn(T, d, generator=gen).to(dev, torch.bfloat16)
w1 = (torch.randn(2 * h, d, generator=gen) / d ** 0.5).to(dev, torch.bfloat16)
b1 = torch.zeros(2 * h, device=dev)
w2 = (torch.randn(d, h, generator=gen) / h ** 0.5).to(dev, torch.bfloat16)
b2 = torch.zeros(d, device=dev)
H = torch.empty(T, h, dtype=torch.bfloat16, device=dev); sc = torch.empty(T, E, device=dev)
y = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
fn = lambda: M.ffn_fused(x, w1, b1, w2, b2, E, ES, k, H_out=H, scores_out=sc, out=y)
def trace():
    buf = (ctypes.c_ulonglong * (256 * 64))()
    assert lib.moe_debug_trace_fused(buf, 256 * 64) == 0
    return torch.tensor(list(buf), dtype=torch.int64).view(256, 64)
for _ in range(50): fn()
torch.cuda.synchronize(); trace(); fn(); torch.cuda.synchronize()
tr = trace()
t0 = tr[:148, 0].min()
for cta in (0, 2, 40):
    r = (tr[cta] - t0).double() / 1e3
    names = {3: "firstMMA", 7: "firstTMA", 8: "c0", 12: "c1", 16: "c2", 20: "c3", 24: "c4", 28: "c5", 9: "eb0", 10: "ed0", 13: "eb1", 14: "ed1", 17: "eb2", 18: "ed2", 21: "eb3", 22: "ed3", 25: "eb4", 26: "ed4",
             50: "m4_reach_tmem", 51: "m4_stagefree", 52: "m4_firstkb", 53: "m4_lastkb", 54: "p4_reach_slot", 55: "p4_lastslotfree"}
    print("CTA", cta, " ".join(f"{names[i]}={r[i]:.2f}" for i in sorted(names) if tr[cta, i] > 0))
