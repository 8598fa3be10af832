import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import moe_b200 as M
from moe_b200.packing import bits_to_sets
DEV = "cuda:0"
for (Tn, E, es, k) in [(1000, 64, 20, 19), (777, 20, 64, 6), (513, 128, 20, 38), (300, 256, 20, 76), (64, 8, 16, 3)]:
    g = torch.Generator().manual_seed(E + k)
    scores = torch.randn(Tn, E, generator=g)
    plain = scores.clone()
    scores[::4] = torch.round(scores[::4] * 2) / 2
    scores[1::7] = scores[1::7].abs() * 3 + 1
    for name, sc in (("plain", plain), ("ties", scores)):
        d = sc.to(DEV)
        os.environ["MOE_ROUTER_LEGACY"] = "1"
        b0, i0 = M.router_topk(d, k, want_idx=True)
        os.environ["MOE_ROUTER_LEGACY"] = "0"
        b1, i1 = M.router_topk(d, k, want_idx=True)
        torch.cuda.synchronize()
        bad = (b0 != b1).any(1).nonzero().flatten().tolist()
        print(Tn, E, k, name, "mismatching tokens:", len(bad), bad[:10])
        want = torch.topk(sc, k, dim=-1)[1].sort(dim=-1)[0]
        print("   legacy==topk", bool((i0.cpu().long() == want).all()), " new==topk", bool((i1.cpu().long() == want).all()))
        if bad:
            t = bad[0]
            s0, s1 = bits_to_sets(b0[t:t+1], E)[0], bits_to_sets(b1[t:t+1], E)[0]
            print("   token", t, "legacy-new", sorted(s0 - s1), "new-legacy", sorted(s1 - s0), "len", len(s0), len(s1))
            srt = torch.sort(sc[t], descending=True)
            print("   kth", srt[0][k-2:k+2].tolist(), srt[1][k-2:k+2].tolist())
