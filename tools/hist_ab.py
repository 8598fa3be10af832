import os, sys, torch
sys.path.insert(0, "diffusion-models-moe_b200")
import moe_b200 as M
dev="cuda:0"
for E in (64, 20, 128, 256):
    idx = torch.randint(0, E, (180_000_000,), dtype=torch.int16, device=dev)
    hist = torch.zeros(E, dtype=torch.int64, device=dev)
    for mode in ("1", "0"):
        os.environ["MOE_HIST_WIDE"] = mode
        for _ in range(3): M.hist_accumulate(idx, E, hist)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): M.hist_accumulate(idx, E, hist)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"E={E} wide={mode}: {us:7.1f} us = {idx.numel()*2/us/1e3:7.1f} GB/s", flush=True)
