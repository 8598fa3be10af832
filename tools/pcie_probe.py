"""Host<->device copy ceilings for the e2e leg of bench.py: the same 16 layer-sized pinned buffers copied H2D only,
D2H only and both directions at once (two streams), timed with CUDA events.  Run on the GPU box."""
import torch

SHAPES = [(8192, 320)] * 6 + [(2048, 640)] * 5 + [(512, 1280)] * 4 + [(128, 1280)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    host_in = [torch.empty(s, dtype=torch.bfloat16).pin_memory() for s in SHAPES]
    host_out = [torch.empty(s, dtype=torch.bfloat16).pin_memory() for s in SHAPES]
    dev_in = [torch.empty(s, dtype=torch.bfloat16, device=dev) for s in SHAPES]
    dev_out = [torch.empty(s, dtype=torch.bfloat16, device=dev) for s in SHAPES]
    nbytes = sum(t.numel() * 2 for t in host_in)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def h2d():
        for h, d in zip(host_in, dev_in):
            d.copy_(h, non_blocking=True)

    def d2h():
        for h, d in zip(host_out, dev_out):
            h.copy_(d, non_blocking=True)

    def both():
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            h2d()
        with torch.cuda.stream(s2):
            d2h()
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    for name, fn in (("H2D only", h2d), ("D2H only", d2h), ("both directions", both)):
        ms = timed(fn)
        print(f"{name:16s}: {ms:.3f} ms for {nbytes / 1e6:.1f} MB per direction = {nbytes / ms / 1e6:.1f} GB/s per direction")


if __name__ == "__main__":
    main()
