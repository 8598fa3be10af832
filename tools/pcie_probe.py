"""Host<->device copy ceilings for the e2e leg of bench.py: the same 16 layer-sized pinned buffers copied H2D only,
D2H only and both directions at once (two streams), timed with CUDA events.  Run on the GPU box.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
                                                                 # N ranks copying CONCURRENTLY (no kernels running):
                                                                 # per-rank and aggregate GB/s = the host-side ceiling"""
import os

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
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
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
        if world > 1:
            dist.barrier()              # all ranks copy at the same time
        ms = timed(fn, reps=50)
        gbs = nbytes / ms / 1e6
        if world > 1:
            t = torch.tensor([gbs, ms], dtype=torch.float64, device=dev)
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            if rank == 0:
                per = [float(v[0]) for v in allv]
                print(f"{name:16s}: {world} ranks concurrently, {nbytes / 1e6:.1f} MB per direction and rank: per rank "
                      f"{min(per):.1f} .. {max(per):.1f} GB/s per direction, aggregate {sum(per):.1f} GB/s per direction "
                      f"(slowest rank {max(float(v[1]) for v in allv):.3f} ms)", flush=True)
        else:
            print(f"{name:16s}: {ms:.3f} ms for {nbytes / 1e6:.1f} MB per direction = {gbs:.1f} GB/s per direction")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
