"""Multi-GPU bit-exactness of the sharded sampling run (BASELINE configs[3]), on real GPUs over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_sampling_check.py [--batches B] [--steps S] [--latent HW]

Every rank builds the same MoEfied FFN stack and the same RemoveExperts receiver (CUDA path, fused layer kernel behind
the hooks, one CUDA graph per sampling run), takes prompt batches r, r + N, ... and accumulates the per-(timestep, layer)
expert counters on its GPU; ONE int64 all-reduce (NCCL) sums them.  Rank 0 then runs ALL prompt batches alone and the
two [T, 16, E_max] histograms must be equal bit for bit.  Prints `BITEXACT OK ...` or raises."""
import argparse, json, os, sys, tempfile
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=6)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--prompts", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import neuron_receivers as nr
    from moefication import helper
    from moe_b200.sd_modules import FFNStackUNet, SyntheticFFNPipeline, GraphedSampling, sd_ffn_shapes

    torch.manual_seed(0)                       # identical weights on every rank
    pipe = SyntheticFFNPipeline(FFNStackUNet(latent_hw=args.latent), num_inference_steps=args.steps, device=dev)
    shapes = sd_ffn_shapes(args.latent)
    labels = {n + ".proj.weight": np.random.RandomState(i).permutation(np.repeat(np.arange(h // 20), 20))
              for i, (n, d, h, s) in enumerate(shapes)}

    class A:
        res_path = ""
        moefication = {"topk_experts": 0.3}
    pipe, names, n_exp = helper.modify_ffn_to_experts(pipe, A(), labels_by_name=labels)
    e_max = max(n_exp.values())
    hist = torch.zeros(args.steps, 16, e_max, dtype=torch.int64, device=dev)
    with tempfile.TemporaryDirectory() as td:
        rs = np.random.RandomState(2)
        for t in range(args.steps):
            for l, nm in enumerate(names):
                E = n_exp[nm]
                json.dump(sorted(int(v) for v in rs.choice(E, E // 10, replace=False)) if t < 3 else [],
                          open(os.path.join(td, f"timestep_{t}_layer_{l}.json"), "w"))
        rec = nr.RemoveExperts(0, td, args.steps, 16, capture_gates=False, hist=hist, count_rows='all')
    gs = GraphedSampling(pipe, rec, args.prompts, args.steps)
    hist.zero_()
    for b in range(rank, args.batches, world):
        gs.load_states(b)
        gs.replay()
    sharded = hist.clone()
    if world > 1:
        dist.all_reduce(sharded, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    if rank == 0:
        hist.zero_()
        for b in range(args.batches):
            gs.load_states(b)
            gs.replay()
        torch.cuda.synchronize()
        tokens = 2 * args.prompts * sum(s for (_, _, _, s) in shapes)
        assert int(hist.sum()) > 0
        assert torch.equal(hist, sharded), f"sharded histogram differs from the single-GPU run in {(hist != sharded).sum().item()} cells"
        print(f"BITEXACT OK world={world} batches={args.batches} steps={args.steps} cells={hist.numel()} "
              f"selections={int(hist.sum())} tokens_per_step={tokens}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
