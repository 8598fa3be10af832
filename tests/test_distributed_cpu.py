"""CPU, world_size 2, gloo: the multi-GPU plan of DESIGN.md section 5 -- prompts sharded r::W, one
all_reduce(SUM) of the int64 expert histogram -- gives exactly the single-process counters."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import moe_ffn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


N_PROMPTS, T_STEPS, N_LAYERS, E, K, S, D, H = 6, 2, 3, 8, 2, 12, 16, 64


def _prompt_counts(prompt):
    """int64 [T_STEPS, N_LAYERS, E] selection counts of one prompt (oracle arithmetic, row 0 only)."""
    layer = O.synthetic_layer(D, H, (2, S), H // E, seed=0)
    pat = O.patterns_from_labels(layer["labels"])
    out = np.zeros((T_STEPS, N_LAYERS, E), dtype=np.int64)
    for t in range(T_STEPS):
        for l in range(N_LAYERS):
            x = torch.randn(2, S, D, generator=torch.Generator().manual_seed(prompt * 100 + t * 10 + l))
            _, labels, _, _ = O.moefy_forward(x, layer["w1"], layer["b1"], pat, K)
            out[t, l] = O.selection_counts(labels, E)
    return out


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "diffusion-models-moe_b200"), os.path.join(root, "oracle")]
    import neuron_receivers as nr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [f"l{i}" for i in range(N_LAYERS)]
    rec = nr.FrequencyMeasure(0, T_STEPS, N_LAYERS, {n: E for n in names}, names, device="cpu")
    hist = rec.int_counts()
    for prompt in range(rank, N_PROMPTS, world):           # shard: rank r takes prompts r::W
        hist += torch.from_numpy(_prompt_counts(prompt))
    local_sum = int(hist.sum())
    total = rec.all_reduce()
    q.put((rank, local_sum, total.clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_histogram_allreduce_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sum(_prompt_counts(p) for p in range(N_PROMPTS))
    for rank, local_sum, total in results:
        assert np.array_equal(total, want)                  # bit-exact integer counters on every rank
        assert local_sum == (N_PROMPTS // world) * T_STEPS * N_LAYERS * S * K
    assert int(want.sum()) == N_PROMPTS * T_STEPS * N_LAYERS * S * K


def test_all_reduce_is_noop_without_process_group():
    import neuron_receivers as nr
    names = ["a"]
    rec = nr.FrequencyMeasure(0, 1, 1, {"a": 4}, names, device="cpu")
    rec.int_counts()[0, 0, 1] = 5
    assert int(rec.all_reduce().sum()) == 5
