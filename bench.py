#!/usr/bin/env python
"""bench.py -- the MoEfied GEGLU-FFN hot path of SD-1.5 on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--experts reference|literal]

One *step* = one denoising step of the hot path: the 16 transformer-block FFNs of the SD-1.5 UNet
(BASELINE.json configs[1]: batch 2 = CFG, 64x64 latents, bf16), each as ONE moe_ffn_fused launch (up-projection
+ GELU gate -> per-token top-k routing + row-0 expert histogram + masking -> down-projection; --path split runs the
K1 geglu_up -> K2 router -> K3 down_proj launches instead), through the C ABI of libmoe_b200.so, on synthetic hidden
states and random-init weights of the SD-1.5 geometry.

  value     tokens/s = token-FFN evaluations per second, whole job (N GPUs x 2 x 26 944 per step);
            inputs resident in HBM, the step replayed as one CUDA graph; time = CUDA events, max over ranks.
  e2e       same metric with HOST buffers: per step the 16 layer inputs are copied from pinned host memory,
            the FFNs run through the C ABI, outputs + histogram are copied back (all inside the timing); copy-in,
            compute and copy-out run on three streams chained per layer, consecutive steps pipelined (PCIe-bound:
            tools/pcie_probe.py measures the two-direction copy ceiling for the same bytes).
  roofline  dominant kernel (ffn_fused_kernel): algorithmic FLOPs (6 d h per token) / CUDA-event time of its launches,
            against the burst bf16 peak of MEASURED_PEAKS.json (sustained peak when the sampled clocks show a power
            cap); `frac_necessary` counts the down-projection at 2 d es k (active experts only).
  roofline_router / roofline_hist   K2 router_topk_kernel and K4 hist_accumulate_kernel at the UNet-batch-16 shapes
            (BASELINE configs[2]): algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth.
  sampling  BASELINE configs[2] / [3]: 50-step sampling, 8 prompts per batch (UNet batch 16), THROUGH the receiver
            hook API (RemoveExperts with skilled-expert lists for t < 20 + per-timestep expert counters), each
            sampling run one CUDA graph captured through the hooks; prompt batches sharded r::W over the ranks, one
            int64 all-reduce of the [50, 16, 256] histogram at the end, inside the timing.
  cpu_baseline  the oracle port of the reference's hook arithmetic (fp32 torch-CPU, all host threads), timed
            here on rank 0 at N=1 on a bounded sample (ONE step = all 16 layers).
  --impl reference  times that CPU arm alone for K steps, all 16 layers every step (the reference is pure Python on
            ATen and needs diffusers and /root/reference, neither of which exists on the GPU box; the oracle port is
            bit-identical to the reference classes on CPU -- oracle/gen_golden.py asserts it).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))

import torch  # noqa: E402

RATIO = 0.3
# (d, tokens per sample at 64x64 latents, number of FFN layers of that shape), firing order groups
SD15_LAYERS = [(320, 4096, 2), (640, 1024, 2), (1280, 256, 2), (1280, 64, 1), (1280, 256, 3), (640, 1024, 3),
               (320, 4096, 3)]


def layer_list():
    out = []
    for d, s, n in SD15_LAYERS:
        out += [(d, 4 * d, s)] * n
    return out


def expert_size_for(mode, h):
    # reference: 20 neurons per expert (experiments/moefy_config.yaml:3 -> 64/128/256 experts);
    # literal:   BASELINE.json config 1's geometry, 20 experts of 64 neurons at h=1280 -> 64-neuron experts
    return 20 if mode == "reference" else 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--experts", default="reference", choices=["reference", "literal"])
    ap.add_argument("--batch", type=int, default=2, help="UNet batch (2 = one prompt with CFG)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=2, help="copy groups per step in the e2e leg (1..16)")
    ap.add_argument("--path", default="fused", choices=["fused", "split"],
                    help="fused: one moe_ffn_fused launch per layer; split: K1 -> K2 -> K3 launches")
    ap.add_argument("--sampling-batches", type=int, default=2,
                    help="prompt batches (8 prompts x 50 steps each) per rank in the sampling leg; 0 = skip")
    ap.add_argument("--sampling-steps", type=int, default=50)
    ap.add_argument("--sampling-prompts", type=int, default=8, help="prompts per batch (UNet batch = 2 x prompts)")
    ap.add_argument("--no-aux", action="store_true", help="skip the router / histogram roofline legs")
    return ap.parse_args()


def workload_config(args, world):
    """The workload, identically worded for both arms (the driver compares the two `config` objects)."""
    tokens = args.batch * sum(s for _, _, s in layer_list())
    return dict(workload="SD-1.5 MoEfied UNet, all 16 transformer-block FFNs, one denoising step, "
                         f"batch {args.batch} (CFG), 64x64 latents (BASELINE configs[1])",
                experts=f"{args.experts}: expert_size {expert_size_for(args.experts, 1280)} @h=1280, top-k ratio {RATIO}",
                tokens_per_step_per_gpu=tokens, parallelism=f"prompt-sharded x{world}",
                l2="working set ~0.5 GB per step > 126 MB L2; no explicit flush",
                counters="row-0 expert-selection histogram per layer call")


# --------------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, steps, warmup):
    """The reference's hook arithmetic (oracle port) on the host cores.  One step = ALL 16 layers of the sweep, each
    with its own weights and inputs; per layer call the reference pays (a) the stock GEGLU forward whose result the
    forward hook discards, (b) the hook (MOEFy.hook_fn + counters) and (c) the stock
    down-projection.  `value` counts (b) + (c) only -- the arithmetic our kernels replace; `with_stock_forward`
    adds (a), what a user of the reference really waits for."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import moe_ffn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    layers = []
    for li, (d, h, s) in enumerate(layer_list()):
        es = expert_size_for(args.experts, h)
        layer = O.synthetic_layer(d, h, (args.batch, s), es, seed=100 + li)
        pat = O.patterns_from_labels(layer["labels"])
        layers.append((layer, pat, O.topk_from_ratio(pat.shape[0], RATIO)))
    tokens_per_step = args.batch * sum(s for _, _, s in layer_list())

    def step():
        t_hook = t_stock = 0.0
        for layer, pat, k in layers:
            with torch.no_grad():
                t0 = time.perf_counter()
                v, g = O.geglu_up(layer["x"], layer["w1"], layer["b1"])     # stock GEGLU.forward, discarded by the hook
                _ = v * g
                t1 = time.perf_counter()
                H, labels, _, _ = O.moefy_forward(layer["x"], layer["w1"], layer["b1"], pat, k)
                O.down_proj(H, layer["w2"], layer["b2"])
                O.selection_counts(labels, pat.shape[0])
                t2 = time.perf_counter()
            t_stock += t1 - t0
            t_hook += t2 - t1
        return t_hook, t_stock

    for _ in range(warmup):
        step()
    wall0 = time.perf_counter()
    times = [step() for _ in range(steps)]
    wall = time.perf_counter() - wall0
    hook = statistics.median(t[0] for t in times)
    both = statistics.median(t[0] + t[1] for t in times)
    return dict(value=tokens_per_step / hook, unit="tokens/s", cores=cores, kind="port",
                sample=f"{steps} steps x all 16 layers (own weights and inputs each), batch {args.batch}; "
                       "oracle.moefy_forward + down_proj + counts, fp32 torch-CPU, median step",
                ms_per_step=hook * 1e3,
                with_stock_forward=dict(value=tokens_per_step / both, unit="tokens/s", ms_per_step=both * 1e3,
                                        note="+ the stock GEGLU forward the reference's forward hook discards"),
                timed_region_s=wall), tokens_per_step


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs local to GPU `index` (sysfs local_cpulist of its PCI function), so that the pinned
    host arenas of the e2e leg are allocated on the NUMA node the GPU hangs off.  Best effort: returns a description."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(f"{base}/numa_node").read().strip()
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"gpu {bdf} node {node}: bound to {len(use)} of {len(allowed)} cpus"
        return f"gpu {bdf} node {node}: {len(allowed)} cpus allowed, no narrowing"
    except Exception as e:      # noqa: BLE001 -- sysfs layout differs between hosts; the bench runs unbound then
        return f"unbound ({type(e).__name__})"


def graph_of(fn, dev):
    """Capture fn() (already warmed up) on a side stream; returns the replayable graph."""
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    return g


def time_graph(g, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n      # ms per replay


def aux_rooflines(M, args, dev, hbm_gbs, peak_src):
    """K2 router_multi_kernel and K4 hist_accumulate_kernel against HBM at the UNet-batch-16 shape of the d = 320
    layers (BASELINE configs[2]: T = 65 536 tokens, the shape that dominates a batch-16 step), each kernel timed alone
    (CUDA events around a graph of launches over REP distinct buffer sets, so nothing is L2-resident between launches),
    plus the 16-launch sweep over all layer shapes of one batch-16 step.  Algorithmic bytes per token (DESIGN.md
    section 4): K2 reads 4 E (scores), writes 2 k (labels) or, when masking, the zero stores 2 es (E - k); K4 reads
    2 k (labels)."""
    B = 16
    gen = torch.Generator(device=dev).manual_seed(7)
    d0, h0, s0 = layer_list()[0]
    es0 = expert_size_for(args.experts, h0)
    E0, T0 = h0 // es0, B * s0
    k0 = int(E0 * RATIO)
    REP = 4
    bufs = [dict(scores=torch.randn(T0, E0, generator=gen, device=dev),
                 H=torch.empty(T0, h0, dtype=torch.bfloat16, device=dev).normal_(generator=gen),
                 hist=torch.zeros(E0, dtype=torch.int64, device=dev)) for _ in range(REP)]

    def router_mask():
        for b in bufs:
            M.router_topk(b["scores"], k0, want_bits=False, want_idx=False, hist=b["hist"], H=b["H"], expert_size=es0,
                          count_rows=(0, s0))

    idx = []

    def router_select():
        idx.clear()
        for b in bufs:
            _, ix = M.router_topk(b["scores"], k0, want_bits=False, want_idx=True)
            idx.append(ix)

    router_mask(); router_select()
    torch.cuda.synchronize()
    us_mask = time_graph(graph_of(router_mask, dev), 10) / REP * 1e3
    us_sel = time_graph(graph_of(router_select, dev), 10) / REP * 1e3
    by_mask = T0 * (4 * E0 + 2 * es0 * (E0 - k0))
    by_sel = T0 * (4 * E0 + 2 * k0)
    shape = f"T={T0} E={E0} es={es0} k={k0}"
    r_router = dict(bound="hbm", kernel="router_multi_kernel (K2: select + histogram + write-only masking of H)", shape=shape,
                    achieved=round(by_mask / us_mask / 1e3, 1), peak=hbm_gbs, unit="GB/s",
                    frac=round(by_mask / us_mask / 1e3 / hbm_gbs, 4), us_per_launch=round(us_mask, 2),
                    # ncu --set full of this launch (profiles/r02_ncu_aux_kernels.csv): dram read + write bytes
                    traffic=1.65e8, peak_source=peak_src,
                    algorithmic=f"per token 4E bytes of scores read + 2 es (E - k) bytes of zero stores = {by_mask / 1e6:.0f} MB per launch",
                    select_only=dict(achieved=round(by_sel / us_sel / 1e3, 1), unit="GB/s", frac=round(by_sel / us_sel / 1e3 / hbm_gbs, 4),
                                     us_per_launch=round(us_sel, 2), traffic=1.68e7,
                                     algorithmic=f"per token 4E read + 2k labels written = {by_sel / 1e6:.0f} MB per launch",
                                     note="bound by the sorting network (integer pipe: ~300 warp instructions per token, "
                                          "ncu source page), not by HBM"))
    # K4 over the labels of 144 prompt batches of this layer (configs[3] accumulation): 0.36 GB of int16 labels
    big = idx[0].repeat(144, 1)
    hist = torch.zeros(E0, dtype=torch.int64, device=dev)

    def hist_fn():
        M.hist_accumulate(big, E0, hist)

    hist_fn()
    torch.cuda.synchronize()
    us_h = time_graph(graph_of(hist_fn, dev), 10) * 1e3
    by_h = big.numel() * 2
    r_hist = dict(bound="hbm", kernel="hist_accumulate_kernel (K4)", shape=f"{big.numel()} labels, E={E0}",
                  achieved=round(by_h / us_h / 1e3, 1), peak=hbm_gbs, unit="GB/s", frac=round(by_h / us_h / 1e3 / hbm_gbs, 4),
                  traffic=None, peak_source=peak_src, us_per_launch=round(us_h, 2),
                  algorithmic=f"2 bytes per (token, slot) label read: {by_h / 1e6:.0f} MB per launch")
    del big
    # the 16 router launches of one batch-16 step (all layer shapes; the small-T layers are launch-latency-bound)
    cells = []
    for (d, h, s) in layer_list():
        es = expert_size_for(args.experts, h)
        E, T = h // es, B * s
        cells.append(dict(T=T, E=E, es=es, k=int(E * RATIO), s=s, scores=torch.randn(T, E, generator=gen, device=dev),
                          H=torch.empty(T, h, dtype=torch.bfloat16, device=dev).normal_(generator=gen),
                          hist=torch.zeros(E, dtype=torch.int64, device=dev)))

    def sweep():
        for c in cells:
            M.router_topk(c["scores"], c["k"], want_bits=False, want_idx=False, hist=c["hist"], H=c["H"],
                          expert_size=c["es"], count_rows=(0, c["s"]))

    sweep()
    torch.cuda.synchronize()
    ms = time_graph(graph_of(sweep, dev), 10)
    by = sum(c["T"] * (4 * c["E"] + 2 * c["es"] * (c["E"] - c["k"])) for c in cells)
    r_router["sweep_16_layers"] = dict(achieved=round(by / (ms * 1e-3) / 1e9, 1), unit="GB/s", ms_per_sweep=round(ms, 4),
                                       frac=round(by / (ms * 1e-3) / 1e9 / hbm_gbs, 4),
                                       algorithmic=f"{by / 1e6:.0f} MB over the 16 launches of one UNet-batch-{B} step")
    return r_router, r_hist


def sampling_leg(M, args, dev, rank, world, dist):
    """BASELINE configs[2] / [3] through the receiver API: RemoveExperts (lists for t < 20) + per-timestep counters,
    8 prompts x 50 steps per CUDA-graph replay; prompt batches b = rank, rank + world, ...; one all-reduce at the end."""
    import numpy as np
    import neuron_receivers as nr
    from moefication import helper
    from moe_b200.sd_modules import FFNStackUNet, SyntheticFFNPipeline, GraphedSampling, sd_ffn_shapes
    steps, n_prompts = args.sampling_steps, args.sampling_prompts
    torch.manual_seed(0)
    unet = FFNStackUNet(latent_hw=64)
    pipe = SyntheticFFNPipeline(unet, num_inference_steps=steps, device=dev)
    shapes = sd_ffn_shapes(64)
    labels = {}
    for i, (n, d, h, s) in enumerate(shapes):
        es = expert_size_for(args.experts, h)
        labels[n + ".proj.weight"] = np.random.RandomState(i).permutation(np.repeat(np.arange(h // es), es))

    class A:
        res_path = ""
        moefication = {"topk_experts": RATIO}
    pipe, names, n_exp = helper.modify_ffn_to_experts(pipe, A(), labels_by_name=labels)
    e_max = max(n_exp.values())
    hist = torch.zeros(steps, 16, e_max, dtype=torch.int64, device=dev)
    with tempfile.TemporaryDirectory() as td:
        rs = np.random.RandomState(2)                       # choice(E, E // 10) for t < 20
        for t in range(steps):
            for l, nm in enumerate(names):
                E = n_exp[nm]
                lst = sorted(int(v) for v in rs.choice(E, E // 10, replace=False)) if t < 20 else []
                json.dump(lst, open(os.path.join(td, f"timestep_{t}_layer_{l}.json"), "w"))
        rec = nr.RemoveExperts(0, td, steps, 16, capture_gates=False, hist=hist, count_rows='row0')
    M.reset_launch_count()
    gs = GraphedSampling(pipe, rec, n_prompts, steps)
    launches_per_run = M.launch_count() // 2                # eager pass + capture pass
    hist.zero_()
    n_batches = args.sampling_batches
    # warm replay
    gs.load_states(10_000 + rank)
    gs.replay()
    torch.cuda.synchronize()
    hist.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in range(n_batches):
        gs.load_states(rank + b * world)                    # prompt batch index = its seed: rank-independent inputs
        gs.replay()
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    hist_host = hist.cpu()                                  # the result a caller reads (freq_expert_select.py:61-72)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms[0])
    tok_row0 = [s for (_, _, _, s) in shapes]
    ks = [int(n_exp[nm] * RATIO) for nm in names]
    want = torch.tensor([[world * n_batches * tok_row0[l] * ks[l] for l in range(16)]] * steps)
    exact = bool((hist_host.sum(-1) == want).all())
    total_steps = world * n_batches * steps
    tokens = 2 * n_prompts * sum(tok_row0)
    return dict(config="BASELINE configs[2]/[3]: 50-step sampling of the 16-FFN stack, 8 prompts per batch (UNet batch 16), "
                       "RemoveExperts lists (10% of the experts, t < 20) + row-0 expert counters per (timestep, layer), "
                       "through neuron_receivers.RemoveExperts hooks; one CUDA graph per sampling run",
                ffn_stack_steps_per_s=total_steps / (ms * 1e-3), prompts_per_s=world * n_batches * n_prompts / (ms * 1e-3),
                tokens_per_s=total_steps * tokens / (ms * 1e-3), ms_per_step=ms / (n_batches * steps),
                prompt_batches_per_rank=n_batches, steps=steps, unet_batch=2 * n_prompts, n_gpus=world,
                fused_launches_per_sampling_run=launches_per_run, histogram_counts_exact=exact,
                timing="CUDA events around load-states + graph replays + all-reduce + D2H of the histogram; max over ranks",
                note="FFN-stack steps (norm3 + FFN + residual of the 16 transformer blocks, stock LayerNorm / add around "
                     "our fused launch); conv / attention layers of the UNet are outside this path")


def gpu_arm(args):
    import moe_b200 as M
    from moe_b200.packing import ExpertLayout, pack_ffn
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)     # before any pinned allocation: host pages land next to the GPU's PCIe root
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M._lib.load()

    B = args.batch
    layers = []
    gen = torch.Generator().manual_seed(1234 + rank)      # each rank = a different prompt shard
    wgen = torch.Generator().manual_seed(0)               # weights are replicated
    e_max = 0
    for li, (d, h, s) in enumerate(layer_list()):
        es = expert_size_for(args.experts, h)
        E = h // es
        e_max = max(e_max, E)
        w1 = ((torch.rand(2 * h, d, generator=wgen) * 2 - 1) / d ** 0.5)
        b1 = ((torch.rand(2 * h, generator=wgen) * 2 - 1) / d ** 0.5)
        w2 = ((torch.rand(d, h, generator=wgen) * 2 - 1) / h ** 0.5)
        b2 = ((torch.rand(d, generator=wgen) * 2 - 1) / h ** 0.5)
        p = pack_ffn(ExpertLayout.contiguous(E, es), w1, b1, w2, b2, device=dev)
        x_host = torch.nn.functional.layer_norm(torch.randn(B * s, d, generator=gen), (d,)).to(torch.bfloat16)
        T = B * s
        layers.append(dict(d=d, h=h, s=s, T=T, E=E, es=es, k=int(E * RATIO), p=p, x_host=x_host,
                           x=x_host.to(dev), H=torch.empty(T, h, dtype=torch.bfloat16, device=dev),
                           scores=torch.empty(T, E, dtype=torch.float32, device=dev),
                           y=torch.empty(T, d, dtype=torch.bfloat16, device=dev)))
    # host side of the e2e leg: one pinned arena per direction, the per-layer host buffers are views into it; on the
    # device two copies of each arena (step parity) so that a step's copy-in never waits for the previous step
    n_elem = sum(L["T"] * L["d"] for L in layers)
    x_arena_host = torch.empty(n_elem, dtype=torch.bfloat16).pin_memory()
    y_arena_host = torch.empty(n_elem, dtype=torch.bfloat16).pin_memory()
    x_arena = [torch.empty(n_elem, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    y_arena = [torch.empty(n_elem, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    off = 0
    for L in layers:
        n = L["T"] * L["d"]
        L["off"], L["n"] = off, n
        x_arena_host[off:off + n].copy_(L["x_host"].reshape(-1))
        L["x_host"] = x_arena_host[off:off + n].view(L["T"], L["d"])
        L["y_host"] = y_arena_host[off:off + n].view(L["T"], L["d"])
        L["x2"] = [a[off:off + n].view(L["T"], L["d"]) for a in x_arena]
        L["y2"] = [a[off:off + n].view(L["T"], L["d"]) for a in y_arena]
        off += n
    hist = torch.zeros(len(layers), e_max, dtype=torch.int64, device=dev)
    hist_host = torch.zeros(len(layers), e_max, dtype=torch.int64).pin_memory()
    tokens_per_step = sum(L["T"] for L in layers)

    fused = args.path == "fused"

    def layer_call(li, x=None, y=None):
        """One hooked FFN layer call on the current stream: fused kernel, or the K1 -> K2 -> K3 triple."""
        L = layers[li]
        p = L["p"]
        x = L["x"] if x is None else x
        y = L["y"] if y is None else y
        if fused:
            M.ffn_fused(x, p.w1p, p.b1p, p.w2p, p.b2, L["E"], L["es"], L["k"], M.ACT_GELU,
                        hist=hist[li, :L["E"]], count_rows=(0, L["s"]), H_out=L["H"], scores_out=L["scores"], out=y)
            return
        M.geglu_up(x, p.w1p, p.b1p, L["E"], L["es"], M.ACT_GELU, out=L["H"], scores_out=L["scores"])
        M.router_topk(L["scores"], L["k"], want_bits=False, hist=hist[li, :L["E"]], H=L["H"], expert_size=L["es"],
                      count_rows=(0, L["s"]))
        M.down_proj(L["H"], p.w2p, p.b2, out=y)

    def ffn_step():
        for li in range(len(layers)):
            layer_call(li)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sets kernel attributes before any capture)
    for _ in range(max(3, args.warmup)):
        ffn_step()
    torch.cuda.synchronize()
    launches_per_step = (1 if fused else 3) * len(layers)

    graph = None if args.no_graph else graph_of(ffn_step, dev)

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            ffn_step()

    # ---- timed region: K steps, device-resident inputs (working set ~0.5 GB >> 126 MB L2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(50):          # keep the GPU under load while the first clock samples are taken
        run_step()
    hist.zero_()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        run_step()
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)       # the path's only exchange: integer histograms
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    counts_ok = bool((hist[:, :].sum(1).cpu() == torch.tensor(
        [world * args.steps * L["s"] * L["k"] for L in layers])).all())

    # ---- per-kernel breakdown: one CUDA graph per kernel type holding that kernel's 16 launches of a step,
    # replayed back to back and timed with CUDA events on the launching stream (a per-launch event pair
    # would time host launch gaps, not the kernels).  Buffers hold valid data from the runs above.
    k1_flops = sum(4.0 * L["d"] * L["h"] * L["T"] for L in layers)
    k3_flops = sum(2.0 * L["d"] * L["h"] * L["T"] for L in layers)
    k3_necessary = sum(2.0 * L["d"] * L["es"] * L["k"] * L["T"] for L in layers)

    def only(kind):
        for li, L in enumerate(layers):
            p = L["p"]
            if kind == 0:
                M.geglu_up(L["x"], p.w1p, p.b1p, L["E"], L["es"], M.ACT_GELU, out=L["H"], scores_out=L["scores"])
            elif kind == 1:
                M.router_topk(L["scores"], L["k"], want_bits=False, hist=hist[li, :L["E"]], H=L["H"],
                              expert_size=L["es"], count_rows=(0, L["s"]))
            else:
                M.down_proj(L["H"], p.w2p, p.b2, out=L["y"])

    per_kernel = []
    n_inst = max(5, min(args.steps, 20))
    for kind in range(0 if fused else 3):
        per_kernel.append(time_graph(graph_of(lambda kind=kind: only(kind), dev), n_inst))   # ms per step in K1 / K2 / K3

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    h2d = sum(L["x_host"].numel() * 2 for L in layers)
    d2h = sum(L["y_host"].numel() * 2 for L in layers) + hist_host.numel() * 8

    # Three streams: copy-in, compute, copy-out.  Each step's layers are cut into --e2e-groups contiguous groups of
    # about equal bytes; a group is copied in with one DMA (host arena slice -> device arena slice), its layers run as
    # soon as it has landed, and its outputs go back with one DMA, so copy-in of group g+1 and copy-out of group g-1
    # overlap the kernels of group g.  Consecutive steps alternate between two device arenas, so the copy-in of step
    # s+1 overlaps the tail of step s; a device arena slice is reused only after the step-before-last's kernels /
    # copy-out that touch it have finished (events below).  Every step copies all inputs in and all results out.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    nL = len(layers)
    n_groups = max(1, min(args.e2e_groups, nL))
    groups, acc, cur = [], 0, []
    for li, L in enumerate(layers):
        cur.append(li)
        acc += L["n"]
        if acc >= n_elem * (len(groups) + 1) / n_groups or li == nL - 1:
            groups.append(cur)
            cur = []
    groups = [g for g in groups if g]

    def e2e_steps(main, n):
        """n steps with host buffers on three streams forked from / joined into `main`."""
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        ev_cmp = {}       # (parity, group) -> kernels of that group done in the last step of that parity
        ev_out = {}       # (parity, group) -> copy-out of that group's outputs done
        for step in range(n):
            par = step & 1
            for gi, g in enumerate(groups):
                lo, hi = layers[g[0]]["off"], layers[g[-1]]["off"] + layers[g[-1]]["n"]
                ev_in = torch.cuda.Event()
                with torch.cuda.stream(s_in):
                    if (par, gi) in ev_cmp:
                        s_in.wait_event(ev_cmp[(par, gi)])
                    x_arena[par][lo:hi].copy_(x_arena_host[lo:hi], non_blocking=True)
                    ev_in.record(s_in)
                main.wait_event(ev_in)
                if (par, gi) in ev_out:
                    main.wait_event(ev_out[(par, gi)])
                for li in g:
                    layer_call(li, layers[li]["x2"][par], layers[li]["y2"][par])
                ev_cmp[(par, gi)] = torch.cuda.Event()
                ev_cmp[(par, gi)].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[(par, gi)])
                    y_arena_host[lo:hi].copy_(y_arena[par][lo:hi], non_blocking=True)
                    ev_out[(par, gi)] = torch.cuda.Event()
                    ev_out[(par, gi)].record(s_out)
            with torch.cuda.stream(s_out):
                hist_host.copy_(hist, non_blocking=True)
        main.wait_stream(s_in)
        main.wait_stream(s_out)       # every step's results (outputs + histogram) are on the host

    n_e2e = min(args.steps, 20)
    e2e_steps(torch.cuda.current_stream(), 3)
    torch.cuda.synchronize()
    e2e_graph = None
    if not args.no_graph:
        # the same n_e2e steps as one CUDA graph (copies included), so that the 16 x (2 copies + launch) per step are
        # not paced by the Python / ctypes host path
        e2e_graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(e2e_graph, stream=side):
                e2e_steps(side, n_e2e)
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            e2e_graph.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if e2e_graph is not None:
        e2e_graph.replay()
    else:
        e2e_steps(torch.cuda.current_stream(), n_e2e)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / n_e2e
    y_check = float(layers[0]["y_host"].float().abs().sum())      # the copied-back result is real data ...
    e2e_same = all(torch.equal(L["y_host"], L["y"].cpu()) for L in layers)   # ... and equals the resident-input run's

    clocks = sampler.stop() if rank == 0 else None   # sampled across the timed region, the breakdown and e2e
    stats = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(stats[0]), float(stats[1])
    ms_step = ms_total / args.steps

    peaks = {}
    peak_src = "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        peaks = {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))

    # ---- the other two legs (every rank takes part: the sampling leg shards prompt batches and all-reduces)
    sampling = None
    if args.sampling_batches > 0:
        sampling = sampling_leg(M, args, dev, rank, world, dist)
    r_router = r_hist = None
    if rank == 0 and not args.no_aux:
        r_router, r_hist = aux_rooflines(M, args, dev, hbm_gbs, peak_src)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    capped = bool(clocks and "sw_power_cap" in (clocks.get("reasons") or []))
    peak_tf = float(peaks.get("bf16_tflops_sustained" if capped else "bf16_tflops", 1590.0))
    peak_kind = "sustained (power cap seen in the sampled clocks)" if capped else "burst (no power cap in the sampled clocks)"
    if fused:
        # the step IS the dominant kernel: 16 launches of ffn_fused_kernel, nothing else in the timed region
        tf = (k1_flops + k3_flops) / (ms_step * 1e-3) / 1e12
        tf_nec = (k1_flops + k3_necessary) / (ms_step * 1e-3) / 1e12
        roofline = dict(bound="tensor", kernel="ffn_fused_kernel (K1 + routing + K3 per layer)", achieved=round(tf, 2),
                        peak=peak_tf, unit="TFLOP/s", frac=round(tf / peak_tf, 4),
                        achieved_necessary=round(tf_nec, 2), frac_necessary=round(tf_nec / peak_tf, 4),
                        # ncu --set full, dram__bytes_read + write of one launch (profiles/r01_ncu_full_fused_layer_d320.csv;
                        # d = 320, 8192 tokens: x 5.2 MB + weights 2.5 MB in, Y still in L2; H never leaves L2)
                        traffic=7.93e6, traffic_launch="ffn_fused_kernel d=320 T=8192", peak_source=peak_src,
                        peak_kind=peak_kind,
                        algorithmic="6*d*h FLOP per token (4dh up-projection + 2dh dense-equivalent down-projection), "
                                    "summed over the step's 16 launches / the step time (CUDA events; the timed region "
                                    "holds nothing but these launches); *_necessary counts the down-projection at "
                                    "2*d*es*k (active experts only)")
    else:
        k1_tf = k1_flops / (per_kernel[0] * 1e-3) / 1e12
        k3_tf = k3_flops / (per_kernel[2] * 1e-3) / 1e12
        roofline = dict(bound="tensor", kernel="geglu_up_kernel (K1)", achieved=round(k1_tf, 2), peak=peak_tf,
                        unit="TFLOP/s", frac=round(k1_tf / peak_tf, 4),
                        traffic=6.93e6, traffic_launch="K1 d=320 T=8192", peak_source=peak_src, peak_kind=peak_kind,
                        algorithmic="4*d*h FLOP per token, summed over the 16 K1 launches of a step / their summed "
                                    "duration (CUDA events around a graph of exactly those launches)",
                        kernel_ms_per_step=dict(K1_geglu_up=round(per_kernel[0], 4), K2_router=round(per_kernel[1], 4),
                                                K3_down_proj=round(per_kernel[2], 4)),
                        K3_down_proj_tflops=round(k3_tf, 2), K3_frac=round(k3_tf / peak_tf, 4),
                        K3_necessary_tflops=round(k3_necessary / (per_kernel[2] * 1e-3) / 1e12, 2))

    line = dict(metric="moe_ffn_tokens_per_s", value=world * tokens_per_step / (ms_step * 1e-3), unit="tokens/s",
                n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms_step, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=workload_config(args, world),
                impl_details=dict(cuda_graph=graph is not None, path=args.path,
                                  counters="row-0 expert histogram fused in the routing stage of the layer kernel"),
                ffn_stack_steps_per_s=world * 1e3 / ms_step,
                e2e=dict(value=world * tokens_per_step / (ms_e2e * 1e-3), unit="tokens/s", h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h, ms_per_step=ms_e2e, cuda_graph=e2e_graph is not None,
                         copies_per_step=2 * len(groups) + 1, pipelined_steps=True, host_binding=numa,
                         h2d_gbs_per_rank=round(h2d / (ms_e2e * 1e-3) / 1e9, 2),
                         d2h_gbs_per_rank=round(d2h / (ms_e2e * 1e-3) / 1e9, 2),
                         output_abs_sum_layer0=y_check, outputs_equal_resident_run=e2e_same),
                gpu_launches=launches_per_step * args.steps, clocks=clocks, roofline=roofline,
                histogram_counts_exact=counts_ok)
    if r_router is not None:
        line["roofline_router"], line["roofline_hist"] = r_router, r_hist
    if sampling is not None:
        line["sampling"] = sampling
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_arm(args, steps=2, warmup=1)
        cb.pop("timed_region_s", None)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cb, tokens = cpu_arm(args, steps=args.steps, warmup=args.warmup)
        ms = cb.pop("ms_per_step")
        wall = cb.pop("timed_region_s")
        line = dict(impl="reference", metric="moe_ffn_tokens_per_s", value=cb["value"], unit="tokens/s",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=workload_config(args, args.gpus),
                    impl_details=dict(note="reference's hook arithmetic on the host CPU, all 16 layers every step (oracle "
                                           "port, bit-identical to the reference classes on CPU; the Python reference needs "
                                           "diffusers and /root/reference, absent from the GPU box)",
                                      timed_region_s=round(wall, 3)),
                    cpu_baseline=dict(cb, value=cb["value"]),
                    e2e=dict(value=cb["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    gpu_arm(args)


if __name__ == "__main__":
    main()
