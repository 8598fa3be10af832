"""Cost of `capture_gates` (the reference's per-layer-call `self.gates.append(gate.detach().cpu())`, moefy.py:25) on
the hooked FFN stack at UNet batch 2: MOEFy.observe_activation over N steps with and without gate capture.
Wall-clock per step (the hook path is eager Python here; the copies are asynchronous into pinned memory)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-moe_b200"))
import neuron_receivers as nr
from moefication import helper
from moe_b200.sd_modules import FFNStackUNet, SyntheticFFNPipeline, sd_ffn_shapes
import moe_b200 as M

dev = "cuda:0"
STEPS = int(os.environ.get("STEPS", "10"))
torch.manual_seed(0)
pipe = SyntheticFFNPipeline(FFNStackUNet(64), num_inference_steps=STEPS, device=dev)
labels = {n + ".proj.weight": np.random.RandomState(i).permutation(np.repeat(np.arange(h // 20), 20))
          for i, (n, d, h, s) in enumerate(sd_ffn_shapes(64))}


class A:
    res_path = ""
    moefication = {"topk_experts": 0.3}


pipe, names, n_exp = helper.modify_ffn_to_experts(pipe, A(), labels_by_name=labels)
for capture in (False, True):
    rec = nr.MOEFy(0, capture_gates=capture)
    rec.observe_activation(pipe, "warm-up")
    rec.gates = []
    torch.cuda.synchronize()
    M.reset_launch_count()
    t0 = time.perf_counter()
    out, gates = rec.observe_activation(pipe, "a photo of a cat")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nbytes = sum(g.numel() * g.element_size() for g in gates)
    print(f"capture_gates={capture}: {dt / STEPS * 1e3:7.2f} ms per step ({STEPS} steps, {M.launch_count() // STEPS} launches of ours per step, "
          f"{len(gates)} gates = {nbytes / 1e6:.0f} MB copied to pinned host memory)", flush=True)
    rec.gates = []
