"""GPU, >= 2 devices: the all-reduced expert counters of a prompt-sharded sampling run equal the single-GPU run's bit
for bit (NCCL, torchrun subprocess; CUDA path through the receiver hooks).  Skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sampling_histogram_is_bit_exact_over_nccl(lib, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "dist_sampling_check.py"), "--batches", str(2 * world + 1), "--steps", "5", "--latent", "32"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "BITEXACT OK" in res.stdout, res.stdout[-2000:]


def test_single_process_sampling_check_runs(lib):
    """The same script on one GPU (world 1): exercises the graph-captured receiver path in the one-GPU test tier."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dist_sampling_check.py"), "--batches", "2", "--steps", "4",
                          "--latent", "16"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "BITEXACT OK world=1" in res.stdout
