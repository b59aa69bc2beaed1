"""Hardware multi-GPU correctness (needs >= 2 GPUs on the box; skipped otherwise): tools/check_ddp.py under torchrun —
the fused symmetric-memory update against the NCCL all-reduce form, N-rank gradients against one rank on the
concatenated batch (SURVEY §4 tier iv), rank bit-identity over real CUDA-graph steps, sharded-state gathering."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_update_and_gradient_equivalence(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tools", "check_ddp.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"OK world={world}" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
    print(r.stdout[-600:])
