"""GPU, >= 2 devices (gpurun --gpus 2): the data-parallel exchange step on real NCCL.  One process per GPU under torchrun
(tools/check_grad_sync.py): ranks start from DIFFERENT parameters, attach_grad_sync broadcasts rank 0's, every rank runs its shard
of clips through the CUDA module, and the ONE fused flat gradient exchange inside backward — the NCCL all-reduce, the library's
peer-memory kernel with the NVSwitch multicast load and with plain peer loads, two steps each — must reproduce the gradients
of a single process on the concatenated batch (reference mechanism: DDP, slowfast/models/build.py:79-83).
Skipped on a 1-GPU box; the host-side logic is covered on CPU by tests/test_distributed_cpu.py (gloo, world size 2)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests._util import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_flat_allreduce_equals_single_process_gradients(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_grad_sync.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "nccl [GradSync] vs single process" in r.stdout
    assert r.stdout.count("max-normalised gradient difference") == 3 or "peer-memory sync unavailable" in r.stdout
