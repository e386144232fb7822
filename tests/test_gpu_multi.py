"""x-slab decomposition over 2 GPUs (NCCL halos) against the single-GPU run of the same global system.
Needs two visible GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["lj", "adress", "adress-cuts"])
def test_two_gpu_slabs_match_single_gpu(mode):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py"), "40", mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-2000:])
    out = json.loads(lines[-1])
    assert out["ok"], out


def test_two_gpu_tetramer_slabs_match_single_gpu():
    """configs[3] physics over x-slabs: whole molecules migrate and travel in the halo, SHAKE / RATTLE stay rank-local"""
    test_two_gpu_slabs_match_single_gpu("tetramer")
