"""Data-parallel equivalence on the CUDA path (MirroredStrategy of src/models/Unets.py:70-75): two ranks,
one process per GPU over NCCL, against oracle.data_parallel_grads.  Needs >= 2 GPUs on the box (gpurun --gpus 2);
skipped otherwise.  The assertions live in tests/dp_worker.py (every rank checks its own side)."""
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
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_two_rank_device_path_matches_oracle(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dp_worker.py'), precision]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert 'dp_worker ok' in r.stdout
