"""Row-sharded path on real GPUs: one process per GPU, NCCL all-reduce of the partials."""
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_sharded_spmd_nccl():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs at least 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()),
           str(ROOT / 'tests' / 'sharded_worker.py'), '--real']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count(': ok') == world
