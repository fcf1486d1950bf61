"""world_size-2 (and 3) CPU test of the row-sharded path: gloo backend, one process per rank."""
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_spmd_gloo(world):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()), str(ROOT / 'tests' / 'sharded_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count(': ok') == world
