"""bench.py contract pieces that run without a GPU: the reference arm (pyMOR's own NumPy backend through the same
pod() call, on the host cores) prints one well-formed JSON line whose `config` equals the CUDA arm's; the flop model
matches SURVEY.md §8d."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--k', '40', '--dofs', '100000'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'TFLOP/s' and line['dtype'] == 'f64'
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['vs_baseline'] is None
    # same metric / config object as the CUDA arm would print for these arguments
    sys.path.insert(0, str(ROOT))
    import argparse

    import bench
    args = argparse.Namespace(n=100000, k=40, scaling='strong')
    assert line['config'] == bench.workload_config(args, 1) and line['metric'] == bench.METRIC
    assert line['scaling'] == 'strong'


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work."""
    import os
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_flop_model():
    sys.path.insert(0, str(ROOT))
    import bench
    n, k = 10_000_000, 1000
    assert bench.pod_flops(n, k, k) == n * k * (k + 1) + 2 * n * k * k + n * k * (k + 1)
    assert abs(bench.pod_flops(n, k, k) / 4.0e13 - 1) < 0.01          # ~4.0e13 flop


def test_ctypes_tables_match_the_header():
    """Every prototype of include/pymor_b200.h is bound in pymor_b200/_lib.py with the same number of arguments and
    matching pointer / integer kinds (an ABI drift between the header and the ctypes tables would corrupt the call
    stack silently)."""
    import ctypes
    import re

    from pymor_b200 import _lib
    text = (ROOT / 'include' / 'pymor_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = re.findall(r'^(?:int|size_t|int64_t|const char \*)\s*(pmc_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.M | re.S)
    assert len(protos) == len(_lib.EXPORTED_SYMBOLS)

    def kind(decl):
        decl = decl.strip()
        if decl == 'void' or not decl:
            return None
        if '*' in decl:
            return 'ptr'
        if decl.startswith('double'):
            return 'double'
        return 'int'

    def ckind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int))) and \
                issubclass(t, ctypes._Pointer):
            return 'ptr'
        if t is ctypes.c_double:
            return 'double'
        return 'int'

    for name, args in protos:
        want = [k for k in (kind(a) for a in args.split(',')) if k]
        if name in _lib._CTX_FUNCS:
            have = ['ptr'] + [ckind(t) for t in _lib._CTX_FUNCS[name]]
        else:
            have = [ckind(t) for t in _lib._OTHER_FUNCS[name][0]]
        assert have == want, (name, have, want)


def test_parity_block_logic_on_the_emulated_abi():
    """bench.py's `parity` block (row sample against the unmodified reference on the same bits + size-independent
    properties) on the NumPy-emulated C ABI: the block itself must report agreement to rounding."""
    import argparse

    import numpy as np
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / 'tests'))
    from emulated_lib import EmulatedLib

    from pymor_b200 import _lib
    _lib.set_library_for_testing(EmulatedLib())
    try:
        import bench
        import torch
        from pymor.algorithms.pod import pod

        from pymor_b200.operators import CudaDiagonalOperator
        from pymor_b200.vectorarray import CudaVectorSpace
        bench.quiet_pymor()
        n, k = 30000, 40
        rng = np.random.default_rng(1)
        Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
        A = (rng.standard_normal((n, k)) / np.sqrt(n)) @ (np.diag(bench.spectrum(k)) @ Qk.T)
        space = CudaVectorSpace(n, device='cpu')
        VA = space.from_numpy(A)
        w = torch.from_numpy(rng.uniform(0.5, 1.5, n))
        prod = CudaDiagonalOperator(w, space)
        out = bench.parity_block(argparse.Namespace(k=k), space, VA, w, prod, None, 'cpu',
                                 lambda arr, pr, **kw: pod(arr, product=pr, method='method_of_snapshots', **kw))
    finally:
        _lib.set_library_for_testing(None)
    smp, full = out['sample'], out['full_size']
    assert smp['ok'] and smp['rows'] == 20000 and smp['modes_kept'] == [k, k]
    assert smp['singular_values_max_rel_err'] < 1e-12 and smp['subspace_angle_rad'] < 1e-10
    assert full['orth_err_max_abs(U^T W U - I)'] < 1e-11 and full['gram_bitwise_symmetric']
    assert full['trace_identity_rel_err(sum s^2 vs sum ||a_j||_W^2)'] < 1e-13 and full['singular_values_sorted']
