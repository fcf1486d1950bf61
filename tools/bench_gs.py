#!/usr/bin/env python
"""Config 3 of BASELINE.json: re-iterated Gram-Schmidt of k vectors with a diagonal mass-matrix
product (HBM-bound: one weighted dot + one axpy per pair-step, a host sync after every dot).

Input family of SURVEY.md §8d: a_i = g_i + 2 s  =>  every vector i >= 1 re-iterates exactly once,
i.e. k(k-1) pair-steps, which makes the byte model exact:
  per pair-step 24 n (weighted dot: a_j, a_i, w) + 24 n (axpy: read a_j, a_i, write a_i)
  per vector    3 weighted norms (16 n each) + 1 scal (16 n);  plus the final check Gramian.
Prints one JSON line (not the driver's bench contract; results are quoted in DESIGN.md).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dofs', dest='n', type=int, default=10_000_000)
    ap.add_argument('--k', type=int, default=256)
    ap.add_argument('--cpu-n', type=int, default=0, help='also time the oracle on this many rows')
    ap.add_argument('--variant', default='mgs', choices=['mgs', 'cgs2', 'bcgs2'],
                    help="cgs2: block projection per pass (streaming inner + block lincomb), two host syncs per pass; "
                         "bcgs2: panels of --panel vectors against all previous ones on the tensor cores, twice")
    ap.add_argument('--panel', default='32', help='panel width(s) of bcgs2, outermost first, e.g. 128,16')
    ap.add_argument('--sweep', action='store_true',
                    help='one fused device sweep per pass (pmc_mgs_sweep, in-kernel sums over the GPUs) instead of one '
                         'launch + host sync per pair')
    args = ap.parse_args()
    import torch

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorArrayImpl, CudaVectorSpace
    n, k = args.n, args.k
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    comm = None
    if world > 1:                                   # torchrun: rows sharded over ranks (strong scaling)
        import torch.distributed as dist

        from pymor_b200.dist import ProcessGroupComm
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = ProcessGroupComm()
    space = CudaVectorSpace(n, device=torch.device('cuda', torch.cuda.current_device()), comm=comm)
    A = space.random_device(k, seed=1, scale=1.0 / np.sqrt(n))
    s = space.random_device(1, seed=2, scale=1.0 / np.sqrt(n))
    A.axpy(2.0, s)
    gen = torch.Generator(device=space.device).manual_seed(3 + rank)
    w = torch.rand(space.local_dim, generator=gen, device=space.device, dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, space)

    counts = {'pairwise_inner': 0, 'axpy': 0, 'scal': 0, 'mgs_sweep': 0, 'inner': 0, 'lincomb': 0}
    cgs_cols = [0]
    sweep_pairs = [0]
    for name in counts:
        orig = getattr(CudaVectorArrayImpl, name)

        def wrapped(self, *a, _orig=orig, _name=name, **kw):
            counts[_name] += 1
            if _name == 'mgs_sweep':
                sweep_pairs[0] += a[0]
            if _name == 'lincomb':
                cgs_cols[0] += len(a[0])
            return _orig(self, *a, **kw)
        setattr(CudaVectorArrayImpl, name, wrapped)

    # warm up the kernels on a small slice
    def run(arr, **kw):
        if args.variant == 'bcgs2':
            return alg.block_gram_schmidt(arr, product=prod, panel=tuple(int(b) for b in args.panel.split(',')),
                                          copy=False, **kw)
        return alg.gram_schmidt(arr, product=prod, copy=False, sweep=args.sweep, variant=args.variant, **kw)

    run(A[:8].copy())
    for c in counts:
        counts[c] = 0
    sweep_pairs[0] = 0
    cgs_cols[0] = 0
    torch.cuda.synchronize()
    l0 = space.ctx.launch_count
    t0 = time.perf_counter()
    Q, R = run(A, return_R=True, check=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = space.ctx.launch_count - l0
    if args.variant == 'bcgs2':  # tensor-bound: report seconds; the MGS byte model does not apply
        pair_steps = k * (k - 1)
        norms = counts['pairwise_inner']
    elif args.variant == 'cgs2': # same byte model: projecting onto c columns = c weighted dots + c axpys' worth of traffic
        pair_steps = cgs_cols[0]
        norms = counts['pairwise_inner']
    elif args.sweep:         # same byte model: every swept pair is a weighted dot + an axpy, every sweep ends in a norm
        pair_steps = sweep_pairs[0]
        norms = counts['pairwise_inner'] + counts['mgs_sweep']
    else:
        pair_steps = counts['axpy']
        norms = counts['pairwise_inner'] - pair_steps
    nbytes = pair_steps * 48.0 * n + norms * 16.0 * n + counts['scal'] * 16.0 * n
    err = float(np.max(np.abs(Q.gramian(prod) - np.eye(len(Q)))))
    peaks = ROOT / 'MEASURED_PEAKS.json'
    hbm = json.loads(peaks.read_text())['hbm_gbs'] if peaks.exists() else 6650.0
    out = {'workload': f'gram_schmidt {n} x {k}, diag-W, re-iterated', 'fused_sweep': bool(args.sweep), 'variant': args.variant, 'panel': args.panel if args.variant == 'bcgs2' else None, 'n_gpus': world, 'seconds': dt, 'pair_steps': pair_steps,
           'weighted_norms': norms, 'scals': counts['scal'], 'algorithmic_bytes': nbytes,
           'achieved_gbs': nbytes / dt * 1e-9, 'hbm_peak_gbs': hbm,
           'peak_source': 'MEASURED_PEAKS.json' if peaks.exists() else 'fallback 6650 GB/s',
           'frac': nbytes / dt * 1e-9 / (hbm * world), 'us_per_pair_step': dt / max(pair_steps, 1) * 1e6,
           'gpu_launches': launches, 'orth_err': err, 'len_Q': len(Q)}
    if args.cpu_n:
        from oracle import numpy_path as O
        rng = np.random.default_rng(1)
        Ah = rng.standard_normal((args.cpu_n, k)) / np.sqrt(args.cpu_n) + 2 * rng.standard_normal((args.cpu_n, 1)) / np.sqrt(args.cpu_n)
        wh = rng.uniform(0.5, 1.5, args.cpu_n)
        st = {}
        t0 = time.perf_counter()
        O.gram_schmidt(Ah, wh, check=False, stats=st)
        tc = time.perf_counter() - t0
        out['cpu_baseline'] = {'rows': args.cpu_n, 'seconds': tc, 'pair_steps': st['pair_steps'], 'kind': 'port',
                               'cores': os.cpu_count(), 'extrapolated_seconds_full_n': tc * n / args.cpu_n}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
