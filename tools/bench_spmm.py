#!/usr/bin/env python
"""SpMM timing: Y = M X for the 5-point stencil CSR operator (Operator.apply)."""
import argparse, json, os, sys, time
from pathlib import Path
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from bench_project import stencil


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=2000)
    ap.add_argument('--ny', type=int, default=1000)
    ap.add_argument('--ks', default='1,8,64,256')
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    import torch
    from pymor_b200.operators import CudaCsrOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    n = args.nx * args.ny
    M = stencil(args.nx, args.ny)
    sp = CudaVectorSpace(n)
    op = CudaCsrOperator(M, sp)
    for k in (int(v) for v in args.ks.split(',')):
        V = sp.random_device(k, seed=1)
        t_end = time.perf_counter() + 0.3                 # bring the clocks up: short runs start from idle clocks
        while time.perf_counter() < t_end:
            Y = op.apply(V)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            Y = op.apply(V)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(json.dumps({'n': n, 'k': k, 'nnz': int(M.nnz), 'ms': ms, 'variant': os.environ.get('PMC_CSR_VARIANT', '0'),
                          'algorithmic_gbs': (12.0 * M.nnz + 16.0 * n * k) / ms * 1e-6}))
        if k == 8:
            ref = M @ V[:2].to_numpy()
            assert np.max(np.abs(Y[:2].to_numpy() - ref)) < 1e-11
        del V, Y


if __name__ == '__main__':
    main()
