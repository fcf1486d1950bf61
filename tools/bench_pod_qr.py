#!/usr/bin/env python
"""POD through the reference's DEFAULT method `qr_svd` (src/pymor/algorithms/pod.py:17, svd_va.py:103-168) with
the QR step done by the block Gram-Schmidt on the tensor cores (`gs_variant='bcgs2'`), next to the method of
snapshots that bench.py measures: same synthetic snapshots (sigma_j = 10^(-2j/(k-1))), diagonal mass-matrix
product.  qr_svd does not square the condition number; with k(k-1) host-synchronised pair-steps its Gram-Schmidt
is out of reach at this size for the reference's algorithm (10M x 1000: ~10^6 pair-steps x 74 us).
Prints one JSON line: seconds of both methods, relative difference of the singular values, orthogonality."""
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
    ap.add_argument('--dofs', dest='n', type=int, default=4_000_000)
    ap.add_argument('--k', type=int, default=500)
    ap.add_argument('--panel', default='128,16')
    ap.add_argument('--reps', type=int, default=2)
    args = ap.parse_args()
    import torch

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    n, k = args.n, args.k
    space = CudaVectorSpace(n)
    rng = np.random.default_rng(42)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    mix = np.diag(10.0 ** (-2 * np.arange(k) / max(k - 1, 1))) @ Qk.T
    Gn = space.random_device(k, seed=42, distribution='normal', scale=1.0 / np.sqrt(n))
    A = Gn.lincomb(mix)
    del Gn
    w = torch.rand(n, device=space.device, dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, space)
    panel = tuple(int(b) for b in args.panel.split(','))
    orig = alg.block_gram_schmidt

    def paneled(*a, **kw):
        kw.setdefault('panel', panel)
        return orig(*a, **kw)
    alg.block_gram_schmidt = paneled
    res = {}
    for name, kw in (('method_of_snapshots', dict(method='method_of_snapshots')),
                     ('qr_svd_bcgs2', dict(method='qr_svd', gs_variant='bcgs2'))):
        times = []
        for _ in range(args.reps + 1):
            torch.cuda.synchronize()
            l0, t0 = space.ctx.launch_count, time.perf_counter()
            U, s = alg.pod(A, product=prod, **kw)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            launches = space.ctx.launch_count - l0
            orth = float(np.max(np.abs(U.gramian(prod) - np.eye(len(U)))))
            m = len(U)
            del U
        res[name] = {'seconds': min(times[1:]), 'modes': m, 'gpu_launches': launches, 'orth_err': orth, 's': s}
    s0, s1 = res['method_of_snapshots'].pop('s'), res['qr_svd_bcgs2'].pop('s')
    mm = min(len(s0), len(s1))
    out = {'workload': f'POD {n} x {k}, diag-W, sigma_j = 10^(-2j/(k-1))', 'panel': list(panel), **res,
           'svals_rel_diff': float(np.max(np.abs(s0[:mm] - s1[:mm]) / s0[:mm]))}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
