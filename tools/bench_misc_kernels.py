#!/usr/bin/env python
"""One launch each of the remaining HBM-bound kernels at n = 10M DOF (for ncu captures): CSR SpMM (5-point stencil,
64 columns), the fused CSR pairwise two-form, amax, the fancy-index gather and the in-place column shift of `del`."""
import os
import sys
from pathlib import Path

os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tools'))


def main():
    import torch
    from bench_project import stencil_rows

    from pymor_b200.operators import CudaCsrOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    nx, ny, k = 5000, 2000, 64
    n = nx * ny
    sp = CudaVectorSpace(n)
    op = CudaCsrOperator.from_local_rows(stencil_rows(nx, ny, 0, n), sp, symmetric=True)
    V = sp.random_device(k, seed=1, scale=1.0 / np.sqrt(n))
    for _ in range(2):
        Y = op.apply(V)                       # csr_apply_kernel
        p = op.pairwise_apply2(V, V)          # csr_pairwise_kernel
        idx, val = V.amax()                   # amax_kernel
        G = V[[5, 1, 1, 60, 33, 2, 7, 9]].copy()   # gather_cols_kernel
        W = V.copy()
        del W[3]                              # shift_cols_kernel
        torch.cuda.synchronize()
    print('ok', float(p[0]), int(idx[0]), len(G), len(W))


if __name__ == '__main__':
    main()
