#!/usr/bin/env python
"""Kernel-level timings (CUDA events, device-resident operands and results, L2 far smaller than the
operands): pmc_gram (SYRK / GEMM, weighted or not) and pmc_lincomb at the benchmark shapes.
Prints one JSON line per case: milliseconds, algorithmic TFLOP/s, fraction of the in-run DMMA peak."""
import argparse
import json
import os
import sys
from pathlib import Path

os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pymor_b200._lib import PMC_SYMMETRIC  # noqa: E402
from pymor_b200.vectorarray import CudaVectorSpace  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shapes', default='2000000x500,10000000x1000,4000000x256,1000000x2000')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--opts', default=os.environ.get('PMC_GRAM_OPTS', '0'),
                    help='comma-separated gram_opts values to A/B in one process (csrc/gram2_plan.h; 16 = gram2 kernel)')
    ap.add_argument('--no-lincomb', action='store_true')
    args = ap.parse_args()
    peak = None
    for shape in args.shapes.split(','):
        n, k = (int(v) for v in shape.split('x'))
        sp = CudaVectorSpace(n)
        ctx = sp.ctx
        if peak is None:
            peak = ctx.measure_fp64_peak(True, 200.0)
            print(json.dumps({'fp64_peaks_tflops': {'dmma': peak, 'dfma': ctx.measure_fp64_peak(0, 100.0),
                                                    'mixed_by_warp': ctx.measure_fp64_peak(2, 100.0),
                                                    'mixed_interleaved': ctx.measure_fp64_peak(3, 100.0)}}), flush=True)
        A = sp.random_device(k, seed=1, scale=1.0 / np.sqrt(n))
        w = torch.rand(n, device='cuda', dtype=torch.float64) + 0.5
        out = torch.empty(k * k, device='cuda', dtype=torch.float64)
        impl = A.impl
        pa, cs = impl._ptr(0), impl.ld
        for opts in (int(v) for v in args.opts.split(',')):
            ctx.set_option('gram_opts', opts)
            res = {'shape': shape, 'dmma_peak_tflops': peak, 'gram_opts': opts}
            for name, wp, flags, flop in (('syrk', 0, PMC_SYMMETRIC, n * k * (k + 1.0)),
                                          ('syrk_w', w.data_ptr(), PMC_SYMMETRIC, n * k * (k + 1.0)),
                                          ('gemm_w', w.data_ptr(), 0, 2.0 * n * k * k)):
                ms = timed(lambda: ctx.call('pmc_gram', pa, cs, k, pa, cs, k, n, wp, out.data_ptr(), k, flags), args.reps)
                res[name] = {'ms': ms, 'tflops': flop / ms * 1e-9, 'frac': flop / ms * 1e-9 / peak}
            print(json.dumps(res), flush=True)
        ctx.set_option('gram_opts', 0)
        res = {'shape': shape, 'dmma_peak_tflops': peak}
        if 2 * n * k * 8 < 150e9 and not args.no_lincomb:
            C = torch.randn((k, k), device='cuda', dtype=torch.float64)
            R = sp.zeros(k)
            ms = timed(lambda: ctx.call('pmc_lincomb', pa, cs, k, n, C.data_ptr(), k, k, R.impl._ptr(0), R.impl.ld, 0),
                       args.reps)
            res['lincomb'] = {'ms': ms, 'tflops': 2.0 * n * k * k / ms * 1e-9, 'frac': 2.0 * n * k * k / ms * 1e-9 / peak}
            del R, C
            print(json.dumps(res), flush=True)
        del A, impl, w, out
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
