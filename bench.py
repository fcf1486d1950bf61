#!/usr/bin/env python
"""Benchmark of the hot path: POD (method of snapshots, diagonal mass-matrix product) of a
device-resident fp64 snapshot matrix -- BASELINE.json's metric, config "POD of 10M DOF x 1000 fp64
snapshots with W-weighted gramian, DOF-sharded 1/2/4/8 GPUs".

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One step = one full `pod(A, product=W, method='method_of_snapshots')` call: weighted Gramian (SYRK,
TMA + FP64 tensor-core kernel) -> NCCL all-reduce of the k x k partials (N > 1) -> host `eigh` ->
block lincomb `A (V / s)` -> orthonormality check (second weighted Gramian).  Nothing is skipped and
nothing is cached between steps.  `value` = algorithmic flop / wall time of the whole call (host
eigensolve included); the per-phase split is reported next to it.

Scaling is WEAK by default: every rank holds a `--dofs` x `--k` row shard (global DOFs = N * dofs), the
work per GPU is fixed, the only exchange is the all-reduce of the k x k Gram partials.
`--scaling strong` shards a fixed `--dofs` x `--k` problem instead (BASELINE.json config 4 as written); there
the host eigensolve, which does not shrink with N, bounds the speed-up (Amdahl).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

# one 74.5 GiB panel next to another: let the caching allocator remap instead of fragmenting
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
# torchrun pins every rank to ONE host thread by default; the k x k eigensolve stays on the host, so give
# each rank its share of the cores instead (must happen before NumPy/OpenBLAS is loaded)
if int(os.environ.get('WORLD_SIZE', '1')) > 1 and os.environ.get('OMP_NUM_THREADS', '1') == '1':
    _share = max(1, (os.cpu_count() or 1) // int(os.environ.get('LOCAL_WORLD_SIZE', os.environ['WORLD_SIZE'])))
    os.environ['OMP_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = str(_share)

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'POD (method_of_snapshots, diag-W product) FP64 TFLOP/s, 10M DOF x 1000 snapshots per GPU'


# dram__bytes_read.sum + dram__bytes_write.sum of ONE lincomb_dmma_kernel launch, from an `ncu --set full`
# capture of this very workload (n per GPU, k, m); equals the algorithmic 8*n*(k+m) = 160.0 GB to 0.1 %.
NCU_TRAFFIC = {(10_000_000, 1000, 1000): 80.138563e9 + 79.989229e9,
               (2_000_000, 500, 500): 8.006061e9 + 7.963813e9}
NCU_TRAFFIC_SOURCE = 'profiles/r01_ncu_full_lincomb_10Mx1000_details.txt, profiles/r01b_ncu_full_details.txt'


def pod_flops(n, k, m):
    """Algorithmic flop of POD by the method of snapshots: SYRK Gramian + lincomb +
    orthonormality-check Gramian."""
    return n * k * (k + 1) + 2 * n * k * m + n * m * (m + 1)


def spectrum(k):
    return 10.0 ** (-2 * np.arange(k) / max(k - 1, 1))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (NumPy/OpenBLAS/LAPACK port of the reference path)
# ------------------------------------------------------------------------------------------------
def cpu_pod_sample(n, k, seed=42):
    """One POD of an n x k sample of the workload on the host cores.  Returns seconds, modes kept."""
    from oracle import numpy_path as O
    rng = np.random.default_rng(seed)
    Gn = rng.standard_normal((n, k)) / np.sqrt(n)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    A = np.asfortranarray(Gn @ (np.diag(spectrum(k)) @ Qk.T))
    w = rng.uniform(0.5, 1.5, n)
    del Gn
    t0 = time.perf_counter()
    U, s = O.pod(A, w, method='method_of_snapshots')
    dt = time.perf_counter() - t0
    return dt, U.shape[1]


def psutil_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 64 << 30


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        blas = [p for p in threadpool_info() if p.get('user_api') == 'blas']
        if blas:
            return max(p['num_threads'] for p in blas)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    k = args.k
    budget = 150.0                                    # seconds for all steps
    t_cal, _ = cpu_pod_sample(20000, k)               # calibration (also warms the BLAS threads)
    n_s = int(20000 * max(budget / ((args.steps + args.warmup) * t_cal), 0.25))
    n_s = max(5000, min(n_s, 400000))
    for _ in range(args.warmup):
        cpu_pod_sample(n_s, k)
    times, m = [], k
    for _ in range(args.steps):
        dt, m = cpu_pod_sample(n_s, k)
        times.append(dt)
    t = float(np.mean(times))
    val = pod_flops(n_s, k, m) / t * 1e-12
    cores = host_threads()
    sample = f'oracle pod() on a {n_s} x {k} row sample of the workload (diag-W), mean of {args.steps}'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'TFLOP/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'POD method_of_snapshots, diag-W product, {n_s} x {k} sample of '
                               f'{args.n} x {k} per GPU', 'modes_kept': int(m)},
        'cpu_baseline': {'value': val, 'unit': 'TFLOP/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'TFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist

    from pymor_b200 import _lib, algorithms as alg
    from pymor_b200.dist import ProcessGroupComm
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorArrayImpl, CudaVectorSpace

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        comm = ProcessGroupComm()
    k = args.k
    # weak (default): --dofs rows per GPU; strong: --dofs rows in total, split evenly (BASELINE config 4 as written)
    n_local = args.n if args.scaling == 'weak' else (args.n // world // 16) * 16
    n_global = n_local * world
    space = CudaVectorSpace(n_global, device=device, comm=comm)
    assert space.local_dim == n_local
    ctx = space.ctx

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    # ---- synthetic snapshots  A = Gn diag(sigma) Q^T  generated on the device (seed 42) ----------
    rng = np.random.default_rng(42)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    mix = np.diag(spectrum(k)) @ Qk.T
    Gn = space.random_device(k, seed=42, distribution='normal', scale=1.0 / np.sqrt(n_global))
    A = Gn.lincomb(mix)
    del Gn
    gen = torch.Generator(device=device).manual_seed(4242 + rank)
    w = torch.rand(n_local, generator=gen, device=device, dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, space)
    torch.cuda.synchronize(device)
    torch.cuda.empty_cache()

    # ---- phase instrumentation (events on the launching stream; host phases by perf_counter) ------
    phases = {'gramian': [], 'lincomb': [], 'eigh_host': []}
    events = []
    orig_inner, orig_lincomb = CudaVectorArrayImpl.inner, CudaVectorArrayImpl.lincomb

    def timed_inner(self, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig_inner(self, *a, **kw); e1.record()
        events.append(('gramian', e0, e1))
        return out

    def timed_lincomb(self, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig_lincomb(self, *a, **kw); e1.record()
        events.append(('lincomb', e0, e1))
        return out

    import scipy.linalg as spla
    orig_eigh = spla.eigh

    def timed_eigh(*a, **kw):
        t0 = time.perf_counter(); out = orig_eigh(*a, **kw); phases['eigh_host'].append((time.perf_counter() - t0) * 1e3)
        return out

    def pod_step(arr):
        U, s = alg.pod(arr, product=prod, method='method_of_snapshots')
        return U, s

    def timed_loop(step_fn, steps):
        times = []
        for _ in range(steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m = step_fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        return times, m

    def resident_step():
        U, s = pod_step(A)
        m = len(U)
        del U
        return m

    for _ in range(args.warmup):
        resident_step()
    CudaVectorArrayImpl.inner, CudaVectorArrayImpl.lincomb, spla.eigh = timed_inner, timed_lincomb, timed_eigh
    alg.spla.eigh = timed_eigh
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    times, m = timed_loop(resident_step, args.steps)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    CudaVectorArrayImpl.inner, CudaVectorArrayImpl.lincomb = orig_inner, orig_lincomb
    alg.spla.eigh = spla.eigh = orig_eigh
    torch.cuda.synchronize(device)
    for name, e0, e1 in events:
        phases[name].append(e0.elapsed_time(e1))
    t_step = float(np.mean(times))                       # ms, max over ranks per step
    flop_global = pod_flops(n_global, k, m)
    value = flop_global / (t_step * 1e-3) * 1e-12
    eigh_ms = float(np.mean(phases['eigh_host'])) if phases['eigh_host'] else 0.0

    # ---- roofline of the dominant kernel (lincomb_dmma_kernel: 2 n k m flop in ONE launch) --------
    peak_dmma = ctx.measure_fp64_peak(True, 300.0)
    peak_dfma = ctx.measure_fp64_peak(False, 100.0)
    lin_ms = float(np.mean(phases['lincomb'])) if phases['lincomb'] else float('nan')
    lin_flop = 2.0 * n_local * k * m
    achieved = lin_flop / (lin_ms * 1e-3) * 1e-12
    gram_ms = float(np.mean(phases['gramian'])) if phases['gramian'] else float('nan')
    gram_tf = n_local * k * (k + 1) / (gram_ms * 1e-3) * 1e-12   # both Gramians here are k x k SYRK (m == k) or m x m
    roofline = {'bound': 'tensor', 'kernel': 'lincomb_dmma_kernel', 'achieved': achieved, 'peak': peak_dmma,
                'unit': 'TFLOP/s', 'frac': achieved / peak_dmma if peak_dmma else None,
                'traffic': NCU_TRAFFIC.get((n_local, k, int(m))),
                'traffic_source': NCU_TRAFFIC_SOURCE if (n_local, k, int(m)) in NCU_TRAFFIC else None,
                'algorithmic_bytes_per_launch': 8.0 * n_local * (k + m),
                'peak_source': 'FP64 tensor-core (DMMA.8x8x4) issue-rate microbenchmark measured in this run '
                               '(pmc_measure_fp64_peak); MEASURED_PEAKS.json carries no fp64 figure',
                'dfma_peak_tflops': peak_dfma,
                'algorithmic_flop_per_launch': lin_flop, 'launch_ms': lin_ms,
                'gram_dmma_kernel': {'achieved': gram_tf, 'frac': gram_tf / peak_dmma if peak_dmma else None,
                                     'launch_ms': gram_ms}}

    # ---- end to end: host snapshots (pinned) -> H2D -> pod -> D2H of the singular values -----------
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError('skipped (--no-e2e)')
        # The snapshots start in pinned host memory and arrive in column chunks, the way pyMOR's
        # reduce_pod builds its array (`snapshots.append(fom.solve(mu))`).  One rank's shard is 80 GB;
        # with N ranks on one host N x 80 GB do not fit into RAM, so the pinned buffer holds as many
        # snapshot columns as a 45 % share of the free host memory allows and is re-used for the
        # later chunks (every byte is still copied host -> device inside the timed region).
        avail = psutil_available()
        k_host = int(min(k, max(1, (0.45 * avail / world) // (n_local * 8))))
        host = None
        while host is None:                                # pinning may be limited below the free RAM
            try:
                host = torch.empty((k_host, n_local), dtype=torch.float64, pin_memory=True)
            except RuntimeError:
                if k_host <= 8:
                    raise
                k_host = max(8, k_host // 2)
        host.copy_(A.to_torch()[:k_host])
        del A
        torch.cuda.empty_cache()
        host_np = host.numpy().T                          # (n_local, k_host) Fortran-ordered view
        h2d = k * n_local * 8
        last = {}

        def e2e_step():
            if k_host == k:                               # public API: host shard -> HBM (copy stream; the
                arr = space.from_local_numpy(host_np)     # Gramian starts on the column blocks as they land)
            else:
                arr = space.empty(reserve=k)
                for c0 in range(0, k, k_host):            # host chunk -> HBM, appended
                    kc = min(k_host, k - c0)
                    arr.append(space.from_local_numpy(host_np[:, :kc]), remove_from_other=True)
            U, s = pod_step(arr)
            last['s'] = np.array(s)                       # singular values arrive on the host
            mm = len(U)
            del U, arr
            return mm

        for _ in range(min(args.warmup, 1)):
            e2e_step()
        n_e2e = max(1, min(args.steps, 3))
        et, m2 = timed_loop(e2e_step, n_e2e)
        t_e2e = float(np.mean(et))
        e2e = {'value': pod_flops(n_global, k, m2) / (t_e2e * 1e-3) * 1e-12, 'unit': 'TFLOP/s',
               'h2d_bytes_per_step': int(h2d + k * k * 8), 'd2h_bytes_per_step': int(2 * k * k * 8 + 8 * k),
               'ms_per_step': t_e2e, 'steps': n_e2e,
               'host_chunk_columns': k_host,
               'note': ('snapshots start in pinned host memory; space.from_local_numpy uploads the shard on a copy '
                        'stream in 8 column blocks and the Gramian contracts the blocks as they land'
                        if k_host == k else
                        'snapshots start in pinned host memory and are appended in chunks of host_chunk_columns '
                        'columns (from_local_numpy + append; N x 80 GB does not fit into host RAM)') +
                       '; H2D of the shard, Gram/check D2H and the coefficient H2D are inside the timed region; '
                       'the modes stay on the device'}
        del host_np
    except Exception as exc:                              # e.g. not enough host RAM to pin the shard
        e2e = {'value': None, 'unit': 'TFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
               'error': f'{type(exc).__name__}: {exc}'}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ---------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_sample
        cpu_pod_sample(5000, k)                            # spin up the BLAS threads
        dt, m_cpu = cpu_pod_sample(n_s, k)
        cpu_baseline = {'value': pod_flops(n_s, k, m_cpu) / dt * 1e-12, 'unit': 'TFLOP/s', 'cores': host_threads(),
                        'kind': 'port', 'seconds': dt,
                        'sample': f'oracle pod() (NumPy/OpenBLAS + SciPy eigh) on a {n_s} x {k} row sample, 1 run'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'TFLOP/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t_step, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'POD method_of_snapshots of {n_local} DOF x {k} fp64 snapshots per GPU '
                                   f'({n_global} DOF global), diagonal mass-matrix product, all {m} modes kept',
                       'n_per_gpu': n_local, 'k': k, 'modes_kept': int(m), 'spectrum': 'sigma_j = 10^(-2j/(k-1))',
                       'l2': 'inputs (80 GB per GPU) are far larger than the 126 MB L2; no flush needed',
                       'parallelism': f'row(DOF)-sharded x{world}, k x k all-reduce (NCCL)'},
            'phases_ms': {'gramian_each': gram_ms, 'eigh_host': eigh_ms, 'lincomb': lin_ms,
                          'device_tflops_excl_host_eigh': flop_global / ((t_step - eigh_ms) * 1e-3) * 1e-12},
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--dofs', dest='n', type=int, default=10_000_000, help='DOFs per GPU')
    ap.add_argument('--k', type=int, default=1000, help='snapshots')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --dofs per GPU (default); strong: --dofs in total, sharded over the GPUs')
    ap.add_argument('--cpu-sample', type=int, default=150_000, help='rows of the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg (profiling runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == '__main__':
    main()
