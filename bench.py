#!/usr/bin/env python
"""Benchmark of the hot path: POD (method of snapshots, diagonal mass-matrix product) of a fp64 snapshot matrix --
BASELINE.json's metric and config 4, "POD of 10M DOF x 1000 fp64 snapshots with W-weighted gramian, DOF-sharded
1/2/4/8 GPUs".

    python bench.py --gpus N --steps K --warmup W            # pyMOR's pod() on this backend (CUDA kernels)
    python bench.py --impl reference --steps K --warmup W    # pyMOR's pod() on its own NumPy backend (host cores)

BOTH arms call the reference's own, unmodified ``pymor.algorithms.pod.pod(A, product=W,
method='method_of_snapshots')`` (src/pymor/algorithms/pod.py:16, svd_va.py:20); they differ only in the classes
behind it: ``CudaVectorSpace`` / ``CudaDiagonalOperator`` (device-resident shards, libpymor_b200.so) versus
``NumpyVectorSpace`` / ``NumpyMatrixOperator`` (NumPy/OpenBLAS + SciPy).

One step = one full call: weighted Gramian (SYRK, TMA + FP64 tensor-core kernel) -> all-reduce of the k x k partials
(N > 1) -> host `eigh` (inside pyMOR) -> block lincomb `A (V / s)` -> orthonormality check (second weighted Gramian).
Nothing is skipped and nothing is cached between steps.  `value` = algorithmic flop / wall time of the whole call,
host eigensolve included.

Scaling is STRONG (BASELINE config 4 as written): the `--dofs` x `--k` problem is fixed and its rows are sharded
over the N GPUs.  The host eigensolve does not shrink with N and bounds the speed-up (Amdahl); the
weak-scaling figure (`--dofs` rows PER GPU) is reported next to it under `weak`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path


def parse_args(argv):
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--dofs', dest='n', type=int, default=10_000_000, help='DOFs (global; per GPU with --scaling weak)')
    ap.add_argument('--k', type=int, default=1000, help='snapshots')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong (default): --dofs in total, rows sharded over the GPUs; weak: --dofs per GPU')
    ap.add_argument('--cpu-sample', type=int, default=150_000, help='rows of the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer legs (profiling runs)')
    ap.add_argument('--no-weak', action='store_true', help='N > 1: skip the extra weak-scaling measurement')
    ap.add_argument('--no-parity', action='store_true', help='skip the parity block')
    ap.add_argument('--c5', default='auto', choices=['auto', 'on', 'off'],
                    help="secondary block: BASELINE config 5, project() of a 5-point CSR operator onto 2000 basis vectors, "
                         "6.25M DOF per GPU (50M at 8 GPUs); auto = at 8 GPUs")
    ap.add_argument('--no-gs', action='store_true',
                    help='skip the secondary block "config 3: re-iterated Gram-Schmidt of 10M x 256 with a diagonal product"')
    ap.add_argument('--no-csr-pod', action='store_true',
                    help='N = 1: skip the secondary block "POD with a CSR (tridiagonal FEM mass matrix) product"')
    ap.add_argument('--accelerate', action='store_true',
                    help='also time pod() with pymor_b200.accelerate_pymor() active (extra key, never the headline)')
    return ap.parse_args(argv)


ARGS = parse_args(sys.argv[1:] if __name__ == '__main__' else [])     # (importing bench must not read argv)
WORLD = int(os.environ.get('WORLD_SIZE', '1'))
RANK = int(os.environ.get('RANK', '0'))

# one 74.5 GiB panel next to another: let the caching allocator remap instead of fragmenting
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
# Host BLAS/LAPACK threads (must be set before NumPy/OpenBLAS is loaded).  torchrun pins every rank to ONE host
# thread by default.  The reference arm runs on rank 0 alone with ALL host cores whatever N is; in the CUDA arm the
# k x k eigensolve stays on the host and is replicated, so every rank gets its share of the cores.
if ARGS.impl == 'reference':
    if RANK != 0:                                     # under torchrun rank 0 alone runs the reference arm
        sys.exit(0)
    os.environ['OMP_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = str(os.cpu_count() or 1)
elif WORLD > 1:
    _local_world = int(os.environ.get('LOCAL_WORLD_SIZE', WORLD))
    _share = max(1, (os.cpu_count() or 1) // _local_world)
    if os.environ.get('OMP_NUM_THREADS', '1') == '1':
        os.environ['OMP_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = str(_share)
    # ... and its own block of cores: N ranks x (LAPACK threads + NCCL / driver helper threads) migrating over one
    # socket make the replicated eigensolves finish at different times, and the slowest rank sets the step time
    try:
        _cores = sorted(os.sched_getaffinity(0))
        _lr = int(os.environ.get('LOCAL_RANK', '0'))
        _per = len(_cores) // _local_world
        if _per >= 1 and os.environ.get('PMC_BENCH_NO_PINNING', '0') != '1':
            os.sched_setaffinity(0, _cores[_lr * _per:(_lr + 1) * _per])
    except (AttributeError, OSError):
        pass

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'POD (pymor pod, method_of_snapshots, diag-W product) FP64 TFLOP/s, 10M DOF x 1000 snapshots'

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel, from `ncu --set full` captures of this
# very workload, keyed by (kernel, rows per GPU, k); see profiles/ for the reports.
NCU_TRAFFIC = {
    ('gram', 10_000_000, 1000): (238.067654e9 + 3.538929e9, 'profiles/r02_ncu_gram2_default_10Mx1000_dram.csv (one weighted SYRK launch; '
                                 '80.08 GB of operands: the tiles of a row range are dispatched staggered and re-read '
                                 'part of it from DRAM, DESIGN.md §3)'),
    ('lincomb', 10_000_000, 1000): (80.138563e9 + 79.989229e9, 'profiles/r01_ncu_full_lincomb_10Mx1000_details.txt'),
    ('lincomb', 2_000_000, 500): (8.006061e9 + 7.963813e9, 'profiles/r01b_ncu_full_details.txt'),
}


def pod_flops(n, k, m):
    """Algorithmic flop of POD by the method of snapshots: SYRK Gramian + lincomb +
    orthonormality-check Gramian."""
    return n * k * (k + 1) + 2 * n * k * m + n * m * (m + 1)


def spectrum(k):
    return 10.0 ** (-2 * np.arange(k) / max(k - 1, 1))


def workload_config(args, world):
    """The `config` object -- identical in both arms (the reference arm times a row sample of this workload and
    says so under `cpu_baseline.sample`)."""
    n_global = args.n if args.scaling == 'strong' else args.n * world
    return {'workload': f'pymor.algorithms.pod.pod(A, product=W, method="method_of_snapshots") of {n_global} DOF x '
                        f'{args.k} fp64 snapshots, W = diagonal mass matrix, all modes kept (rtol 1e-7)',
            'n_global': n_global, 'k': args.k, 'spectrum': 'A = Gn diag(sigma) Q^T, sigma_j = 10^(-2j/(k-1)), seed 42',
            'scaling': args.scaling,
            'parallelism': f'rows (DOFs) sharded over {world} GPU(s); k x k Gram partials all-reduced (NCCL)',
            'l2': 'operands (>= 10 GB per GPU) are far larger than the 126 MB L2; no flush needed'}


def quiet_pymor():
    from pymor_b200 import interface  # noqa: F401  (puts pyMOR on sys.path: installed, or baseline/_ref)
    from pymor.core.logger import set_log_levels
    set_log_levels({'pymor': 'ERROR'})


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: pyMOR's own NumPy backend through the same pod() call
# ------------------------------------------------------------------------------------------------
_cpu_samples = {}


def cpu_pod_sample(n, k, seed=42):
    """One `pod` of an n x k row sample of the workload with pyMOR's NumpyVectorSpace + NumpyMatrixOperator on the
    host cores.  Returns seconds, modes kept.  The sample (same construction as the device-side workload) is
    generated once per size and re-used: pod() does not modify its input, and only the call itself is timed."""
    import scipy.sparse as sps
    from pymor.algorithms.pod import pod
    from pymor.operators.numpy import NumpyMatrixOperator
    from pymor.vectorarrays.numpy import NumpyVectorSpace
    if (n, k, seed) not in _cpu_samples:
        _cpu_samples.clear()                               # one sample at a time (400 000 x 1000 is 3.2 GB)
        rng = np.random.default_rng(seed)
        Gn = rng.standard_normal((n, k)) / np.sqrt(n)
        Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
        A = np.asfortranarray(Gn @ (np.diag(spectrum(k)) @ Qk.T))
        w = rng.uniform(0.5, 1.5, n)
        del Gn
        _cpu_samples[(n, k, seed)] = (NumpyVectorSpace(n).from_numpy(A), NumpyMatrixOperator(sps.diags(w).tocsr()))
    VA, W = _cpu_samples[(n, k, seed)]
    t0 = time.perf_counter()
    U, s = pod(VA, product=W, method='method_of_snapshots')
    dt = time.perf_counter() - t0
    return dt, len(U)


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
    quiet_pymor()
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
    sample = (f'the unmodified reference (pyMOR {_pymor_version()}, NumpyVectorSpace + NumpyMatrixOperator, NumPy/OpenBLAS + '
              f'SciPy eigh) through the same pod() call on a {n_s} x {k} row sample of the workload, mean of {args.steps} '
              f'steps, {cores} BLAS threads (all host cores, whatever --gpus is); rate = algorithmic flop of the sample '
              f'/ time (the path is linear in the number of rows)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'TFLOP/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, args.gpus),
        'cpu_baseline': {'value': val, 'unit': 'TFLOP/s', 'cores': cores, 'kind': 'reference', 'sample': sample,
                         'sample_rows': n_s, 'modes_kept': int(m)},
        'e2e': {'value': val, 'unit': 'TFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def _pymor_version():
    import pymor
    return pymor.__version__


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

    quiet_pymor()
    import pymor.algorithms.svd_va as svd_va_mod
    from pymor.algorithms.pod import pod

    from pymor_b200.dist import ProcessGroupComm
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace

    world, rank = WORLD, RANK
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        comm = ProcessGroupComm()
    k = args.k

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    def make_problem(n_global, seed=42):
        """Synthetic snapshots A = Gn diag(sigma) Q^T generated on the device, rows sharded over the ranks."""
        space = CudaVectorSpace(n_global, device=device, comm=comm)
        rng = np.random.default_rng(seed)
        Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
        mix = np.diag(spectrum(k)) @ Qk.T
        Gn = space.random_device(k, seed=seed, distribution='normal', scale=1.0 / np.sqrt(n_global))
        A = Gn.lincomb(mix)
        del Gn
        gen = torch.Generator(device=device).manual_seed(100 * seed + 42 + rank)
        w = torch.rand(space.local_dim, generator=gen, device=device, dtype=torch.float64) + 0.5
        prod = CudaDiagonalOperator(w, space)
        torch.cuda.synchronize(device)
        torch.cuda.empty_cache()
        return space, A, w, prod

    def timed_loop(step_fn, steps):
        times, out = [], None
        for _ in range(steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step_fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        return times, out

    def pod_step(arr, prod, **kw):
        return pod(arr, product=prod, method='method_of_snapshots', **kw)     # pyMOR's own function

    n_global = args.n if args.scaling == 'strong' else args.n * world
    space, A, w, prod = make_problem(n_global)
    n_local = space.local_dim
    ctx = space.ctx

    def resident_step():
        U, s = pod_step(A, prod)
        m = len(U)
        del U
        return m

    # ---- host-side phase timer: the eigensolve pyMOR runs between the two device phases ------------------
    eigh_ms = []
    orig_eigh = svd_va_mod.spla.eigh

    def timed_eigh(*a, **kw):
        t0 = time.perf_counter()
        out = orig_eigh(*a, **kw)
        eigh_ms.append((time.perf_counter() - t0) * 1e3)
        return out

    for _ in range(args.warmup):
        resident_step()
    svd_va_mod.spla.eigh = timed_eigh
    ctx.kernel_times()                                   # clear
    ctx.set_option('kernel_timing', 1)                   # CUDA events around each tensor-core kernel launch
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    times, m = timed_loop(resident_step, args.steps)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_option('kernel_timing', 0)
    ktimes = ctx.kernel_times()
    svd_va_mod.spla.eigh = orig_eigh
    t_step = float(np.mean(times))                       # ms, max over ranks per step
    flop_global = pod_flops(n_global, k, m)
    value = flop_global / (t_step * 1e-3) * 1e-12
    eigh_mean = float(np.mean(eigh_ms)) if eigh_ms else 0.0

    # ---- roofline of the time-dominant kernel (kernel times from events around the launches) --------------
    peak_dmma = ctx.measure_fp64_peak(True, 300.0)
    peak_dfma = ctx.measure_fp64_peak(False, 100.0)
    gram_ms = [ms for kind, ms in ktimes if kind == 'gram']
    lin_ms = [ms for kind, ms in ktimes if kind == 'lincomb']
    per_step = max(args.steps, 1)
    kern = {}
    if gram_ms:
        g = float(np.mean(gram_ms))
        # both Gramians of a step are SYRKs (k x k; the check is m x m with m == k here)
        flop = n_local * (k * (k + 1.0) + m * (m + 1.0)) / 2.0
        kern['gram'] = {'name': f'gram{"2" if ctx.get_option("gram_opts") & 16 else ""}_dmma_kernel<weighted> (SYRK A^T W A)',
                        'launches_per_step': len(gram_ms) / per_step, 'launch_ms': g, 'ms_per_step': sum(gram_ms) / per_step,
                        'algorithmic_flop_per_launch': flop, 'algorithmic_bytes_per_launch': 8.0 * n_local * (k + 1),
                        'achieved': flop / (g * 1e-3) * 1e-12}
    if lin_ms:
        g = float(np.mean(lin_ms))
        flop = 2.0 * n_local * k * m
        kern['lincomb'] = {'name': 'lincomb_dmma_kernel (A C)', 'launches_per_step': len(lin_ms) / per_step, 'launch_ms': g,
                           'ms_per_step': sum(lin_ms) / per_step, 'algorithmic_flop_per_launch': flop,
                           'algorithmic_bytes_per_launch': 8.0 * n_local * (k + m), 'achieved': flop / (g * 1e-3) * 1e-12}
    for v in kern.values():
        v['frac'] = v['achieved'] / peak_dmma if peak_dmma else None
        v['share_of_step'] = v['ms_per_step'] / t_step
    dom = max(kern, key=lambda name: kern[name]['ms_per_step']) if kern else None
    roofline = None
    if dom:
        d = kern[dom]
        traffic = NCU_TRAFFIC.get((dom, n_local, k))
        roofline = {'bound': 'tensor', 'kernel': d['name'], 'achieved': d['achieved'], 'peak': peak_dmma,
                    'unit': 'TFLOP/s', 'frac': d['frac'], 'traffic': traffic[0] if traffic else None,
                    'traffic_source': traffic[1] if traffic else None,
                    'algorithmic_flop_per_launch': d['algorithmic_flop_per_launch'],
                    'algorithmic_bytes_per_launch': d['algorithmic_bytes_per_launch'], 'launch_ms': d['launch_ms'],
                    'launches_per_step': d['launches_per_step'], 'share_of_step': d['share_of_step'],
                    'timing': 'CUDA events recorded by the library around each launch of the kernel, on the launching '
                              'stream, inside the timed steps (pmc_ctx_kernel_times)',
                    'peak_source': 'FP64 tensor-core (DMMA.8x8x4) issue-rate microbenchmark measured in this run '
                                   '(pmc_measure_fp64_peak); MEASURED_PEAKS.json carries no fp64 figure',
                    'dfma_peak_tflops': peak_dfma, 'kernels': kern}

    # ---- parity (outside the timed region) -----------------------------------------------------------------
    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(args, space, A, w, prod, comm, device, pod_step)
        except Exception as exc:                          # never lose the contract line to the checker
            parity = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- optional: the same call with accelerate_pymor() active ------------------------------------------
    accelerated = None
    if args.accelerate:
        from pymor_b200.accelerate import accelerate_pymor, restore_pymor
        accelerate_pymor()
        try:
            for _ in range(min(args.warmup, 2)):
                resident_step()
            at, am = timed_loop(resident_step, max(2, min(args.steps, 5)))
            accelerated = {'value': pod_flops(n_global, k, am) / (float(np.mean(at)) * 1e-3) * 1e-12, 'unit': 'TFLOP/s',
                           'ms_per_step': float(np.mean(at)),
                           'note': 'pod() with pymor_b200.accelerate_pymor() installed (opt-in dispatcher)'}
        finally:
            restore_pymor()

    # ---- end to end: host snapshots (pinned) -> H2D -> pod -> D2H ------------------------------------------
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError('skipped (--no-e2e)')
        host, k_host = pinned_copy_of(A, world)
        del A                                             # the e2e legs re-create the array from the host copy
        torch.cuda.empty_cache()
        e2e = e2e_legs(args, space, host, k_host, prod, world, n_global, timed_loop, pod_step, device)
    except Exception as exc:                              # e.g. not enough host RAM to pin the shard
        e2e = {'value': None, 'unit': 'TFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
               'error': f'{type(exc).__name__}: {exc}'}
    A = None
    del prod, w, space
    torch.cuda.empty_cache()

    # ---- N = 1: the same POD with a CSR mass-matrix product (north star: "diagonal or CSR mass-matrix weight") ----
    csr_pod = None
    if world == 1 and not args.no_csr_pod:
        try:
            csr_pod = csr_weighted_pod(args, make_problem, timed_loop, pod_step)
        except Exception as exc:
            csr_pod = {'value': None, 'error': f'{type(exc).__name__}: {exc}'}
        torch.cuda.empty_cache()

    # ---- N > 1: the weak-scaling figure next to the strong one ---------------------------------------------
    weak = None
    if world > 1 and args.scaling == 'strong' and not args.no_weak:
        try:
            wspace, wA, ww, wprod = make_problem(args.n * world, seed=43)

            def weak_step():
                U, s = pod_step(wA, wprod)
                mm = len(U)
                del U
                return mm

            weak_step()
            wt, wm = timed_loop(weak_step, 2)
            weak = {'value': pod_flops(args.n * world, k, wm) / (float(np.mean(wt)) * 1e-3) * 1e-12, 'unit': 'TFLOP/s',
                    'ms_per_step': float(np.mean(wt)), 'steps': 2, 'n_per_gpu': args.n, 'n_global': args.n * world}
            del wA, wprod, ww, wspace
        except Exception as exc:
            weak = {'value': None, 'error': f'{type(exc).__name__}: {exc}'}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload -------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_sample
        cpu_pod_sample(5000, k)                            # spin up the BLAS threads
        dt, m_cpu = cpu_pod_sample(n_s, k)
        cpu_baseline = {'value': pod_flops(n_s, k, m_cpu) / dt * 1e-12, 'unit': 'TFLOP/s', 'cores': host_threads(),
                        'kind': 'reference', 'seconds': dt,
                        'sample': f'the unmodified reference (pyMOR {_pymor_version()} NumPy backend) through the same pod() '
                                  f'call on a {n_s} x {k} row sample, 1 run'}

    line = None
    if rank == 0:
        cfg = workload_config(args, world)
        cfg.update({'n_per_gpu': n_local, 'modes_kept': int(m)})
        device_ms = sum(v['ms_per_step'] for v in kern.values())
        line = {
            'metric': METRIC, 'value': value, 'unit': 'TFLOP/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t_step, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'phases_ms': {'tensor_kernels': device_ms, 'eigh_host': eigh_mean,
                          'other (reduce kernels, all-reduce, D2H/H2D of k x k, Python)': t_step - device_ms - eigh_mean,
                          'device_tflops_excl_host_eigh': flop_global / ((t_step - eigh_mean) * 1e-3) * 1e-12,
                          'host_blas_threads_per_rank': host_threads()},
            'roofline': roofline, 'parity': parity, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'weak': weak,
            'secondary': {'pod_csr_product': csr_pod} if csr_pod else None, 'accelerated': accelerated, 'gpu_launches': int(launches), 'clocks': clocks,
        }

    # ---- secondary blocks: BASELINE config 3 (re-iterated Gram-Schmidt, 10M x 256) at every N and config 5 (RB
    # projection V^T A V, 50M DOF x 2000 basis) at 8 GPUs.  They run last and under a watchdog: whatever happens
    # here, the contract line above is printed exactly once.
    want_c5 = args.c5 == 'on' or (args.c5 == 'auto' and world == 8)
    if want_c5 or not args.no_gs:
        extra = {}

        def give_up():
            if rank == 0:
                extra['error'] = 'secondary blocks timed out after 400 s'
                line['secondary'] = dict(line['secondary'] or {}, **extra)
                print(json.dumps(line), flush=True)
            os._exit(0)
        watchdog = threading.Timer(400.0, give_up)
        watchdog.daemon = True
        watchdog.start()
        torch.cuda.empty_cache()
        if not args.no_gs:
            try:
                extra['gram_schmidt_c3'] = gram_schmidt_c3(device, comm, barrier)
            except Exception as exc:
                extra['gram_schmidt_c3'] = {'error': f'{type(exc).__name__}: {exc}'}
            torch.cuda.empty_cache()
        if want_c5:
            try:
                sys.path.insert(0, str(ROOT / 'tools'))
                from bench_project import run_projection
                nx, ny = 5000, 1250 * world                # 6.25M DOF per GPU; the shard boundary is one grid line
                pspace = CudaVectorSpace(nx * ny, device=device, comm=comm)
                barrier()
                extra['projection_c5'] = run_projection(pspace, nx, ny, 2000, reps=2, comm=comm)
            except Exception as exc:
                extra['projection_c5'] = {'error': f'{type(exc).__name__}: {exc}'}
        watchdog.cancel()
        if rank == 0:
            line['secondary'] = dict(line['secondary'] or {}, **extra)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gram_schmidt_c3(device, comm, barrier, n=10_000_000, k=256):
    """BASELINE config 3: re-iterated Gram-Schmidt of k vectors with a diagonal mass-matrix product, rows sharded
    over the ranks.  Input family of SURVEY.md §8d (a_i = g_i + 2 s: every vector re-iterates exactly once, k(k-1)
    pair-steps).  pyMOR's own `gram_schmidt` (one host-synchronised dot + axpy per pair: gram_schmidt.py:84-91) next to
    the device-friendly formulations of the same factorisation: `sweep=True` (one fused kernel per pass, sums over
    the GPUs inside the kernel), `variant='bcgs2'` (panels on the tensor cores) and pyMOR's `shifted_chol_qr`.  Every
    result is checked: orthonormality, A = Q R, and R against pyMOR's R."""
    import torch
    from pymor.algorithms.chol_qr import shifted_chol_qr
    from pymor.algorithms.gram_schmidt import gram_schmidt

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    rank = comm.rank if comm is not None else 0
    space = CudaVectorSpace(n, device=device, comm=comm)
    A0 = space.random_device(k, seed=1, scale=1.0 / np.sqrt(n))
    A0.axpy(2.0, space.random_device(1, seed=2, scale=1.0 / np.sqrt(n)))
    gen = torch.Generator(device=device).manual_seed(3 + rank)
    w = torch.rand(space.local_dim, generator=gen, device=device, dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, space)
    # (name, function, runs): the variants that allocate full-size temporaries (block lincombs) are run more than
    # once and the best run is reported -- their first call pays for the caching allocator's fresh 20 GB mappings
    variants = [
        ('pymor_gram_schmidt', lambda A: gram_schmidt(A, product=prod, return_R=True, check=False, copy=False), 1),
        ('sweep', lambda A: alg.gram_schmidt(A, product=prod, return_R=True, check=False, copy=False, sweep=True), 1),
        ('bcgs2', lambda A: alg.gram_schmidt(A, product=prod, return_R=True, check=False, copy=False, variant='bcgs2'), 2),
        ('pymor_shifted_chol_qr', lambda A: shifted_chol_qr(A, product=prod, return_R=True, copy=False,
                                                            product_norm=float(np.sqrt(1.5))), 3),
    ]
    out, R_ref = {}, None
    anorm = float(np.max(A0.norm(prod)))
    for name, fn, runs in variants:
        fn(A0[:8].copy())                                       # warm the kernels / the code path
        all_dt = []
        for _ in range(runs):
            Q = R = None
            A = A0.copy(deep=True)
            barrier()
            t0 = time.perf_counter()
            Q, R = fn(A)
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            if comm is not None:
                dt = max(comm.allgather_object(dt))
            all_dt.append(dt)
        dt = min(all_dt)
        D = Q.lincomb(R)
        D.axpy(-1.0, A0)
        res = {'seconds': dt, 'all_seconds': all_dt, 'orth_err_max_abs(Q^T W Q - I)': float(np.max(np.abs(Q.gramian(prod) - np.eye(len(Q))))),
               'reconstruction_rel_err(A - Q R)': float(np.max(D.norm(prod)) / anorm), 'len_Q': len(Q)}
        if name == 'pymor_gram_schmidt':
            R_ref = R
            res['us_per_pair_step'] = dt / (k * (k - 1)) * 1e6
        elif R_ref is not None and R.shape == R_ref.shape:
            res['R_vs_pymor_gram_schmidt_rel'] = float(np.max(np.abs(R - R_ref)) / np.max(np.abs(R_ref)))
        out[name] = res
        del Q, D, A
    out['workload'] = (f'orthonormalisation of {n} x {k}, diagonal product, a_i = g_i + 2 s (every vector re-iterates once: '
                       f'{k * (k - 1)} pair-steps in modified Gram-Schmidt), rows sharded over {comm.size if comm else 1} GPU(s)')
    return out


def csr_weighted_pod(args, make_problem, timed_loop, pod_step):
    """pod(A, product=M, method='method_of_snapshots') with M the 1-D FEM mass matrix h/6 [1 4 1] as a device CSR
    operator: the weighted Gramians run as the fused CSR two-form -- row panels of M A contracted by
    the tensor-core kernel on the lower tile triangle only (M is symmetric), no n x k temporary."""
    import scipy.sparse as sps

    from pymor_b200.operators import CudaCsrOperator
    space, A, w, prod = make_problem(args.n, seed=44)
    del w, prod
    n, k = space.dim, args.k
    h = 1.0 / (n + 1)
    M = sps.diags([np.full(n - 1, h / 6), np.full(n, 4 * h / 6), np.full(n - 1, h / 6)], [-1, 0, 1], format='csr')
    op = CudaCsrOperator.from_local_rows(M, space, symmetric=True)
    nnz = M.nnz
    del M
    A.scal(1.0 / np.sqrt(h))                                   # ||a_j||_M = O(1) like in the diagonal case

    def step():
        U, s = pod_step(A, op)
        mm = len(U)
        del U
        return mm

    step(); step()
    times, m = timed_loop(step, 3)
    t = float(np.mean(times))
    U, s = pod_step(A, op)
    orth = float(np.max(np.abs(U.gramian(op) - np.eye(len(U)))))
    trace = float(np.sum(A.norm2(op)))
    flop = n * k * (k + 1.0) + 2.0 * n * k * m + n * m * (m + 1.0) + 2.0 * nnz * (k + m)
    return {'value': flop / (t * 1e-3) * 1e-12, 'unit': 'TFLOP/s', 'ms_per_step': t, 'steps': 3, 'modes_kept': int(m),
            'workload': f'pymor pod(method_of_snapshots) of {n} x {k}, product = CSR tridiagonal FEM mass matrix ({nnz} nnz)',
            'algorithmic_flop': flop,
            'parity': {'orth_err_max_abs(U^T M U - I)': orth,
                       'trace_identity_rel_err': float(abs(np.sum(s ** 2) - trace) / trace)}}


def parity_block(args, space, A, w, prod, comm, device, pod_step):
    """Correctness evidence from THIS run, at this N.  (1) A row sample of the workload -- the first rows of every
    rank's shard, i.e. a DOF-sharded problem of its own -- goes through the same sharded pod() call and, gathered on
    rank 0, through the unmodified reference (pyMOR's NumPy backend) on the identical bits.  (2) Size-independent
    properties of the full-size result."""
    import scipy.sparse as sps
    import torch
    import torch.distributed as dist
    from pymor.algorithms.pod import pod
    from pymor.operators.numpy import NumpyMatrixOperator
    from pymor.vectorarrays.numpy import NumpyVectorSpace

    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    world, rank = (comm.size, comm.rank) if comm is not None else (1, 0)
    k = args.k
    out = {}
    # (1) sample parity against the reference
    r = max(16, (20000 // world) // 16 * 16)
    r = min(r, space.local_dim)
    sspace = CudaVectorSpace(r * world, device=device, comm=comm)
    shard = A.to_torch()[:, :r].contiguous()                       # (k, r): this rank's rows of the sample
    ws = w[:r].contiguous()
    SA = sspace.make_array(shard.clone())
    sprod = CudaDiagonalOperator(ws, sspace)
    U, s = pod_step(SA, sprod)
    Uh = U.to_torch()[:, :r].contiguous()
    if world > 1:
        parts = [torch.empty_like(shard) for _ in range(world)] if rank == 0 else None
        dist.gather(shard, parts, dst=0)
        wparts = [torch.empty_like(ws) for _ in range(world)] if rank == 0 else None
        dist.gather(ws, wparts, dst=0)
        uparts = [torch.empty_like(Uh) for _ in range(world)] if rank == 0 else None
        dist.gather(Uh, uparts, dst=0)
    else:
        parts, wparts, uparts = [shard], [ws], [Uh]
    if rank == 0:
        host = np.asfortranarray(torch.cat(parts, dim=1).cpu().numpy().T)       # (r * world, k)
        wh = torch.cat(wparts).cpu().numpy()
        Ug = torch.cat(uparts, dim=1).cpu().numpy().T
        nsp = NumpyVectorSpace(host.shape[0])
        Uref, sref = pod(nsp.from_numpy(host), product=NumpyMatrixOperator(sps.diags(wh).tocsr()),
                         method='method_of_snapshots')
        Ur = Uref.to_numpy()
        mm = min(len(s), len(sref))
        # largest principal angle between the two POD subspaces in the W inner product, from its SINE:
        # || (I - Ur Ur^T W) Ug ||_2 (small angles are invisible in the cosines)
        resid = Ug[:, :mm] - Ur[:, :mm] @ ((Ur[:, :mm] * wh[:, None]).T @ Ug[:, :mm])
        sine = float(np.sqrt(np.linalg.norm((resid * wh[:, None]).T @ resid, 2)))
        out['sample'] = {
            'rows': int(host.shape[0]), 'k': k, 'reference': f'pyMOR {_pymor_version()} NumPy backend, same pod() call, same bits',
            'modes_kept': [int(len(s)), int(len(sref))],
            'singular_values_max_rel_err': float(np.max(np.abs(s[:mm] - sref[:mm]) / sref[:mm])),
            'subspace_angle_rad': float(np.arcsin(min(1.0, sine))),
            'orth_err_device': float(np.max(np.abs((Ug * wh[:, None]).T @ Ug - np.eye(Ug.shape[1])))),
            'orth_err_reference': float(np.max(np.abs((Ur * wh[:, None]).T @ Ur - np.eye(Ur.shape[1])))),
            'tolerances': {'singular_values_rel': 1e-10, 'subspace_angle': 1e-8}}
        out['sample']['ok'] = bool(out['sample']['singular_values_max_rel_err'] < 1e-10
                                   and out['sample']['subspace_angle_rad'] < 1e-8
                                   and len(s) == len(sref))
    del SA, sprod, U, shard, Uh
    # (2) full-size properties
    U, s = pod_step(A, prod)
    G = U.gramian(prod)
    trace_norms = float(np.sum(A.norm2(prod)))                      # sum_j ||a_j||_W^2 through the streaming reduction kernel
    out['full_size'] = {
        'orth_err_max_abs(U^T W U - I)': float(np.max(np.abs(G - np.eye(len(U))))),
        'trace_identity_rel_err(sum s^2 vs sum ||a_j||_W^2)': float(abs(np.sum(s ** 2) - trace_norms) / trace_norms),
        'sum_s2': float(np.sum(s ** 2)), 'sum_norm2': trace_norms,
        'gram_bitwise_symmetric': bool(np.array_equal(G, G.T)),
        'singular_values_sorted': bool(np.all(np.diff(s) <= 0)), 'modes_kept': int(len(U))}
    del U
    return out


def pinned_copy_of(A, world):
    """This rank's snapshot shard in pinned host memory -- the starting point of the e2e legs.  One rank's shard is
    n_local x k doubles; if the host cannot pin all shards at once (N ranks share one host), the buffer holds as
    many snapshot columns as a 45 % share of the free host memory allows and is re-used for the later chunks (every
    byte is still copied inside the timed region)."""
    import torch
    k, n_local = len(A), A.space.local_dim
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
    return host, k_host


def e2e_legs(args, space, host, k_host, prod, world, n_global, timed_loop, pod_step, device):
    """The same call measured end to end through the public API with HOST buffers: every step copies this rank's
    snapshot shard from pinned host memory to the device (`space.from_local_numpy`), runs pod(), and reads the result
    back.  Leg 1 (the `e2e` value): all modes kept, singular values come back, the modes stay on the device (where
    the next pyMOR step -- projection -- uses them).  Leg 2 (`modes100`): pod(modes=100), the 100 modes are copied
    back to pinned host memory as well."""
    import torch
    k, n_local = args.k, space.local_dim
    host_np = host.numpy().T                          # (n_local, k_host) Fortran-ordered view
    h2d = k * n_local * 8
    last = {}

    def upload():
        if k_host == k:                               # public API: host shard -> HBM (copy stream; the
            return space.from_local_numpy(host_np)    # Gramian starts on the column blocks as they land)
        arr = space.empty(reserve=k)
        for c0 in range(0, k, k_host):                # host chunk -> HBM, appended
            kc = min(k_host, k - c0)
            arr.append(space.from_local_numpy(host_np[:, :kc]), remove_from_other=True)
        return arr

    def e2e_step():
        arr = upload()
        U, s = pod_step(arr, prod)
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
           'ms_per_step': t_e2e, 'steps': n_e2e, 'host_chunk_columns': k_host,
           'h2d_GBps_per_gpu_if_copy_bound': h2d / (t_e2e * 1e-3) * 1e-9,
           'note': ('snapshots start in pinned host memory; space.from_local_numpy uploads the shard on a copy stream in '
                    'column blocks and the Gramian contracts the blocks as they land'
                    if k_host == k else
                    'snapshots start in pinned host memory and are appended in chunks of host_chunk_columns columns') +
                   '; H2D of the shard, Gram/check D2H and the coefficient H2D are inside the timed region; all modes are '
                   'kept and stay on the device (see modes100 for the leg that returns modes to the host)'}
    # leg 2: truncated POD whose modes go back to the host
    mk = min(100, k)
    modes_host = torch.empty((mk, n_local), dtype=torch.float64, pin_memory=True)

    def e2e_modes_step():
        arr = upload()
        U, s = pod_step(arr, prod, modes=mk)
        mm = len(U)
        modes_host[:mm].copy_(U.to_torch(), non_blocking=True)
        last['s100'] = np.array(s)
        torch.cuda.current_stream(device).synchronize()
        del U, arr
        return mm

    e2e_modes_step()
    mt, m3 = timed_loop(e2e_modes_step, max(1, min(args.steps, 2)))
    t3 = float(np.mean(mt))
    e2e['modes100'] = {'value': pod_flops(n_global, k, m3) / (t3 * 1e-3) * 1e-12, 'unit': 'TFLOP/s', 'ms_per_step': t3,
                       'modes_kept': int(m3), 'h2d_bytes_per_step': int(h2d + k * m3 * 8),
                       'd2h_bytes_per_step': int(m3 * n_local * 8 + k * k * 8 + m3 * m3 * 8 + 8 * m3),
                       'algorithmic_flop': pod_flops(n_global, k, m3),
                       'note': 'pod(modes=100): H2D of the shard, Gramian, host eigh (subset), lincomb to 100 modes, '
                               'orthonormality check, D2H of the 100 modes into pinned host memory'}
    return e2e


def main():
    if ARGS.impl == 'reference':
        run_reference(ARGS)
    else:
        run_cuda(ARGS)


if __name__ == '__main__':
    main()
