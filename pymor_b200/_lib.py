"""ctypes binding of libpymor_b200.so (the C ABI declared in include/pymor_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100a device is present, creating
a :class:`DeviceContext` raises.  All pointers cross the boundary as plain integers.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from pathlib import Path

import numpy as np
import torch

PMC_OK = 0
PMC_ERR_WORKSPACE = -3
PMC_SYNC = 1
PMC_SYMMETRIC = 2
PMC_FORCE_SIMPLE = 4
PMC_LOWER_ONLY = 8
PMC_LOCAL_ONLY = 16
PMC_MAX_PEERS = 8
PMC_IPC_HANDLE_BYTES = 64
PMC_COMM_ID_BYTES = 128

LIB_PATH = Path(__file__).resolve().parent / 'lib' / 'libpymor_b200.so'

_P, _I64, _I32, _U32 = c_void_p, c_int64, c_int32, c_uint32

# name -> argtypes (after the leading pmc_ctx*); every symbol of include/pymor_b200.h is listed here
_CTX_FUNCS = {
    'pmc_gram': [_P, _I64, _I32, _P, _I64, _I32, _I64, _P, _P, _I64, _U32],
    'pmc_lincomb': [_P, _I64, _I32, _I64, _P, _I64, _I32, _P, _I64, _U32],
    'pmc_pairwise_inner': [_P, _I64, _P, _I64, _I32, _I64, _P, _P, _U32],
    'pmc_norm2': [_P, _I64, _I32, _I64, _P, _P, _U32],
    'pmc_axpy': [_P, _I64, _P, _I64, _I32, _I64, _P, _I32, _U32],
    'pmc_scal': [_P, _I64, _I32, _I64, _P, _I32, _U32],
    'pmc_axpy_dot': [_P, _P, _P, _P, _I64, _P, _P, _U32],
    'pmc_copy': [_P, _I64, _P, _I64, _I32, _I64, _U32],
    'pmc_gather_cols': [_P, _I64, _P, _I64, _P, _I32, _I64, _U32],
    'pmc_shift_cols': [_P, _I64, _I64, _I64, _I64, _I64, _U32],
    'pmc_dofs': [_P, _I64, _I32, _I64, _P, _I32, _P, _U32],
    'pmc_amax': [_P, _I64, _I32, _I64, _P, _P, _U32],
    'pmc_diag_apply': [_P, _P, _I64, _P, _I64, _I32, _I64, _U32],
    'pmc_csr_apply': [_I64, _P, _P, _P, _P, _I64, _P, _P, _I64, _I32, _U32],
    'pmc_gather_rows': [_P, _I64, _I32, _I64, _P, _I64, _P, _U32],
    'pmc_csr_gram': [_I64, _P, _P, _P, _P, _I64, _I32, _P, _I64, _I32, _P, _P, _I64, _U32],
    'pmc_csr_pairwise': [_I64, _P, _P, _P, _P, _I64, _P, _I64, _P, _I32, _P, _U32],
    'pmc_measure_fp64_peak': [c_int, c_double, _P],
    'pmc_host_register': [_P, c_size_t, ctypes.POINTER(_P)],
    'pmc_host_unregister': [_P],
    'pmc_mgs_sweep': [_P, _I64, _P, _I64, _I32, _P, _I64, _P, _P, _U32],
    'pmc_peer_create': [c_int, c_int, _P],
    'pmc_peer_connect': [_P],
    'pmc_peer_destroy': [],
    'pmc_comm_init': [c_int, c_int, _P],
    'pmc_comm_destroy': [],
    'pmc_allreduce_sum': [_P, _I64, _U32],
    'pmc_mirror_lower': [_P, _I64, _I64, _U32],
    'pmc_ctx_set_option': [ctypes.c_char_p, c_int],
    'pmc_ctx_get_option': [ctypes.c_char_p, ctypes.POINTER(c_int)],
    'pmc_ctx_kernel_times': [_P, _P, _I32, ctypes.POINTER(_I32)],
}
_OTHER_FUNCS = {
    'pmc_version': ([], c_int),
    'pmc_last_error': ([], ctypes.c_char_p),
    'pmc_ctx_create': ([c_int, _P, ctypes.POINTER(_P)], c_int),
    'pmc_ctx_destroy': ([_P], c_int),
    'pmc_ctx_set_workspace': ([_P, _P, c_size_t], c_int),
    'pmc_ctx_workspace_needed': ([_P], c_size_t),
    'pmc_ctx_sync': ([_P], c_int),
    'pmc_ctx_sm_count': ([_P], c_int),
    'pmc_ctx_launch_count': ([_P], c_int64),
    'pmc_comm_unique_id': ([_P], c_int),
    'pmc_comm_size': ([_P], c_int),
}
EXPORTED_SYMBOLS = sorted(list(_CTX_FUNCS) + list(_OTHER_FUNCS))


class PmcError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the in-tree shared library and declare the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PmcError(f'{LIB_PATH} not found: build it with `python -m pymor_b200.build` '
                       '(nvcc, sm_100a).  There is no CPU fallback.')
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, args in _CTX_FUNCS.items():
        fn = getattr(lib, name)
        fn.argtypes = [_P] + args
        fn.restype = c_int
    for name, (args, res) in _OTHER_FUNCS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def set_library_for_testing(lib):
    """Inject an object exposing the same entry points (tests/emulated_lib.py) -- test use only."""
    global _lib
    _lib = lib
    _contexts.clear()


class DeviceContext:
    """One ``pmc_ctx`` per device: stream, workspace and pinned result buffers.

    Device memory (workspace included) is allocated through torch and lent to the library."""

    def __init__(self, device):
        self.lib = load_library()
        self._fns = {}
        self.device = torch.device(device)
        on_cuda = self.device.type == 'cuda'
        if on_cuda:
            index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.device = torch.device('cuda', index)
            stream = torch.cuda.current_stream(self.device).cuda_stream
        else:
            index, stream = 0, 0      # only reachable with an injected test library
        handle = _P()
        rc = self.lib.pmc_ctx_create(index, stream, ctypes.byref(handle))
        if rc != PMC_OK:
            raise PmcError(f'pmc_ctx_create failed: {self._err()}')
        self.handle = handle.value if isinstance(handle.value, int) else handle
        self.stream_ptr = int(stream)                  # the stream the library launches on
        self._pin = on_cuda
        self._ws = None
        self.ensure_workspace(64 << 20)
        # deferred single-vector axpy (see CudaVectorArrayImpl.axpy): at most one per device, flushed by
        # any other operation; merged with the next dot/norm on its target by pmc_axpy_dot
        self.uploads = []                              # impls whose host -> device copy is still in flight
        self._copy_stream = None
        self.pending = None
        self._peer_comm = None
        self._native_comm = None                       # (rank, size) of the library-side NCCL communicator
        self.fuse_axpy = os.environ.get('PMC_NO_FUSE_AXPY', '0') != '1'
        self.scalar = np.zeros(1)                      # staging for scalar coefficients (read at launch)
        self.scalar_ptr = self.scalar.ctypes.data
        self._coeff_stage = None
        self._coeff_event = None
        self._host_f64 = torch.empty(1 << 16, dtype=torch.float64, pin_memory=self._pin)
        self._host_i64 = torch.empty(1 << 12, dtype=torch.int64, pin_memory=self._pin)

    def _err(self):
        msg = self.lib.pmc_last_error()
        return msg.decode() if isinstance(msg, bytes) else str(msg)

    def ensure_workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            rc = self.lib.pmc_ctx_set_workspace(self.handle, self._ws.data_ptr(), self._ws.numel())
            if rc != PMC_OK:
                raise PmcError(self._err())

    def call(self, name, *args):
        """Invoke ``name(ctx, *args)``; grows the workspace and retries on PMC_ERR_WORKSPACE."""
        fn = self._fns.get(name)
        if fn is None:
            fn = self._fns[name] = getattr(self.lib, name)
        rc = fn(self.handle, *args)
        if rc == PMC_ERR_WORKSPACE:
            need = int(self.lib.pmc_ctx_workspace_needed(self.handle))
            self.ensure_workspace(need + min(need >> 2, 64 << 20))
            rc = fn(self.handle, *args)
        if rc != PMC_OK:
            raise PmcError(f'{name} failed ({rc}): {self._err()}')

    def copy_stream(self):
        """Side stream for host -> device uploads that overlap with compute."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def wait_uploads(self):
        """Make the compute stream wait for every upload still in flight."""
        ups, self.uploads = self.uploads, []
        cur = torch.cuda.current_stream(self.device)
        for impl in ups:
            for _, _, ev in impl._arrivals or ():
                cur.wait_event(ev)
            impl._arrivals = None
            impl._arrival_src = None

    def flush(self):
        """Launch the deferred axpy, if any; order the stream after uploads still in flight."""
        if self.uploads:
            self.wait_uploads()
        pend = self.pending
        if pend is not None:
            self.pending = None
            timpl, tcol, alpha, ximpl, xcol = pend
            self.scalar[0] = alpha
            self.call('pmc_axpy', timpl._ptr(tcol), timpl.ld, ximpl._ptr(xcol), ximpl.ld, 1, timpl.n,
                      self.scalar_ptr, 1, 0)

    def upload_coefficients(self, coefficients, k, m, ldc):
        """Host coefficient matrix ``(k, m)`` -> device tensor ``(m, ldc)`` (= Fortran-ordered ``(ldc, m)``) through a
        pinned staging buffer and one asynchronous DMA on the compute stream.  The staging buffer is re-used: the
        event of the previous upload is waited for before it is overwritten."""
        if not self._pin:                                   # injected test library: plain host tensors
            host = np.zeros((ldc, m), dtype=np.float64, order='F')
            host[:k] = coefficients
            return torch.from_numpy(host.T).to(self.device)
        count = ldc * m
        stage = self._coeff_stage
        if stage is None or stage.numel() < count:
            if self._coeff_event is not None:
                self._coeff_event.synchronize()
            stage = self._coeff_stage = torch.empty(max(int(count), 1 << 16), dtype=torch.float64, pin_memory=True)
        elif self._coeff_event is not None:
            self._coeff_event.synchronize()
        view = stage[:count].numpy().reshape((ldc, m), order='F')
        view[:k] = coefficients
        if ldc > k:
            view[k:] = 0.0
        ct = torch.empty((m, ldc), dtype=torch.float64, device=self.device)
        ct.copy_(stage[:count].view(m, ldc), non_blocking=True)
        ev = self._coeff_event
        if ev is None:
            ev = self._coeff_event = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return ct

    def host_f64(self, count):
        """Pinned (device-writable) host buffer of at least ``count`` doubles."""
        if self._host_f64.numel() < count:
            self._host_f64 = torch.empty(int(count), dtype=torch.float64, pin_memory=self._pin)
        return self._host_f64

    def host_i64(self, count):
        if self._host_i64.numel() < count:
            self._host_i64 = torch.empty(int(count), dtype=torch.int64, pin_memory=self._pin)
        return self._host_i64

    def sync(self):
        self.flush()
        rc = self.lib.pmc_ctx_sync(self.handle)
        if rc != PMC_OK:
            raise PmcError(self._err())

    @property
    def sm_count(self):
        return int(self.lib.pmc_ctx_sm_count(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.pmc_ctx_launch_count(self.handle))

    def register_host(self, host_ptr, nbytes):
        """Pin caller-owned host memory; returns the address kernels may write to."""
        out = _P()
        rc = self.lib.pmc_host_register(self.handle, host_ptr, nbytes, ctypes.byref(out))
        if rc != PMC_OK:
            raise PmcError(f'pmc_host_register failed: {self._err()}')
        return int(out.value)

    def unregister_host(self, host_ptr):
        self.lib.pmc_host_unregister(self.handle, host_ptr)

    def ensure_peers(self, comm):
        """Open the peer inboxes of all ranks of ``comm`` (one process per GPU on one box) so that
        pmc_mgs_sweep can sum its coefficients across the GPUs inside the kernel.  Collective."""
        if self._peer_comm is not None:
            if self._peer_comm != comm:
                raise PmcError('peer inboxes are already connected to a different process group')
            return
        if comm.size > PMC_MAX_PEERS:
            raise PmcError(f'in-kernel sums support at most {PMC_MAX_PEERS} ranks')
        handle = ctypes.create_string_buffer(PMC_IPC_HANDLE_BYTES)
        self.call('pmc_peer_create', comm.rank, comm.size, ctypes.cast(handle, _P))
        blobs = comm.allgather_object(handle.raw)
        table = ctypes.create_string_buffer(b''.join(blobs), PMC_IPC_HANDLE_BYTES * comm.size)
        self.call('pmc_peer_connect', ctypes.cast(table, _P))
        comm.barrier()                 # nobody posts before everybody has connected
        self._peer_comm = comm

    def attach_comm(self, comm):
        """Give the library its own NCCL communicator over the ranks of ``comm`` (collective; the 128-byte NCCL id is
        created by rank 0 and handed round through ``comm``).  Afterwards :meth:`allreduce_sum` adds device buffers
        in place on the library's stream.  Returns whether the native path is in place on ALL ranks."""
        if self._native_comm is not None:
            return self._native_comm == (comm.rank, comm.size)
        ident = None
        if comm.rank == 0:
            buf = ctypes.create_string_buffer(PMC_COMM_ID_BYTES)
            if self.lib.pmc_comm_unique_id(ctypes.cast(buf, _P)) == PMC_OK:
                ident = buf.raw
        ident = comm.broadcast_object(ident)
        ok = False
        if ident is not None:
            buf = ctypes.create_string_buffer(ident, PMC_COMM_ID_BYTES)
            ok = self.lib.pmc_comm_init(self.handle, comm.rank, comm.size, ctypes.cast(buf, _P)) == PMC_OK
        ok = all(comm.allgather_object(bool(ok)))
        if not ok:
            self.lib.pmc_comm_destroy(self.handle)
            return False
        self._native_comm = (comm.rank, comm.size)
        return True

    def allreduce_sum(self, tensor):
        """In-place sum of a device fp64 tensor over the ranks of the attached communicator (ncclAllReduce on the
        library's stream, ordered behind the kernel that produced the data)."""
        self.call('pmc_allreduce_sum', tensor.data_ptr(), tensor.numel(), 0)
        return tensor

    def set_option(self, name, value):
        """Kernel tuning switch (``gram_opts``, ``gram_waves``, ``csr_variant``, ``l2_keep``)."""
        self.flush()
        self.call('pmc_ctx_set_option', name.encode(), int(value))

    def get_option(self, name):
        out = c_int(0)
        self.call('pmc_ctx_get_option', name.encode(), ctypes.byref(out))
        return int(out.value)

    KERNEL_KINDS = {0: 'gram', 1: 'lincomb', 2: 'spmm'}

    def kernel_times(self):
        """Per-launch times (ms) of the hot kernels recorded since the last call, as ``[(kind, ms), ...]`` in launch
        order (``set_option('kernel_timing', 1)`` switches the recording on); waits for those launches."""
        cap = 1024
        kinds, ms, count = np.zeros(cap, dtype=np.int32), np.zeros(cap), _I32(0)
        self.call('pmc_ctx_kernel_times', kinds.ctypes.data, ms.ctypes.data, cap, ctypes.byref(count))
        return [(self.KERNEL_KINDS.get(int(kinds[i]), str(int(kinds[i]))), float(ms[i])) for i in range(count.value)]

    def measure_fp64_peak(self, use_dmma=True, ms=200.0):
        out = np.zeros(1)
        self.call('pmc_measure_fp64_peak', int(use_dmma), float(ms), out.ctypes.data)   # 0 DFMA, 1 DMMA, 2/3 mixed
        return float(out[0])


_contexts = {}


def get_context(device) -> DeviceContext:
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = DeviceContext(device)
    return ctx
