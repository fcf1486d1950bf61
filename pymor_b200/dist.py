"""Row (DOF) sharding of snapshot arrays across the GPUs of one box.

This mirrors ``MPIVectorArrayAutoComm`` (reference: src/pymor/vectorarrays/mpi.py:316-445): every
rank holds the rows ``[offset, offset + local_dim)`` of every vector, ``len`` is global, reductions
are local partial results summed over ranks.  The mechanism differs: instead of a rank-0
controller broadcasting pickled calls to an event loop (src/pymor/tools/mpi.py:111-158) every rank
runs the same Python program (SPMD, one process per GPU, launched by torchrun) and the k1 x k2 / k
partials are summed with one NCCL all-reduce over NVLink, so all ranks see identical host scalars
and take identical control-flow decisions in ``gram_schmidt`` / ``pod``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


class SingleRank:
    """No communication (one GPU)."""
    rank = 0
    size = 1

    def allreduce_sum_(self, tensor):
        return tensor

    def allgather(self, tensor):
        return [tensor]

    def barrier(self):
        pass

    def allgather_object(self, obj):
        return [obj]

    def broadcast_object(self, obj, src=0):
        return obj

    def attach(self, ctx):
        pass

    def exchange(self, sends, recvs):
        assert not sends and not recvs

    def __eq__(self, other):
        return isinstance(other, SingleRank)

    def __hash__(self):
        return 0


class ProcessGroupComm:
    """torch.distributed process group (NCCL for device tensors, gloo in the CPU tests)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._native = {}            # DeviceContext -> whether the library-side NCCL communicator is in place

    def attach(self, ctx):
        """Collective, called once per (communicator, device context) by ``CudaVectorSpace``: let the LIBRARY own
        the all-reduce of the k x k / k partials (``pmc_allreduce_sum``: NCCL behind the C ABI, on the library's
        stream) instead of ``torch.distributed`` (which stays the bootstrap channel for the NCCL id and carries the
        halo exchange).  ``PMC_NATIVE_COMM=0`` keeps the all-reduce in ``torch.distributed`` as well; the gloo process
        groups of the CPU tests always do."""
        if ctx in self._native:
            return
        use = (os.environ.get('PMC_NATIVE_COMM', '1') != '0' and ctx.device.type == 'cuda'
               and dist.get_backend(self.group) == 'nccl' and hasattr(ctx.lib, 'pmc_allreduce_sum'))
        self._native[ctx] = bool(use) and ctx.attach_comm(self)

    def allreduce_sum_(self, tensor):
        if tensor.is_cuda and tensor.dtype == torch.float64 and tensor.is_contiguous():
            for ctx, native in self._native.items():
                if native and ctx.device == tensor.device:
                    return ctx.allreduce_sum(tensor)
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def broadcast_object(self, obj, src=0):
        box = [obj]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                                   group=self.group)
        return box[0]

    def allgather(self, tensor):
        out = [torch.empty_like(tensor) for _ in range(self.size)]
        dist.all_gather(out, tensor, group=self.group)
        return out

    def barrier(self):
        dist.barrier(group=self.group)

    def allgather_object(self, obj):
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def exchange(self, sends, recvs):
        """Neighbour exchange: ``sends`` / ``recvs`` map peer rank -> contiguous tensor.  Point-to-point
        (NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests); every rank must call it."""
        ops = [dist.P2POp(dist.irecv, t, p, group=self.group) for p, t in sorted(recvs.items())]
        ops += [dist.P2POp(dist.isend, t, p, group=self.group) for p, t in sorted(sends.items())]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def __eq__(self, other):
        return isinstance(other, ProcessGroupComm) and other.group is self.group

    def __hash__(self):
        return hash(('pg', id(self.group)))


def partition(dim, size):
    """Even row partition: local dims and offsets (cumulative local dims, mpi.py:380-384)."""
    base, rem = divmod(int(dim), int(size))
    dims = np.array([base + (1 if r < rem else 0) for r in range(size)], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(dims)[:-1])).astype(np.int64)
    return dims, offsets


def merge_amax(inds, vals):
    """Per-column winner over shards; the lowest shard wins ties (mpi.py:341-347).

    ``inds`` are already global row indices, shape (size, k); ``vals`` (size, k)."""
    inds, vals = np.asarray(inds), np.asarray(vals)
    win = np.argmax(vals, axis=0)
    cols = np.arange(vals.shape[1])
    return inds[win, cols], vals[win, cols]


class SharedHostReducer:
    """Sum of tiny per-rank partial results through a pinned shared-memory segment.

    ``pairwise_inner`` / ``norm`` inside Gram-Schmidt return one double that the HOST needs (it drives
    the control flow and fills ``R``).  Going device -> NCCL all-reduce -> device -> host costs several
    launch/sync latencies per pair-step; the reference's own MPI path does a ``Gather`` to rank 0 plus a
    host sum (src/pymor/vectorarrays/mpi.py:394-416).  Here every rank's reduction kernel writes its
    partial straight into its slot of a POSIX shared-memory segment that all ranks of the box have
    pinned (``cudaHostRegister``); after its own stream sync a rank publishes a sequence number, spins
    until the other slots carry the same number and adds the slots in rank order -- every rank gets the
    bitwise identical sum, with no collective launch.  Two slot generations (parity of the sequence
    number) make it safe for a fast rank to start the next reduction while a slow one still reads.
    """

    HEADER = 8          # doubles reserved per slot for the sequence number (64-byte aligned payload)

    def __init__(self, comm, ctx, capacity=2048):
        from multiprocessing import shared_memory
        self.comm, self.ctx, self.capacity = comm, ctx, int(capacity)
        self.slot = self.HEADER + self.capacity
        nbytes = comm.size * 2 * self.slot * 8
        name = None
        if comm.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            np.ndarray((nbytes // 8,), dtype=np.float64, buffer=self._shm.buf)[:] = 0.0
            name = self._shm.name
        name = comm.allgather_object(name)[0]
        if comm.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name)
        self._data = np.ndarray((comm.size, 2, self.slot), dtype=np.float64, buffer=self._shm.buf)
        self._seqs = self._data.view(np.int64)
        self._host_ptr = self._data.ctypes.data
        self._dev_ptr = ctx.register_host(self._host_ptr, nbytes)
        self._seq = 0
        comm.barrier()
        from multiprocessing import resource_tracker
        try:                            # every rank keeps its mapping; the name goes away right now
            resource_tracker.unregister(self._shm._name, 'shared_memory')
        except Exception:
            pass
        if comm.rank == 0:
            try:
                import os
                os.unlink('/dev/shm/' + self._shm.name.lstrip('/'))
            except OSError:
                pass

    def close(self):
        if self._shm is None:
            return
        try:
            self.ctx.unregister_host(self._host_ptr)
        except Exception:
            pass
        self._data = self._seqs = None
        try:
            self._shm.close()
        except Exception:
            pass
        self._shm = None

    def target(self):
        """Device-accessible address of this rank's payload for the NEXT reduction."""
        gen = (self._seq + 1) & 1
        return self._dev_ptr + ((self.comm.rank * 2 + gen) * self.slot + self.HEADER) * 8

    def finish(self, count, timeout=120.0):
        """Call after the kernel's stream has been synchronised: publish, wait for the peers, sum."""
        import time
        self._seq += 1
        seq, gen = self._seq, self._seq & 1
        self._seqs[self.comm.rank, gen, 0] = seq
        total = np.zeros(count)
        deadline = None
        for r in range(self.comm.size):
            spins = 0
            while self._seqs[r, gen, 0] < seq:
                spins += 1
                if spins & 0xffff == 0:
                    deadline = deadline or time.monotonic() + timeout
                    if time.monotonic() > deadline:
                        raise RuntimeError(f'rank {self.comm.rank}: timed out waiting for rank {r} in host reduction')
            total += self._data[r, gen, self.HEADER:self.HEADER + count]
        return total


_reducers = {}


def get_host_reducer(comm, ctx, capacity):
    """One reducer per (process group, device): spaces come and go, the pinned segment stays."""
    key = (id(getattr(comm, 'group', None)), str(ctx.device))
    red = _reducers.get(key)
    if red is None:
        red = _reducers[key] = SharedHostReducer(comm, ctx, capacity)
    return red


def _close_reducers():
    for red in _reducers.values():
        red.close()
    _reducers.clear()


import atexit  # noqa: E402

atexit.register(_close_reducers)
