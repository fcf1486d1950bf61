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

    def allreduce_sum_(self, tensor):
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

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
