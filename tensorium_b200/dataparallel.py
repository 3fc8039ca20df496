"""Data-parallel training plumbing: one process per GPU, the batch sharded across ranks,
and ONE all-reduce per step over a flat bucket that aliases every layer's update buffer
(weight_updates / bias_updates / scale_updates — the buffers TConvolutionalLayer.updateGPU consumes,
source/nConvolutionLayer.pas:1294-1330).  The reference is single-device; this is the only exchange step
the path has.

Two buckets with the same layout:

* NativeBucket — device memory from the C ABI (createDeviceBuffer) and the library's own NCCL entry points
  (tnn_comm_init_rank / tnn_allreduce_sum): exactly what the FreePascal host would call.  The collective runs on the
  backend's stream, so TNNet.enableGraph captures forward + backward + all-reduce + update into ONE CUDA graph.
  torch.distributed is used only to hand rank 0's 128-byte NCCL id to the other ranks.
* GradientBucket — a flat torch tensor + torch.distributed.all_reduce: the host-logic twin used by the world-size-2
  gloo tests on CPU (tests/test_train_cpu.py).

The update buffers double as the momentum state (update: w += rate*u; u *= momentum).  Summing them across
R replicas every step would multiply the carried momentum by R, so updateGPU keeps momentum/R in each
replica's buffer (TUpdateArgs.replicas) — the next SUM then restores momentum * (previous summed update), and
the weights of all replicas stay bit-identical because every replica applies the same summed buffer.

Batch-norm under batch sharding: the reference computes the statistics over the whole mini-batch
(source/nbaselayer.pas:352, source/ntensors.pas:9102-9177).  `sync_bn=True` reproduces that across replicas
(SyncStats below: the per-channel sums are all-reduced, mean first and then the squared deviations from the GLOBAL
mean — the reference's two-pass formula — and the three backward sums likewise); `sync_bn=False` keeps per-replica
statistics (standard data parallelism, different numbers from the single-device run at the same global batch).
"""
from __future__ import annotations

import numpy as np

ALIGN = 64  # elements: every slice starts on a 256-byte boundary


def _layout(sizes):
    offsets, off = [], 0
    for n in sizes:
        offsets.append(off)
        off += (int(n) + ALIGN - 1) // ALIGN * ALIGN
    return offsets, max(off, 1)


class GradientBucket:
    """A flat float32 torch tensor carved into per-tensor slices (256-byte aligned); torch.distributed all-reduce."""

    ALIGN = ALIGN

    def __init__(self, sizes, device):
        import torch
        import torch.distributed as dist
        self.offsets, total = _layout(sizes)
        self.sizes = [int(n) for n in sizes]
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.capturable = False

    def view(self, i):
        return self.flat[self.offsets[i]:self.offsets[i] + self.sizes[i]]

    def ptr(self, i) -> int:
        return self.flat.data_ptr() + 4 * self.offsets[i]

    def all_reduce(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def exchange_unique_id(cu, rank: int, world: int) -> bytes:
    """Rank 0's NCCL id reaches the other ranks through whatever the host has: here torch.distributed's store."""
    import torch.distributed as dist
    box = [cu.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class NativeBucket:
    """The same layout in one C-ABI device buffer, all-reduced by the library's NCCL entry point on the backend's
    stream (capturable into the step's CUDA graph)."""

    def __init__(self, cu, sizes, comm=None, world=1):
        self.cu = cu
        self.offsets, self.total = _layout(sizes)
        self.sizes = [int(n) for n in sizes]
        self.base = cu.createDeviceBuffer(self.total * 4)
        cu.fill(self.total, self.base, 0, 0.0, 1)
        self.comm, self.world = comm, world
        self.capturable = True

    def ptr(self, i) -> int:
        return self.base + 4 * self.offsets[i]

    def all_reduce(self):
        if self.world > 1:
            self.cu.allreduceSum(self.comm, self.base, 0, self.total)

    def read(self) -> np.ndarray:
        out = np.empty(self.total, np.float32)
        self.cu.readBuffer(self.base, out.nbytes, out)
        return out


class SyncStats:
    """Per-channel batch-norm sums shared across replicas (SyncBN): small all-reduces on the backend's stream."""

    def __init__(self, cu, comm, world):
        self.cu, self.comm, self.world = cu, comm, world

    def all_reduce(self, dev_ptr, n):
        if self.world > 1:
            self.cu.allreduceSum(self.comm, dev_ptr, 0, n)


_COMM = {}


def native_comm(cu):
    """(comm, world, rank) of this process: one NCCL communicator per backend instance, created on first use."""
    import torch.distributed as dist
    key = id(cu)
    if key not in _COMM:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            world, rank = dist.get_world_size(), dist.get_rank()
            uid = exchange_unique_id(cu, rank, world)
            _COMM[key] = (cu.commInitRank(world, rank, uid), world, rank)
        else:
            _COMM[key] = (None, 1, 0)
    return _COMM[key]


def attach(net, device=None, sync_bn=False):
    """Moves every update buffer of `net` into one bucket and tells the net how many replicas there are.  On a GPU the
    bucket is native (C ABI + the library's NCCL); the all-reduce is ordered on the backend's stream after the
    backward kernels and before the update kernels."""
    from .tensor import cuda
    cu = cuda()
    grads = net.gradients()
    comm, world, _ = native_comm(cu)
    bucket = NativeBucket(cu, [g.Size() for g in grads], comm, world)
    for i, g in enumerate(grads):
        if not g.wasGPU():
            g.pushToDevice()
        g.rebind(bucket.ptr(i))
        g.setCUDA()
    net.bucket = bucket
    net.replicas = world
    if sync_bn and world > 1:
        stats = SyncStats(cu, comm, world)
        for l in net.layers:
            if getattr(l, "isBatchNormalized", False):
                l.sync_stats = stats
    return bucket


def broadcast_parameters(net, src=0):
    """Replicas must start from the same weights: rank `src`'s device copies are broadcast over the library's NCCL."""
    from .tensor import cuda
    cu = cuda()
    comm, world, _ = native_comm(cu)
    if world == 1:
        return
    for l in net.layers:
        for name in ("weights", "biases", "scales", "rolling_mean", "rolling_variance"):
            t = getattr(l, name, None)
            if t is None:
                continue
            if not t.wasGPU():
                t.pushToDevice()
            cu.broadcast(comm, t.devData, 0, t.Size(), src)
            t.setCUDA()
    cu.finish()
