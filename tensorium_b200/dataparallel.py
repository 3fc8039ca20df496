"""Data-parallel training plumbing: one process per GPU, the batch sharded across ranks,
and ONE all-reduce per step over a flat bucket that aliases every layer's update buffer
(weight_updates / bias_updates / scale_updates — the buffers TConvolutionalLayer.updateGPU consumes,
source/nConvolutionLayer.pas:1294-1330).  The reference is single-device; this is the only exchange step
the path has.

The update buffers double as the momentum state (update: w += rate*u; u *= momentum).  Summing them across
R replicas every step would multiply the carried momentum by R, so updateGPU keeps momentum/R in each
replica's buffer (TUpdateArgs.replicas) — the next SUM then restores momentum * (previous summed update), and
the weights of all replicas stay bit-identical because every replica applies the same summed buffer.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class GradientBucket:
    """A flat float32 torch tensor carved into per-tensor slices (256-byte aligned)."""

    ALIGN = 64  # elements

    def __init__(self, sizes, device):
        self.offsets = []
        off = 0
        for n in sizes:
            self.offsets.append(off)
            off += (int(n) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.sizes = [int(n) for n in sizes]
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=device)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def view(self, i) -> torch.Tensor:
        return self.flat[self.offsets[i]:self.offsets[i] + self.sizes[i]]

    def ptr(self, i) -> int:
        return self.flat.data_ptr() + 4 * self.offsets[i]

    def all_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def attach(net, device=None) -> GradientBucket:
    """Moves every update buffer of `net` into one bucket on `device` and tells the net how many replicas
    there are.  The backend must be launching on torch's current stream (TNNCuda.set_stream) so that the
    NCCL all-reduce is ordered after the backward kernels and before the update kernels."""
    grads = net.gradients()
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    bucket = GradientBucket([g.Size() for g in grads], device)
    for i, g in enumerate(grads):
        if not g.wasGPU():
            g.pushToDevice()
        g.rebind(bucket.ptr(i))
        g.setCUDA()
    net.bucket = bucket
    net.replicas = bucket.world
    return bucket


def broadcast_parameters(net, src=0):
    """Replicas must start from the same weights: broadcast rank `src`'s host copies."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for l in net.layers:
        for name in ("weights", "biases", "scales", "rolling_mean", "rolling_variance"):
            t = getattr(l, name, None)
            if t is None:
                continue
            if t.wasGPU():
                t.pullFromDevice()
            if dist.get_backend() == "nccl":
                buf = torch.from_numpy(t.Data).cuda()
                dist.broadcast(buf, src)
                t.Data[...] = buf.cpu().numpy()
            else:
                dist.broadcast(torch.from_numpy(t.Data), src)
            t.setCPU()
