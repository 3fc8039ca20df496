"""CIFAR-10 training throughput of deepCIFAR10 (BASELINE config 3) on synthetic data.

  python tools/cifar_train.py --batch 128 --steps 50
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/cifar_train.py --batch 128

One training iteration = host batch -> device, forward, backward, all-reduce of the update bucket when world > 1
(the library's own NCCL entry point, captured into the step's CUDA graph), SGD update; the loss is read back every
--loss-every iterations.  Prints one JSON line on rank 0.  bench.py imports train_throughput() for its secondary
`cifar10_train` object.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MATH = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16x3": 3}


def train_throughput(cu, stream, batch, steps=50, warmup=5, rank=0, world=1, sync_bn=False, fused=True, graph=True,
                     loss_every=10, seed=1):
    """deepCIFAR10 (nmodels.pas:82-111) on synthetic batches of `batch` images per GPU.  `stream`: the torch stream
    the backend launches on (events are recorded on it).  Returns a dict (max-over-ranks time)."""
    import torch
    import torch.distributed as dist
    from tensorium_b200 import dataparallel, models
    net = models.deepCIFAR10(batch, fused=fused, seed=seed, learningRate=0.01)
    dataparallel.broadcast_parameters(net)
    if world > 1:
        dataparallel.attach(net, sync_bn=sync_bn)
    rng = np.random.default_rng(100 + rank)
    x = rng.uniform(0, 1, (4, batch, 3, 32, 32)).astype(np.float32)
    t = np.zeros((4, batch, 10), np.float32)
    for k in range(4):
        t[k, np.arange(batch), rng.integers(0, 10, batch)] = 1
    costs = []

    def step(i):
        net.trainBatch(x[i % 4], t[i % 4])
        if loss_every and i % loss_every == 0:
            costs.append(net.cost() / batch)

    for i in range(warmup):
        step(i)
    if graph and fused:
        net.enableGraph()
        for i in range(2):
            step(i)
    l0 = cu.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for i in range(steps):
        step(warmup + i)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = (cu.launch_count() - l0) / steps
    if world > 1:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    # the exchange step alone (all-reduce of the whole update bucket), timed back to back on the backend's stream
    allreduce_us = None
    if world > 1:
        for _ in range(3):
            net.bucket.all_reduce()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(20):
            net.bucket.all_reduce()
        e1.record(stream)
        torch.cuda.synchronize()
        allreduce_us = e0.elapsed_time(e1) / 20 * 1e3
        cu.fill(net.bucket.total, net.bucket.base, 0, 0.0, 1)
    # replicas must hold bit-identical weights after the summed updates
    in_sync = True
    if world > 1:
        sums = []
        for l in net.layers:
            if getattr(l, "trainable", False):
                sums.append(float(np.abs(l.weights.pullFromDevice()).sum(dtype=np.float64)))
                l.weights.setCUDA()
        mine = torch.tensor(sums, dtype=torch.float64, device="cuda")
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        in_sync = all(bool(torch.equal(allv[0], v)) for v in allv)
    bucket_bytes = net.bucket.total * 4 if world > 1 else sum(g.Size() for g in net.gradients()) * 4
    return {"images_per_s": batch * world * steps / (ms / 1e3), "ms_per_step": ms / steps, "per_gpu_batch": batch,
            "global_batch": batch * world, "n_gpus": world, "steps": steps, "launches_per_step": launches,
            "cuda_graph": bool(graph and fused), "sync_bn": bool(sync_bn and world > 1), "allreduce_us": allreduce_us,
            "bucket_bytes": bucket_bytes, "replicas_in_sync": in_sync, "loss_first": costs[0] if costs else None,
            "loss_last": costs[-1] if costs else None}


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--math", default="3xtf32", choices=list(MATH))
    ap.add_argument("--unfused", action="store_true", help="issue the reference's own cuda.<op> sequence")
    ap.add_argument("--loss-every", type=int, default=10)
    ap.add_argument("--sync-bn", action="store_true", help="batch-norm statistics over the global batch (SyncBN)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying "
                                                              "the captured CUDA graph of the step")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tensorium_b200 import tensor
    cu = tensor.initCUDART(local)
    stream = torch.cuda.Stream()
    cu.set_stream(stream.cuda_stream)
    cu.set_math_mode(MATH[a.math])
    with torch.cuda.stream(stream):
        r = train_throughput(cu, stream, a.batch, a.steps, a.warmup, rank, world, a.sync_bn, not a.unfused,
                             not a.no_graph, a.loss_every)
    if rank == 0:
        r.update(metric="cifar10_train_images_per_s", value=r["images_per_s"], unit="images/s", math=a.math,
                 fused=not a.unfused, data="synthetic", e2e="host batch copied in every step")
        print(json.dumps(r))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
