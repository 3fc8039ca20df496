"""CIFAR-10 training throughput of deepCIFAR10 (BASELINE config 3) on synthetic data.

  python tools/cifar_train.py --batch 128 --steps 50
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/cifar_train.py --batch 128

One training iteration = host batch -> device, forward, backward, (all-reduce of the update bucket when
world > 1), SGD update; the loss is read back every --loss-every iterations.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--math", default="3xtf32", choices=["fp32", "3xtf32", "tf32"])
    ap.add_argument("--unfused", action="store_true", help="issue the reference's own cuda.<op> sequence")
    ap.add_argument("--loss-every", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying "
                                                              "the captured CUDA graph of the step")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tensorium_b200 import dataparallel, models, tensor
    cu = tensor.initCUDART(local)
    stream = torch.cuda.Stream()
    cu.set_stream(stream.cuda_stream)
    cu.set_math_mode({"fp32": 0, "3xtf32": 1, "tf32": 2}[a.math])
    with torch.cuda.stream(stream):
        net = models.deepCIFAR10(a.batch, fused=not a.unfused, seed=1, learningRate=0.01)
        dataparallel.broadcast_parameters(net)
        if world > 1:
            dataparallel.attach(net)
        rng = np.random.default_rng(100 + rank)
        x = rng.uniform(0, 1, (4, a.batch, 3, 32, 32)).astype(np.float32)
        t = np.zeros((4, a.batch, 10), np.float32)
        for k in range(4):
            t[k, np.arange(a.batch), rng.integers(0, 10, a.batch)] = 1
        costs = []

        def step(i):
            net.trainBatch(x[i % 4], t[i % 4])
            if a.loss_every and i % a.loss_every == 0:
                costs.append(net.cost() / a.batch)

        for i in range(a.warmup):
            step(i)
        if not a.no_graph and not a.unfused:
            net.enableGraph()
            for i in range(2):
                step(i)
        l0 = cu.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(a.steps):
            step(a.warmup + i)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
            ms = float(tms.item())
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
        torch.distributed.all_gather(allv, mine)
        in_sync = all(bool(torch.equal(allv[0], v)) for v in allv)
    if rank == 0:
        print(json.dumps({
            "metric": "cifar10_train_images_per_s", "value": a.batch * world * a.steps / (ms / 1e3), "unit": "images/s",
            "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps, "per_gpu_batch": a.batch,
            "math": a.math, "fused": not a.unfused, "cuda_graph": (not a.no_graph and not a.unfused), "launches_per_step": (cu.launch_count() - l0) / a.steps,
            "loss_first": costs[0] if costs else None, "loss_last": costs[-1] if costs else None,
            "replicas_in_sync": in_sync, "data": "synthetic", "e2e": "host batch copied in every step"}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
