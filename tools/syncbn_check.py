"""N replicas (batch sharded, NCCL all-reduce of the update bucket through the C ABI) against ONE device at the same
global batch, with batch-norm on: with SyncBN the replicas follow the single-device run (the reference's semantics:
statistics over the whole mini-batch, nbaselayer.pas:352); without it they drift apart from it.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/syncbn_check.py
Prints one JSON line on rank 0 (max relative weight difference after `--steps` SGD steps, per mode)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def weights_of(net):
    out = []
    for l in net.layers:
        if getattr(l, "trainable", False):
            for name in ("weights", "biases", "scales", "rolling_mean", "rolling_variance"):
                t = getattr(l, name, None)
                if t is not None:
                    out.append(t.pullFromDevice().copy().reshape(-1))
                    t.setCUDA()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tensorium_b200 import dataparallel, models, tensor
    cu = tensor.initCUDART(local)
    B = a.global_batch
    b = B // world
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 1, (a.steps, B, 3, 32, 32)).astype(np.float32)
    T = np.zeros((a.steps, B, 10), np.float32)
    for s in range(a.steps):
        T[s, np.arange(B), rng.integers(0, 10, B)] = 1
    res = {}
    trace = {"sync_bn": [], "local_bn": []}
    for mode in ("sync_bn", "local_bn"):
        net = models.deepCIFAR10(b, seed=1, dropout=0.0, learningRate=0.01)
        dataparallel.broadcast_parameters(net)
        dataparallel.attach(net, sync_bn=(mode == "sync_bn"))
        for s in range(a.steps):
            net.trainBatch(X[s, rank * b:(rank + 1) * b], T[s, rank * b:(rank + 1) * b])
            cu.finish()
            trace[mode].append(weights_of(net))
        res[mode] = weights_of(net)
        # replicas identical?
        mine = torch.tensor([float(np.abs(w).sum(dtype=np.float64)) for w in res[mode]], dtype=torch.float64, device="cuda")
        allv = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(allv, mine)
        res[mode + "_replicas_identical"] = all(bool(torch.equal(allv[0], v)) for v in allv)
    if rank == 0:
        single = models.deepCIFAR10(B, seed=1, dropout=0.0, learningRate=0.01)
        per_step = {"sync_bn": [], "local_bn": []}
        for s in range(a.steps):
            single.trainBatch(X[s], T[s])
            cu.finish()
            ref_s = weights_of(single)
            for mode in per_step:
                per_step[mode].append(max(float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-12))
                                          for g, r in zip(trace[mode][s], ref_s)))
        ref = weights_of(single)
        out = {"global_batch": B, "replicas": world, "steps": a.steps, "max_rel_diff_after_each_step": per_step}
        for mode in ("sync_bn", "local_bn"):
            worst = 0.0
            for g, r in zip(res[mode], ref):
                worst = max(worst, float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-12)))
            out[mode + "_max_rel_diff_vs_single_device"] = worst
            out[mode + "_replicas_identical"] = res[mode + "_replicas_identical"]
        print(json.dumps(out))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
