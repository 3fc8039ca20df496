"""world_size-2 gloo tests (CPU) of the N>1 host logic of bench.py: batch sharding with no data-path collective,
max-over-ranks timing, whole-job throughput, and the reference arm running on rank 0 only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    got = bench.reduce_max_ms(ms, world, device="cpu")
    x = bench.synthetic_images(2, bench.rank_seed(rank))
    out[rank] = (got, float(x.mean()), bench.aggregate_images_per_s(world, 3, got))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_max_over_ranks_gloo():
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == res[1][0] == 15.0                 # MAX over ranks, same on every rank
    assert res[0][1] != res[1][1]                         # ranks draw different images (independent shards)
    assert res[0][2] == pytest.approx(64 * 2 * 3 / 0.015)  # whole-job images/s = all ranks' images / max time


def test_synthetic_images_are_reproducible():
    sys.path.insert(0, ROOT)
    import bench
    a, b = bench.synthetic_images(1, 7), bench.synthetic_images(1, 7)
    assert a.dtype == np.float32 and a.shape == (1, 3, 416, 416) and np.array_equal(a, b)
    assert 0.0 <= a.min() and a.max() < 1.0


def test_reference_arm_runs_on_rank0_only():
    # rank 1 of a 2-rank launch must exit 0 without printing or working
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
