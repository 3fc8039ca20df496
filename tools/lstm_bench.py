"""SURVEY config 5 (char-LSTM step): TLSTMLayer forward, steps=128, batch=256, through the C ABI.
Prints one JSON line: tokens/s (= steps*batch sequence positions per second) for the reference's operator order
(`--unfused`) or the up-front input projections (default), eager or as one captured CUDA graph (`--graph`)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import tensor  # noqa: E402
from tensorium_b200.layers import TNNetState  # noqa: E402
from tensorium_b200.rnn_layers import TLSTMLayer  # noqa: E402
from tensorium_b200.tensor import TSingleTensor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--inputs", type=int, default=256)
ap.add_argument("--hidden", type=int, default=1024)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--graph", action="store_true")
ap.add_argument("--math", default="3xtf32")
a = ap.parse_args()
cu = tensor.initCUDART(0)
cu.set_math_mode({"fp32": 0, "3xtf32": 1, "tf32": 2}[a.math])
x = np.random.default_rng(0).uniform(-1, 1, (a.steps * a.batch, a.inputs)).astype(np.float32)
l = TLSTMLayer(a.steps * a.batch, a.inputs, a.hidden, a.steps, fused=not a.unfused)
st = TNNetState(input=TSingleTensor(x.shape, a.steps * a.batch, data=x))
n0 = cu.launch_count()
l.forwardGPU(st)
cu.finish()
launches = cu.launch_count() - n0
run = lambda: l.forwardGPU(st)
if a.graph:
    l.forwardGPU(st)
    cu.finish()
    cu.beginCapture()
    l.forwardGPU(st)
    g = cu.endCapture()
    run = lambda: cu.launchGraph(g)
run()
cu.finish()
t0 = time.perf_counter()
for _ in range(a.iters):
    run()
cu.finish()
dt = (time.perf_counter() - t0) / a.iters
flops = 2.0 * a.steps * a.batch * a.hidden * (a.inputs + a.hidden) * 4
print(json.dumps({"metric": "lstm_forward_tokens_per_s", "value": a.steps * a.batch / dt, "unit": "tokens/s",
                  "ms_per_sequence_batch": dt * 1e3, "tflops": flops / dt / 1e12, "steps": a.steps, "batch": a.batch,
                  "inputs": a.inputs, "hidden": a.hidden, "math": a.math, "fused_projections": not a.unfused,
                  "cuda_graph": a.graph, "launches_per_forward": launches}))
