"""How long does the HOST take to enqueue a CIFAR training step vs how long the GPU takes to run it?"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorium_b200 import models, tensor
cu = tensor.initCUDART(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = models.deepCIFAR10(B, seed=1, learningRate=0.01)
rng = np.random.default_rng(0)
x = rng.uniform(0, 1, (B, 3, 32, 32)).astype(np.float32)
t = np.zeros((B, 10), np.float32); t[np.arange(B), rng.integers(0, 10, B)] = 1
for _ in range(5):
    net.trainBatch(x, t)
cu.finish()
n = 30
t0 = time.perf_counter()
for _ in range(n):
    net.trainBatch(x, t)
t1 = time.perf_counter()
cu.finish()
t2 = time.perf_counter()
print(f"batch {B}: host enqueue {1e3*(t1-t0)/n:.3f} ms/step, total {1e3*(t2-t0)/n:.3f} ms/step")
