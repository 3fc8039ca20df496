"""Run-to-run reproducibility of a training step: two identical deepCIFAR10 replicas, same batches and dropout
seeds, stepped eagerly side by side; prints the per-layer weight difference after every step.  The split
reductions accumulate with fp32 atomics, so differences of a few ulp are expected; anything larger is a mask flip
(ReLU / max-pool route) downstream of such an ulp, or a bug."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import models  # noqa: E402
from tensorium_b200.tensor import initCUDART  # noqa: E402
from tensorium_b200.train_layers import TDropoutLayer  # noqa: E402


def batch(B, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (B, 3, 32, 32)).astype(np.float32)
    t = np.zeros((B, 10), np.float32)
    t[np.arange(B), rng.integers(0, 10, B)] = 1
    return x, t


def main():
    initCUDART(0)
    B = int(os.environ.get("B", 16))
    steps = int(os.environ.get("STEPS", 6))
    nets = [models.deepCIFAR10(B, seed=21, learningRate=0.01) for _ in range(2)]
    for i in range(steps):
        x, t = batch(B, 300 + i)
        for net in nets:
            for li, l in enumerate(net.layers):
                if isinstance(l, TDropoutLayer):
                    l.seed = 1000 * i + li
            net.trainBatch(x, t)
        errs = []
        for la, lb in zip(nets[0].layers, nets[1].layers):
            if getattr(la, "trainable", False):
                a, b = la.weights.pullFromDevice(), lb.weights.pullFromDevice()
                la.weights.setCUDA(); lb.weights.setCUDA()
                errs.append(float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)))
        if max(errs) > 1e-5 and not os.environ.get("QUIET"):
            for li, (la, lb) in enumerate(zip(nets[0].layers, nets[1].layers)):
                for name in ("output", "delta", "indexes", "x_norm", "mean", "variance", "weight_updates", "bias_updates",
                             "scale_updates", "rand"):
                    ta, tb = getattr(la, name, None), getattr(lb, name, None)
                    if ta is None or tb is None or not getattr(ta, "devData", None):
                        continue
                    a, b = ta.pullFromDevice().copy(), tb.pullFromDevice().copy()
                    ta.setCUDA(); tb.setCUDA()
                    if a.dtype.kind == "i":
                        nd = int((a != b).sum())
                        if nd:
                            print(f"   layer {li} {type(la).__name__}.{name}: {nd} of {a.size} indexes differ")
                    else:
                        d = float(np.abs(a - b).max())
                        if d > 1e-5 * max(float(np.abs(b).max()), 1e-30):
                            k = int(np.abs(a - b).argmax())
                            print(f"   layer {li} {type(la).__name__}.{name}: max diff {d:.3e} (scale {np.abs(b).max():.3e}) at {k}: "
                                  f"{a.flat[k]!r} vs {b.flat[k]!r}; {int((a != b).sum())} of {a.size} differ")
            os.environ["QUIET"] = "1"
        print(f"step {i} cost {nets[0].cost():.6f} {nets[1].cost():.6f}  " + " ".join(f"{e:.1e}" for e in errs), flush=True)


if __name__ == "__main__":
    main()
