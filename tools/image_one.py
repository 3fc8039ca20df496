"""The YOLOv3 image layer (3 -> 32, 3x3, 416 x 416, batch 64) through tnn_conv2d_forward_split, for ncu / timing.
usage: image_one.py [bf16x3|tf32|3xtf32] [rows|planes|both]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev

MATH = {"3xtf32": 1, "tf32": 2, "bf16x3": 3}
math = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
what = sys.argv[2] if len(sys.argv) > 2 else "planes"
B, C, H, F, k = int(os.environ.get("B", "64")), 3, 416, 32, 3
cu = tensor.initCUDART(0)
cu.set_math_mode(MATH[math])
rng = np.random.default_rng(0)
x = rng.random((B, C, H, H), dtype=np.float32)
w = rng.uniform(-0.2, 0.2, (F, C * k * k)).astype(np.float32)
dx, dw, db = Dev(cu, x), Dev(cu, w), Dev(cu, np.zeros(F, np.float32))
M = B * H * H
do, dsp = Dev(cu, n=M * F), Dev(cu, n=M * F)
desc = cu.conv_desc(B, C, H, H, F, k, k, 1, 1, 1, 1)
run = lambda: cu.conv2dForwardSplit(desc, dx.ptr, 0, 0, 0, 0, dw.ptr, None, db.ptr, 9, None, 0, 0,
                                    do.ptr if what != "planes" else None, 0, F, 1,
                                    dsp.ptr if what != "rows" else None, 0, F, M * F)
for _ in range(2):
    run()
cu.finish()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    run()
cu.finish()
dt = (time.perf_counter() - t0) / n
print(f"image layer B{B}: {dt*1e3:.3f} ms ({math}, {what}, {cu.lastConvKernel()})  out bytes {M*F*4*(2 if what=='both' else 1)/1e9:.2f} GB")
