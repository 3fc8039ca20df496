"""Launch each HBM-bound kernel of the path a few times at roofline sizes (tensors far larger than L2) — the command
line ncu wraps for the memory-bound captures of tools/profile_round2.sh.  usage: membound_driver.py [op ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import tensor  # noqa: E402

ops = sys.argv[1:] or ["rows_add", "activate", "bn_fwd", "bn_bwd", "sgd", "upsample", "nchw2nhwc"]
cu = tensor.initCUDART(0)
B, C, HW = 64, 64, 208 * 208
n = B * C * HW
a, b, c = (torch.rand(n, device="cuda") for _ in range(3))
Bb, Cb, HWb = 2048, 32, 1024
nb = Bb * Cb * HWb
x, xn = torch.empty(nb, device="cuda"), torch.empty(nb, device="cuda")
vec = [torch.ones(Cb, device="cuda") for _ in range(9)]
p = [v.data_ptr() for v in vec]
torch.cuda.synchronize()
for _ in range(3):
    if "rows_add" in ops:
        cu.rowsAdd(B * HW, C, a.data_ptr(), C, b.data_ptr(), C, c.data_ptr(), C)
    if "activate" in ops:
        cu.ActivateArray(n, a.data_ptr(), 0, 9)
    if "bn_fwd" in ops:
        cu.batchNormForwardTraining(Bb, Cb, HWb, a.data_ptr(), x.data_ptr(), xn.data_ptr(), p[0], p[1], p[2], p[3], 0.1, p[4],
                                    p[5], 1)
    if "bn_bwd" in ops:
        cu.batchNormBackward(Bb, Cb, HWb, a.data_ptr(), 1, b.data_ptr(), x.data_ptr(), xn.data_ptr(), p[0], p[1], p[4], p[6],
                             p[7], p[8])
    if "sgd" in ops:
        cu.sgdUpdate(1 << 27, a.data_ptr(), b.data_ptr(), 1e-6, 0.9, -1e-6)
    if "upsample" in ops:
        cu.upSampleNhwc(64, 256, 26, 26, 2, 1.0, a.data_ptr(), 256, c.data_ptr(), 256)
    if "nchw2nhwc" in ops:
        cu.nchwToNhwc(64, 64, HW, a.data_ptr(), c.data_ptr(), 64)
cu.finish()
print("ok", ops)
