"""The image layer (3 -> 32, 3x3, 416x416, batch 64): direct fp32 kernel writing channels-innermost rows vs
nchw->nhwc(pitch 4) + the TMA im2col tensor-core kernel.  usage: first_layer.py [math]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import tensor

MATH = {"3xtf32": 1, "tf32": 2, "bf16x3": 3}
math = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
cu = tensor.initCUDART(0); cu.set_stream(stream.cuda_stream); cu.set_math_mode(MATH[math])
B, C, H, F, k = 64, 3, 416, 32, 3
x = torch.rand(B * C * H * H, device="cuda")
x4 = torch.zeros(B * H * H * 4, device="cuda")
w = (torch.rand(F * C * k * k, device="cuda") - 0.5) * 0.5
b = torch.zeros(F, device="cuda")
y1, y2 = torch.empty(B * H * H * F, device="cuda"), torch.empty(B * H * H * F, device="cuda")
desc = cu.conv_desc(B, C, H, H, F, k, k, 1, 1, 1, 1)
cu.setWeightsStatic(w.data_ptr(), True)
def direct(): cu.conv2dForwardNhwc(desc, x.data_ptr(), 0, 0, 0, w.data_ptr(), None, b.data_ptr(), 9, y1.data_ptr(), 0, F, 1)
def conv(): cu.nchwToNhwc(B, C, H * H, x.data_ptr(), x4.data_ptr(), 4)
def tma():
    cu.conv2dForwardNhwc(desc, x4.data_ptr(), 0, 4, 1, w.data_ptr(), None, b.data_ptr(), 9, y2.data_ptr(), 0, F, 1)
def both(): conv(); tma()
for name, fn in (("direct fp32 (NHWC out)", direct), ("nchw->nhwc4", conv), ("tma im2col conv", tma), ("convert + tma conv", both)):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): fn()
    e1.record(stream); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1)/5:.3f} ms  [{cu.lastConvKernel()}]")
a, c = y1.cpu().numpy(), y2.cpu().numpy()
print("max rel diff tensor-core vs direct fp32:", float(np.abs(a - c).max() / np.abs(a).max()), "| hbm floor", round(4.0*(B*C*H*H+B*F*H*H)/6.45e12*1e3,3), "ms")
