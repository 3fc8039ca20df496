"""3xTF32 / TF32 accuracy characterisation on a B200 (run via gpurun): error vs K against fp64 truth."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from oracle import oracle as orc
from util import Dev, rel_max

cu = tensor.initCUDART(0)
rng = np.random.default_rng(0)
M = N = 256
print("mode K  rel_max_vs_fp64  mean_signed_err*sign(truth)/rms  cpu_fp32_rel_max")
for K in (32, 64, 256, 1024, 2048, 4608, 8192):
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    truth = A.astype(np.float64) @ B.astype(np.float64).T
    cpu = np.zeros((M, N), np.float32)
    orc.sgemm(0, 1, M, N, K, 1.0, A, K, B, K, 0.0, cpu, N)
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=M * N)
    for mode, name in ((0, "fp32"), (1, "3xtf32"), (2, "tf32")):
        cu.set_math_mode(mode)
        cu.gemm(False, True, M, N, K, 1.0, dA.ptr, 0, K, dB.ptr, 0, K, 0.0, dC.ptr, 0, N)
        got = dC.read((M, N)).astype(np.float64)
        err = got - truth
        bias = float(np.mean(err * np.sign(truth)) / np.sqrt(np.mean(truth ** 2)))
        print(f"{name:7s} {K:5d}  {rel_max(got, truth):.3e}  {bias:+.3e}  {rel_max(cpu, truth):.3e}")
cu.set_math_mode(1)

# ---- the channels-innermost convolution kernel (a 1x1 convolution over 256 pixels is the same GEMM), all tensor modes
print("\nnhwc conv 1x1 (M = 256 pixels, F = 256):  mode K  rel_max_vs_fp64  signed_bias/rms")
for K in (32, 256, 1024, 2048, 4608, 8192):
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)       # [pixel][channel]  (NHWC rows)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)       # [filter][channel]
    truth = A.astype(np.float64) @ B.astype(np.float64).T   # [pixel][filter]
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=M * N)
    desc = cu.conv_desc(1, K, 16, 16, N, 1, 1, 0, 0, 1, 1)
    for mode, name in ((1, "3xtf32"), (2, "tf32"), (3, "bf16x3")):
        cu.set_math_mode(mode)
        cu.conv2dForwardNhwc(desc, dA.ptr, 0, K, 1, dB.ptr, None, None, 4, dC.ptr, 0, N, 1)
        got = dC.read((M, N)).astype(np.float64)
        err = got - truth
        bias = float(np.mean(err * np.sign(truth)) / np.sqrt(np.mean(truth ** 2)))
        print(f"{name:7s} {K:5d}  {rel_max(got, truth):.3e}  {bias:+.3e}")
cu.set_math_mode(1)
