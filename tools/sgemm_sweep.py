"""BASELINE config 1: standalone SGEMM sweep through tnn_gemm (M = N = K in 256..8192; NN / NT / TN) with a sampled
fp64 check; host-timed around a stream synchronise, so the n <= 512 rows are launch latency.
usage: sgemm_sweep.py [math [n]]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev
MATH = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16x3": 3}
math = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
cu = tensor.initCUDART(0)
cu.set_math_mode(MATH[math])
rng = np.random.default_rng(0)
print(f"math={math}")
SIZES = (int(sys.argv[2]),) if len(sys.argv) > 2 else (256, 512, 1024, 2048, 4096, 8192)
for n in SIZES:
    A = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    B = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=n * n)
    row = []
    for name, ta, tb in (("NN", 0, 0), ("NT", 0, 1), ("TN", 1, 0)):
        for _ in range(2):
            cu.gemm(ta, tb, n, n, n, 1.0, dA.ptr, 0, n, dB.ptr, 0, n, 0.0, dC.ptr, 0, n)
        cu.finish()
        reps = 3 if n >= 4096 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            cu.gemm(ta, tb, n, n, n, 1.0, dA.ptr, 0, n, dB.ptr, 0, n, 0.0, dC.ptr, 0, n)
        cu.finish()
        dt = (time.perf_counter() - t0) / reps
        C = dC.read((n, n))
        idx = rng.integers(0, n, (64, 2))
        opA = A.T if ta else A
        opB = B.T if tb else B
        ref = np.array([np.dot(opA[i].astype(np.float64), opB[:, j].astype(np.float64)) for i, j in idx])
        err = np.abs(C[idx[:, 0], idx[:, 1]] - ref).max() / np.abs(ref).max()
        row.append(f"{name} {dt*1e3:8.3f} ms {2.0*n**3/dt/1e12:7.1f} TF/s err {err:.1e}")
    print(f"n={n:5d}  " + "  |  ".join(row), flush=True)

if len(sys.argv) <= 2:
    # conv-shaped rectangles (M = F, N = OH*OW*b, K = C*k*k) in the three modes the layers use
    # (nConvolutionLayer.pas:1087 NN beta=0, :1228 NT beta=1, :1255 TN beta=0)
    print("rectangles: M=F N=OH*OW*b K=C*k*k")
    for F, C, k, hw, b in ((64, 32, 3, 208, 4), (128, 64, 3, 104, 8), (256, 128, 3, 52, 16), (512, 256, 3, 26, 64),
                           (1024, 512, 3, 13, 64), (512, 1024, 1, 13, 64)):
        M, N, K = F, hw * hw * b, C * k * k
        row = []
        for name, ta, tb, beta, (m, n, kk) in (("NN b=0", 0, 0, 0.0, (M, N, K)), ("NT b=1", 0, 1, 1.0, (M, K, N)),
                                               ("TN b=0", 1, 0, 0.0, (K, N, M))):
            # NN: W[M,K] x col[K,N];  NT (wgrad): delta[M,N] x col[K,N]^T -> [M,K];  TN (dgrad): W[M,K]^T x delta[M,N] -> [K,N]
            ar, ac = (kk, m) if ta else (m, kk)
            br, bc = (n, kk) if tb else (kk, n)
            A = rng.uniform(-1, 1, (ar, ac)).astype(np.float32)
            B = rng.uniform(-1, 1, (br, bc)).astype(np.float32)
            C0 = rng.uniform(-1, 1, (m, n)).astype(np.float32)
            dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, C0)
            call = lambda: cu.gemm(ta, tb, m, n, kk, 1.0, dA.ptr, 0, ac, dB.ptr, 0, bc, beta, dC.ptr, 0, n)
            call()
            C1 = dC.read((m, n))
            idx = np.stack([rng.integers(0, m, 64), rng.integers(0, n, 64)], 1)
            opA = A.T if ta else A
            opB = B.T if tb else B
            ref = np.array([np.dot(opA[i].astype(np.float64), opB[:, j].astype(np.float64)) + beta * C0[i, j] for i, j in idx])
            err = np.abs(C1[idx[:, 0], idx[:, 1]] - ref).max() / np.abs(ref).max()
            call(); cu.finish()
            t0 = time.perf_counter()
            for _ in range(5):
                call()
            cu.finish()
            dt = (time.perf_counter() - t0) / 5
            row.append(f"{name} {dt*1e3:7.3f} ms {2.0*m*n*kk/dt/1e12:6.1f} TF/s err {err:.0e}")
        print(f"F={F:4d} C={C:4d} k={k} {hw:3d}^2 x{b:2d}  " + " | ".join(row), flush=True)
