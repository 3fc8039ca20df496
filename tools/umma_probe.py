"""Diagnostic for the tcgen05 operand layouts (run on a B200 via gpurun).

Feeds the UMMA GEMM with index-encoded operands against one-hot partners so that, if a shared-memory layout or
descriptor field were wrong, the output shows WHERE each element landed.  Writes gpurun_out/umma_probe.npz and
prints a per-case verdict.  Not part of the product or the test-suite.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tensorium_b200 import tensor  # noqa: E402
from util import Dev  # noqa: E402


def main():
    cu = tensor.initCUDART(0)
    print("device:", cu.name, cu.device_info())
    out = {}
    M, N, K = 128, 64, 64
    ok_all = True
    for mode, mname in ((2, "tf32"), (1, "3xtf32")):
        cu.set_math_mode(mode)
        for ta in (0, 1):
            for tb in (0, 1):
                # logical A(m,k) = m*64 + k (exact only in 3xTF32; tf32 keeps 11 bits -> use m + k/64 style check)
                Al = (np.arange(M)[:, None] * 64 + np.arange(K)[None, :]).astype(np.float32)
                if mode == 2:
                    Al = (np.arange(M)[:, None] + np.arange(K)[None, :] * 128).astype(np.float32) % 2048
                Bl = np.eye(K, N, dtype=np.float32)  # B(k,n) = delta(k==n)  -> D[m][n] = A(m,n)
                A = np.ascontiguousarray(Al.T) if ta else Al
                B = np.ascontiguousarray(Bl.T) if tb else Bl
                dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=M * N)
                cu.gemm(ta, tb, M, N, K, 1.0, dA.ptr, 0, A.shape[1], dB.ptr, 0, B.shape[1], 0.0, dC.ptr, 0, N)
                D = dC.read((M, N))
                want = Al[:, :N]
                good = np.array_equal(D, want)
                ok_all &= good
                print(f"{mname} ta={ta} tb={tb} A-encoded: {'OK' if good else 'MISMATCH'}")
                if not good:
                    bad = np.argwhere(D != want)
                    print("   first mismatches (m,n,got,want):", [(int(i), int(j), float(D[i, j]), float(want[i, j])) for i, j in bad[:8]])
                out[f"{mname}_{ta}{tb}_Aenc"] = D
                # B encoded, A one-hot: A(m,k) = delta(k == m % K) -> D[m][n] = B(m%K, n)
                Al2 = np.zeros((M, K), np.float32)
                Al2[np.arange(M), np.arange(M) % K] = 1
                Bl2 = (np.arange(K)[:, None] * 32 + np.arange(N)[None, :] % 32).astype(np.float32)
                A = np.ascontiguousarray(Al2.T) if ta else Al2
                B = np.ascontiguousarray(Bl2.T) if tb else Bl2
                dA, dB = Dev(cu, A), Dev(cu, B)
                cu.gemm(ta, tb, M, N, K, 1.0, dA.ptr, 0, A.shape[1], dB.ptr, 0, B.shape[1], 0.0, dC.ptr, 0, N)
                D = dC.read((M, N))
                want = Bl2[np.arange(M) % K]
                good = np.array_equal(D, want)
                ok_all &= good
                print(f"{mname} ta={ta} tb={tb} B-encoded: {'OK' if good else 'MISMATCH'}")
                if not good:
                    bad = np.argwhere(D != want)
                    print("   first mismatches (m,n,got,want):", [(int(i), int(j), float(D[i, j]), float(want[i, j])) for i, j in bad[:8]])
                out[f"{mname}_{ta}{tb}_Benc"] = D
    cu.set_math_mode(1)
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/umma_probe.npz", **out)
    print("ALL OK" if ok_all else "SOME MISMATCH")


if __name__ == "__main__":
    main()
