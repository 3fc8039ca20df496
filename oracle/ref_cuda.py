"""Witness: the reference's OWN CUDA kernels, run here — TEST INFRASTRUCTURE ONLY (never imported by the product).

`oracle/_ref/cuda_sgemm.cubin` is built from the reference's source where it lies (/root/reference/source/cuda_sgemm.cu,
never copied into this repository) by `make -C oracle ref` with the flags the reference passes to NVRTC
(source/nncuda.pas:1600-1605: -arch=sm_<cc>, --use_fast_math, -DBLOCK=512, -Dnfloat=float).  This module loads the
cubin through the CUDA driver API and launches the kernels with the argument lists and launch shapes of the
reference's host wrappers (TNNCuda<T> methods, source/nncuda.pas; cited per function), i.e. it replays what
`cuda.<op>(...)` does in the reference.  tests/test_gpu_ref_witness.py uses it to pin the C oracle
(oracle/tensorium_oracle.c) against outputs of the reference itself wherever the reference's CPU and GPU paths
compute the same thing (im2col / col2im index mapping, the k-sequential fp32 dot of sgemm1_*, bias / scale,
piecewise-linear activations, max-pool arg-max), and to document the three known CPU/GPU divergences
(variance / N vs / (N-1), clamp order, tanh) with the exact factor between them.

Device pointers are plain integers (the same cudaMalloc'ed buffers the product's C ABI hands out: both use the
device's primary context).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(HERE, "_ref", "cuda_sgemm.cubin")

NUM_THREADS = 0x200   # nncuda.pas:37
COPY_DIMX = COPY_DIMY = 8  # nncuda.pas:38-39


def available() -> bool:
    return os.path.exists(CUBIN)


def _ceil(x, y):  # cl_las.pas:971-979 (Ceil = CeilDiv*y)
    return (1 + (x - 1) // y) * y


def euclid_gcd(a, b):
    """nncuda.pas:1189-1210 -> (p, q, r)"""
    p, q, p_1, q_1 = 0, 1, 1, 0
    while True:
        c = a % b
        if c == 0:
            break
        p_2, q_2 = p_1, q_1
        p_1, q_1 = p, q
        p = p_2 - p_1 * (a // b)
        q = q_2 - q_1 * (a // b)
        a, b = b, c
    return p, q, b


class RefCuda:
    def __init__(self, device: int = 0):
        if not available():
            raise RuntimeError(f"{CUBIN} is missing: run `make -C oracle ref` where /root/reference exists")
        self.cu = C.CDLL("libcuda.so.1")
        self._ck(self.cu.cuInit(0))
        dev = C.c_int()
        self._ck(self.cu.cuDeviceGet(C.byref(dev), device))
        ctx = C.c_void_p()
        self._ck(self.cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev))
        self._ck(self.cu.cuCtxSetCurrent(ctx))
        self.mod = C.c_void_p()
        self._ck(self.cu.cuModuleLoad(C.byref(self.mod), CUBIN.encode()))
        self._fn = {}

    def _ck(self, r):
        if r != 0:
            s = C.c_char_p()
            self.cu.cuGetErrorString(r, C.byref(s))
            raise RuntimeError(f"CUDA driver error {r}: {(s.value or b'').decode()}")

    def kernel(self, name):
        if name not in self._fn:
            f = C.c_void_p()
            self._ck(self.cu.cuModuleGetFunction(C.byref(f), self.mod, name.encode()))
            self._fn[name] = f
        return self._fn[name]

    def launch(self, name, grid, block, args):
        """args: ctypes scalars in the kernel's parameter order (c_long for `long`, c_float, c_void_p for pointers)."""
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        arr = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p) for a in args])
        self._ck(self.cu.cuLaunchKernel(self.kernel(name), *[C.c_uint(int(g)) for g in grid],
                                        *[C.c_uint(int(b)) for b in block], C.c_uint(0), C.c_void_p(0), arr, C.c_void_p(0)))
        self._ck(self.cu.cuCtxSynchronize())

    # ---- wrappers, one per TNNCuda<T> method used by the conv path --------------------------------------------
    @staticmethod
    def _l(v):
        return C.c_long(int(v))

    @staticmethod
    def _p(v):
        return C.c_void_p(int(v))

    def gemm(self, transA, transB, M, N, K, alpha, A, aoff, lda, B, boff, ldb, beta, Cm, coff, ldc):
        """TNNCuda.gemm, useBLAS = 0 branch (nncuda.pas:691-716): kernel choice, 8x16 thread blocks."""
        l, p = self._l, self._p
        if not transA and not transB:
            name, dim = ("sgemm2_nn", (N, M)) if N > M else ("sgemm1_nn", (M, N))
        elif not transA and transB:
            name, dim = "sgemm1_nt", (M, N)
        elif transA and not transB:
            name, dim = "sgemm1_tn", (M, N)
        else:
            raise ValueError("the reference has no TT kernel")
        grid = ((dim[0] + 7) // 8, (dim[1] + 15) // 16)
        self.launch(name, grid, (8, 16), [l(M), l(N), l(K), C.c_float(alpha), p(A), l(aoff), l(lda), p(B), l(boff), l(ldb),
                                          C.c_float(beta), p(Cm), l(coff), l(ldc)])
        return name

    def forward_bias(self, dstSize, dst, offset, srcSize, src, incb, batch):
        """nncuda.pas:582-600"""
        l, p = self._l, self._p
        blockSize = dstSize // (srcSize * batch)
        self.launch("forward_bias", ((dstSize + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,),
                    [l(dstSize), l(srcSize), l(blockSize), p(dst), l(offset), p(src), l(1), l(incb)])

    def backward_bias(self, dstSize, dst, srcSize, src, srcOffset, incb, batch):
        """nncuda.pas:602-622 (grid = dstSize blocks)"""
        l, p = self._l, self._p
        blockSize = srcSize // (dstSize * batch)
        self.launch("backward_bias", (dstSize,), (NUM_THREADS,), [l(dstSize), p(dst), l(blockSize), p(src), l(srcOffset), l(batch)])

    def activate_array(self, N, x, offset, act):
        """nncuda.pas:524-542 (skipped on the host when linear)"""
        if act == 4:
            return
        self.launch("activate_array", ((N + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,),
                    [self._l(N), self._p(x), self._l(offset), C.c_int(act)])

    def derive_array(self, N, x, offset, act, delta):
        """nncuda.pas:562-580"""
        self.launch("gradient_array", ((N + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,),
                    [self._l(N), self._p(x), self._l(offset), C.c_int(act), self._p(delta)])

    @staticmethod
    def _col_dims(h, w, kh, kw, ph, pw, sy, sx, dy, dx):
        size_h, padding_h = h + 2 * ph, dy * (kh - 1) + 1
        col_h = (size_h - padding_h) // sy + 1 if size_h >= padding_h else 1
        size_w, padding_w = w + 2 * pw, dx * (kw - 1) + 1
        col_w = (size_w - padding_w) // sx + 1 if size_w >= padding_w else 1
        return col_h, col_w

    def im2col(self, c, h, w, kh, kw, ph, pw, sy, sx, dy, dx, im, imOffset, col, colOffset):
        """nncuda.pas:1144-1185: Xim2colKernelNormal, block (4,64)"""
        l, p = self._l, self._p
        col_h, col_w = self._col_dims(h, w, kh, kw, ph, pw, sy, sx, dy, dx)
        ceiled_w, ceiled_h = _ceil(col_w, COPY_DIMX), _ceil(col_h, COPY_DIMY)
        N = ceiled_h * c
        grid = ((ceiled_w + 3) // 4, (N + 63) // 64)
        self.launch("Xim2colKernelNormal", grid, (4, 64),
                    [l(ceiled_w), l(N), l(h), l(w), l(c), l(col_h), l(col_w), l(kh), l(kw), l(ph), l(pw), l(sy), l(sx),
                     l(dy), l(dx), p(im), l(imOffset), p(col), l(colOffset)])

    def col2im(self, c, h, w, kh, kw, ph, pw, sy, sx, dy, dx, col, colOffset, im, imOffset):
        """nncuda.pas:1212-1265: Xcol2imKernelNormal with the Bezout / GCD index algebra, accumulates into im"""
        l, p = self._l, self._p
        col_h, col_w = self._col_dims(h, w, kh, kw, ph, pw, sy, sx, dy, dx)
        sbh, dbh, gh = euclid_gcd(sy, dy)
        sbw, dbw, gw = euclid_gcd(sx, dx)
        w_ceiled = _ceil((w - 1) // gw + 1, COPY_DIMX)
        h_ceiled = _ceil((h - 1) // gh + 1, COPY_DIMY)
        N = h_ceiled * c
        grid = ((w_ceiled + 3) // 4, (N + 63) // 64)
        self.launch("Xcol2imKernelNormal", grid, (4, 64),
                    [l(w_ceiled), l(N), l(h), l(w), l(c), l(col_h), l(col_w), l(kh), l(kw), l(ph), l(pw), l(sy), l(sx),
                     l(dy), l(dx), l(sbh), l(sbw), l(dbh), l(dbw), l(gh), l(gw), p(col), l(colOffset), p(im), l(imOffset)])

    def forward_maxpool(self, aBatch, outC, outH, outW, inp, c, h, w, stride_x, stride_y, padding, kernelSize, indexes, out):
        """nncuda.pas:1088-1119: block (8,8,8) over (batch*channels, outH, outW)"""
        l, p = self._l, self._p
        N = aBatch * outC
        grid = ((N + 7) // 8, (outH + 7) // 8, (outW + 7) // 8)
        self.launch("forward_maxpool", grid, (8, 8, 8), [l(N), l(outH), l(outW), p(inp), l(c), l(h), l(w), l(stride_x),
                                                         l(stride_y), l(padding), l(kernelSize), p(indexes), p(out)])

    def means(self, srcSize, dstSize, groups, src, offset, means):
        """nncuda.pas:1330-1348: one block per channel"""
        l, p = self._l, self._p
        blockSize = srcSize // (dstSize * groups)
        self.launch("means", (dstSize,), (NUM_THREADS,), [l(dstSize), l(blockSize), l(groups), p(src), l(offset), p(means)])

    def variances(self, srcSize, dstSize, groups, src, offset, means, vars_):
        """nncuda.pas:1350-1369 (kernel divides by N, not N-1: cuda_sgemm.cu:1743)"""
        l, p = self._l, self._p
        blockSize = srcSize // (dstSize * groups)
        self.launch("vars", (dstSize,), (NUM_THREADS,), [l(dstSize), l(blockSize), l(groups), p(src), l(offset), p(means), p(vars_)])

    def normalize(self, srcSize, dstSize, groups, means, meansStride, vars_, varsStride, dst, dstOffset):
        """nncuda.pas:1371-1396: normblkvv, block (16,8,8) over (channels, blockSize, groups)"""
        l, p = self._l, self._p
        blockSize = dstSize // (srcSize * groups)
        M, N, K = srcSize, blockSize, groups
        grid = ((M + 15) // 16, (N + 7) // 8, (K + 7) // 8)
        self.launch("normblkvv", grid, (16, 8, 8), [l(M), l(N), l(K), p(means), l(meansStride), p(vars_), l(varsStride), p(dst), l(dstOffset)])

    def forward_scale(self, dstSize, dst, offset, scaleSize, scale, incb, batch):
        """nncuda.pas:1468-1488"""
        l, p = self._l, self._p
        blockSize = dstSize // (scaleSize * batch)
        self.launch("forward_scale", ((dstSize + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,),
                    [l(dstSize), l(scaleSize), l(blockSize), p(dst), l(offset), p(scale), l(1), l(incb)])

    def add_dots(self, N, dstSize, groups, src1, src2, srcOffset, dst):
        """nncuda.pas:1448-1466: one block per channel"""
        l, p = self._l, self._p
        blockSize = N // (dstSize * groups)
        self.launch("add_dots", (dstSize,), (NUM_THREADS,), [l(dstSize), l(groups), l(blockSize), p(src1), p(src2), l(srcOffset), p(dst)])

    def axpy(self, N, a, x, xOffset, incx, y, yOffset, incy):
        """nncuda.pas:945-964"""
        l, p = self._l, self._p
        self.launch("axpy", ((N + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,),
                    [l(N), C.c_float(a), p(x), l(xOffset), l(incx), p(y), l(yOffset), l(incy)])

    def scale(self, N, a, x, incx):
        """nncuda.pas:986-1003"""
        self.launch("scale", ((N + NUM_THREADS - 1) // NUM_THREADS,), (NUM_THREADS,), [self._l(N), C.c_float(a), self._p(x), self._l(incx)])
