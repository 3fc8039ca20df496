"""ctypes front-end of the CPU oracle (oracle/tensorium_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of tensorium_oracle.c.  Nothing under
tensorium_b200/ may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.  PARITY UNPINNED (the reference
has no golden vectors and cannot be built here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtensorium_oracle.so")

MODE_PAS = 0  # default Makefile builds: saxpy_pas / sdot_pas (ntensors.pas:1130-1145)
MODE_FMA = 1  # -dUSE_AVX2 builds: vfmadd231ps + 8-lane partial sums (ntensors.pas:1233-1435)

# TActivationType ordinals (ntypes.pas:66-71)
ACT = dict(
    LOGISTIC=0, RELU=1, RELU6=2, RELIE=3, LINEAR=4, RAMP=5, TANH=6, PLSE=7, REVLEAKY=8, LEAKY=9,
    ELU=10, LOGGY=11, STAIR=12, HARDTAN=13, LHTAN=14, SELU=15, SOFTMAX=16, GELU=17, SWISH=18,
)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tensorium_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
i64 = C.c_int64
f32 = C.c_float
fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_dot.restype = f32
        _lib.orc_sum.restype = f32
        _lib.orc_out_dim.restype = i64
        _lib.orc_get_mode.restype = C.c_int
        _lib.orc_get_max_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(fp)


def set_mode(mode: int):
    lib().orc_set_mode(C.c_int(mode))


def set_threads(n: int):
    lib().orc_set_threads(C.c_int(n))


def max_threads() -> int:
    return lib().orc_get_max_threads()


def out_dim(n, pad, k, stride, dil=1) -> int:
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


# ---------------------------------------------------------------- BLAS

def sgemm(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, aoff=0, boff=0, coff=0):
    """cblas_sgemm (ntensors.pas:2231-2286), row-major, in place on Cm (flat float32 array)."""
    a = A.reshape(-1)[aoff:]
    b = B.reshape(-1)[boff:]
    c = Cm.reshape(-1)[coff:]
    lib().orc_sgemm(C.c_int(int(transA)), C.c_int(int(transB)), i64(M), i64(N), i64(K), f32(alpha),
                    _p(a), i64(lda), _p(b), i64(ldb), f32(beta), _p(c), i64(ldc))
    return Cm


def sgemm_strided_batched(transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, Cm, ldc, sC, batch):
    lib().orc_sgemm_strided_batched(C.c_int(int(transA)), C.c_int(int(transB)), i64(M), i64(N), i64(K),
                                    f32(alpha), _p(A.reshape(-1)), i64(lda), i64(sA), _p(B.reshape(-1)),
                                    i64(ldb), i64(sB), f32(beta), _p(Cm.reshape(-1)), i64(ldc), i64(sC),
                                    i64(batch))
    return Cm


def dgemm_truth(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta=0.0, Cin=None, ldc=None):
    ldc = ldc or N
    out = np.zeros((M, ldc), dtype=np.float64)
    cin = _p(Cin.reshape(-1)) if Cin is not None else None
    lib().orc_dgemm_truth(C.c_int(int(transA)), C.c_int(int(transB)), i64(M), i64(N), i64(K), f32(alpha),
                          _p(A.reshape(-1)), i64(lda), _p(B.reshape(-1)), i64(ldb), f32(beta), cin,
                          out.ctypes.data_as(dp), i64(ldc))
    return out[:, :N]


def axpy(a, x, y):
    lib().orc_axpy(i64(x.size), f32(a), _p(x.reshape(-1)), i64(1), _p(y.reshape(-1)), i64(1))
    return y


def scal(alpha, x):
    lib().orc_scal(i64(x.size), f32(alpha), _p(x.reshape(-1)), i64(1))
    return x


# ---------------------------------------------------------------- im2col / col2im

def im2col(x, kh, kw, padH, padW, sy, sx, dy=1, dx=1):
    """sim2Col on one image x[C,H,W] -> col[C*kh*kw, OH*OW] (ntensors.pas:11415-11491)."""
    c, h, w = x.shape
    oh, ow = out_dim(h, padH, kh, sy, dy), out_dim(w, padW, kw, sx, dx)
    col = np.empty((c * kh * kw, oh * ow), dtype=np.float32)
    lib().orc_im2col(i64(c), i64(h), i64(w), i64(kh), i64(kw), i64(padH), i64(padW), i64(sy), i64(sx),
                     i64(dy), i64(dx), _p(x), i64(0), _p(col), i64(0))
    return col


def col2im(col, c, h, w, kh, kw, padH, padW, sy, sx, dy=1, dx=1, im=None):
    """scol2im: scatter-ADD col into im[C,H,W] (ntensors.pas:11650-11763)."""
    if im is None:
        im = np.zeros((c, h, w), dtype=np.float32)
    lib().orc_col2im(i64(c), i64(h), i64(w), i64(kh), i64(kw), i64(padH), i64(padW), i64(sy), i64(sx),
                     i64(dy), i64(dx), _p(col.reshape(-1)), i64(0), _p(im.reshape(-1)), i64(0))
    return im


def im2col_batched(x, kh, kw, padH, padW, sy, sx, dy=1, dx=1):
    b, c, h, w = x.shape
    oh, ow = out_dim(h, padH, kh, sy, dy), out_dim(w, padW, kw, sx, dx)
    col = np.empty((b, c * kh * kw, oh * ow), dtype=np.float32)
    lib().orc_im2col_strided_batched(i64(c), i64(h), i64(w), i64(kh), i64(kw), i64(padH), i64(padW), i64(sy),
                                     i64(sx), i64(dy), i64(dx), _p(x), i64(c * h * w), i64(0), _p(col),
                                     i64(col[0].size), i64(0), i64(b))
    return col


def col2im_batched(col, b, c, h, w, kh, kw, padH, padW, sy, sx, dy=1, dx=1, im=None):
    if im is None:
        im = np.zeros((b, c, h, w), dtype=np.float32)
    lib().orc_col2im_strided_batched(i64(c), i64(h), i64(w), i64(kh), i64(kw), i64(padH), i64(padW), i64(sy),
                                     i64(sx), i64(dy), i64(dx), _p(col.reshape(-1)), i64(col.size // b), i64(0),
                                     _p(im.reshape(-1)), i64(c * h * w), i64(0), i64(b))
    return im


def conv2d(x, weights, pad, sy, sx, dil=1, kh=None, kw=None):
    """TTensor.Conv2D (ntensors.pas:8252-8349).  x[B,C,H,W], weights[F, C*kh*kw] -> out[B,F,OH,OW]."""
    b, c, h, w = x.shape
    f = weights.shape[0]
    if kh is None:
        k2 = weights.size // (f * c)
        kh = kw = int(round(k2 ** 0.5))
    oh, ow = out_dim(h, pad, kh, sy, dil), out_dim(w, pad, kw, sx, dil)
    out = np.zeros((b, f, oh, ow), dtype=np.float32)
    need_ws = kh * kw != 1 or dil != 1 or sx * sy != 1
    ws = np.empty(b * c * kh * kw * oh * ow if need_ws else 1, dtype=np.float32)
    lib().orc_conv2d(i64(b), i64(c), i64(h), i64(w), i64(f), i64(kh), i64(kw), i64(pad), i64(pad), i64(sy),
                     i64(sx), i64(dil), i64(dil), _p(x), _p(weights.reshape(-1)), _p(out), _p(ws))
    return out


# ---------------------------------------------------------------- bias / scale / BN

def forward_bias(dst, bias, groups):
    lib().orc_forward_bias(_p(dst.reshape(-1)), i64(dst.size), i64(groups), _p(bias), i64(bias.size))
    return dst


def forward_scale(dst, scale, groups):
    lib().orc_forward_scale(_p(dst.reshape(-1)), i64(dst.size), i64(groups), _p(scale), i64(scale.size))
    return dst


def add_sums(dst, src, groups):
    lib().orc_add_sums(_p(dst), i64(dst.size), _p(src.reshape(-1)), i64(src.size), i64(groups))
    return dst


def add_dots(dst, src1, src2, groups):
    lib().orc_add_dots(_p(dst), i64(dst.size), _p(src1.reshape(-1)), _p(src2.reshape(-1)), i64(src1.size),
                       i64(groups))
    return dst


def means_and_vars(x, groups, n):
    m = np.empty(n, dtype=np.float32)
    v = np.empty(n, dtype=np.float32)
    lib().orc_means_and_vars(_p(x.reshape(-1)), i64(x.size), i64(groups), _p(m), _p(v), i64(n))
    return m, v


def block_normalize(x, mean, var, groups):
    lib().orc_block_normalize(_p(x.reshape(-1)), i64(x.size), i64(groups), _p(mean), _p(var), i64(mean.size))
    return x


def mean_and_variance_delta(delta, x, mean, var, groups):
    n = mean.size
    md = np.empty(n, dtype=np.float32)
    vd = np.empty(n, dtype=np.float32)
    lib().orc_mean_and_variance_delta(_p(delta.reshape(-1)), _p(x.reshape(-1)), i64(delta.size), i64(groups),
                                      _p(mean), _p(var), _p(md), _p(vd), i64(n))
    return md, vd


def normalize_delta(delta, x, mean, var, mean_delta, var_delta, groups):
    lib().orc_normalize_delta(_p(delta.reshape(-1)), _p(x.reshape(-1)), i64(delta.size), i64(groups), _p(mean),
                              _p(var), _p(mean_delta), _p(var_delta), i64(mean.size))
    return delta


def activate_array(x, act):
    lib().orc_activate_array(_p(x.reshape(-1)), i64(x.size), C.c_int(act))
    return x


def gradient_array(out, act, delta):
    lib().orc_gradient_array(_p(out.reshape(-1)), i64(out.size), C.c_int(act), _p(delta.reshape(-1)))
    return delta


def fuse_batchnorm(weights, biases, scales, rmean, rvar):
    f = biases.size
    lib().orc_fuse_batchnorm(_p(weights.reshape(-1)), _p(biases), _p(scales), _p(rmean), _p(rvar), i64(f),
                             i64(weights.size // f))
    return weights, biases


# ---------------------------------------------------------------- CIFAR-10 training path neighbours

def clamp(x, lo, hi):
    lib().orc_clamp(i64(x.size), f32(lo), f32(hi), _p(x.reshape(-1)))
    return x


def maxpool_forward(x, ksize, stride_x, stride_y, padding=0, want_indexes=True):
    b, c, h, w = x.shape
    outH = (h + padding - ksize) // stride_y + 1  # nMaxPoolLayer.pas:106-107
    outW = (w + padding - ksize) // stride_x + 1
    out = np.empty((b, c, outH, outW), dtype=np.float32)
    idx = np.empty(out.shape, dtype=np.int64) if want_indexes else None
    lib().orc_maxpool_forward(i64(b), i64(c), i64(h), i64(w), i64(outH), i64(outW), i64(ksize), i64(stride_x),
                              i64(stride_y), i64(padding), _p(x.reshape(-1)), _p(out.reshape(-1)),
                              idx.ctypes.data_as(C.c_void_p) if want_indexes else None)
    return out, idx


def maxpool_backward(indexes, delta, upstream):
    lib().orc_maxpool_backward(i64(delta.size), indexes.ctypes.data_as(C.c_void_p), _p(delta.reshape(-1)),
                               _p(upstream.reshape(-1)))
    return upstream


def softmax_batch(x, groups=1, temp=1.0):
    b, n = x.shape
    out = np.empty_like(x)
    lib().orc_softmax_batch(_p(x.reshape(-1)), i64(n // groups), i64(b), i64(n), i64(groups), i64(n // groups),
                            i64(1), f32(temp), _p(out.reshape(-1)))
    return out


def softmax_cross_entropy(pred, truth):
    delta = np.empty_like(pred)
    err = np.empty_like(pred)
    lib().orc_softmax_cross_entropy(i64(pred.size), _p(pred.reshape(-1)), _p(truth.reshape(-1)),
                                    _p(delta.reshape(-1)), _p(err.reshape(-1)))
    return delta, err


def dropout_rand(n, seed):
    r = np.empty(n, dtype=np.float32)
    lib().orc_dropout_rand(i64(n), C.c_uint64(seed), _p(r))
    return r


def dropout_apply(x, rnd, probability, scale):
    lib().orc_dropout_apply(i64(x.size), _p(rnd.reshape(-1)), f32(probability), f32(scale), _p(x.reshape(-1)))
    return x


def vsum(x):
    return float(lib().orc_sum(i64(x.size), _p(x.reshape(-1)), i64(1)))
