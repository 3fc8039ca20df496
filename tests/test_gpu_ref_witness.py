"""Pinning the oracle against THE REFERENCE ITSELF, run here: the reference's own CUDA kernels
(/root/reference/source/cuda_sgemm.cu compiled where it lies into oracle/_ref/cuda_sgemm.cubin by `make -C oracle ref`,
launched by oracle/ref_cuda.py with the argument lists and launch shapes of the TNNCuda<T> wrappers in
source/nncuda.pas) against the C oracle (oracle/tensorium_oracle.c, the restatement of the reference's CPU path), and
against the product's kernels on the same device buffers.

Where the reference's CPU and GPU paths compute the same thing the comparison is bit-exact (im2col / col2im index
mapping, the k-sequential FMA dot of sgemm1_nn / sgemm2_nn / sgemm1_nt / sgemm1_tn, bias, scale, leaky / relu,
max-pool values and arg-max, axpy / scale).  Where they are known to differ the test states the exact
relation: variance / N on the GPU vs / (N-1) on the CPU; sqrt(max(var, eps)) vs max(sqrt(var), eps); --use_fast_math
transcendental approximations.  The product follows the CPU side in those cases."""
import numpy as np
import pytest

from util import Dev, rel_elem, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(cuda_backend):
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.fail("oracle/_ref/cuda_sgemm.cubin is missing: build it with `make -C oracle ref` (needs /root/reference)")
    return ref_cuda.RefCuda(0)


GEMM_CASES = [  # transA, transB, M, N, K
    (False, False, 37, 19, 150),   # M > N: sgemm1_nn
    (False, False, 16, 200, 75),   # N > M: sgemm2_nn (the conv forward shape: filters x pixels)
    (False, True, 24, 27, 169),    # NT: wgrad shape
    (True, False, 27, 40, 24),     # TN: dgrad shape
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_reference_sgemm_kernels_match_the_oracle_bit_for_bit(cuda_backend, oracle, ref, case):
    """sgemm1_* = one thread per C element, `d += a*b` over k (cuda_sgemm.cu:156-275; an FFMA chain under
    --use_fast_math).  The CPU path accumulates every C element in the same k order with vfmadd231ps
    (s_nn -> saxpy_avx2 per k; s_nt -> sdot: 8-lane partial sums, hence NT only agrees to rounding)."""
    cu = cuda_backend
    tA, tB, M, N, K = case
    rng = np.random.default_rng(M * 1000 + N)
    A = rng.uniform(-1, 1, (K, M) if tA else (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K) if tB else (K, N)).astype(np.float32)
    lda, ldb = A.shape[1], B.shape[1]
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, np.zeros((M, N), np.float32))
    name = ref.gemm(tA, tB, M, N, K, 1.0, dA.ptr, 0, lda, dB.ptr, 0, ldb, 0.0, dC.ptr, 0, N)
    got = dC.read((M, N))
    oracle.set_mode(oracle.MODE_FMA)
    try:
        want = np.zeros((M, N), np.float32)
        oracle.sgemm(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, want, N)
    finally:
        oracle.set_mode(oracle.MODE_PAS)
    if tB and not tA:
        assert rel_elem(got, want, floor=1.0) < 2e-6, name   # sdot's 8-lane partial sums vs the GPU's single chain
    else:
        assert np.array_equal(got, want), name
    # and the product's fp32 mode (exact FMA chain, k ascending) carries the same bits
    dP = Dev(cu, np.zeros((M, N), np.float32))
    cu.set_math_mode(0)
    try:
        cu.gemm(tA, tB, M, N, K, 1.0, dA.ptr, 0, lda, dB.ptr, 0, ldb, 0.0, dP.ptr, 0, N)
    finally:
        cu.set_math_mode(1)
    assert np.array_equal(dP.read((M, N)), got), name


IM2COL_CASES = [  # c, h, w, k, pad, stride, dil
    (3, 9, 11, 3, 1, 1, 1), (4, 12, 12, 3, 1, 2, 1), (2, 10, 7, 5, 2, 1, 1), (5, 8, 8, 1, 0, 1, 1),
    (2, 13, 13, 3, 0, 2, 1), (3, 11, 9, 3, 2, 1, 2),
]


@pytest.mark.parametrize("case", IM2COL_CASES)
def test_reference_im2col_col2im_kernels_match_oracle_and_product(cuda_backend, oracle, ref, case):
    cu = cuda_backend
    c, h, w, k, p, s, d = case
    rng = np.random.default_rng(sum(case))
    im = rng.uniform(-1, 1, (c, h, w)).astype(np.float32)
    want = oracle.im2col(im, k, k, p, p, s, s, d, d)
    dim, dcol, dcol2 = Dev(cu, im), Dev(cu, n=want.size), Dev(cu, n=want.size)
    cu.fill(want.size, dcol.ptr, 0, 123.0, 1)
    ref.im2col(c, h, w, k, k, p, p, s, s, d, d, dim.ptr, 0, dcol.ptr, 0)
    got = dcol.read(want.shape)
    assert np.array_equal(got, want)                       # reference GPU kernel == CPU restatement, bit for bit
    cu.im2col(c, h, w, k, k, p, p, s, s, d, d, dim.ptr, 0, dcol2.ptr, 0)
    assert np.array_equal(dcol2.read(want.shape), got)     # product == reference GPU kernel
    # col2im accumulates; small integers keep every summation order exact
    col = rng.integers(-4, 5, want.shape).astype(np.float32)
    im0 = rng.integers(-4, 5, im.shape).astype(np.float32)
    want_im = oracle.col2im(col, c, h, w, k, k, p, p, s, s, d, d, im=im0.copy())
    dcol.write(col)
    dA, dB = Dev(cu, im0), Dev(cu, im0)
    ref.col2im(c, h, w, k, k, p, p, s, s, d, d, dcol.ptr, 0, dA.ptr, 0)
    cu.col2im(c, h, w, k, k, p, p, s, s, d, d, dcol.ptr, 0, dB.ptr, 0)
    a, b = dA.read(im.shape), dB.read(im.shape)
    if d == 1:  # the CPU's (kernel_row - pad)*dil differs from im2col's -pad + kernel_row*dil for dil > 1
        assert np.array_equal(a, want_im)
    assert np.array_equal(b, want_im)


def test_reference_bias_scale_activation_kernels(cuda_backend, oracle, ref):
    cu = cuda_backend
    rng = np.random.default_rng(2)
    B, F, P = 3, 7, 45
    x = rng.uniform(-2, 2, (B, F, P)).astype(np.float32)
    bias = rng.uniform(-1, 1, F).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    dx, db, ds = Dev(cu, x), Dev(cu, bias), Dev(cu, sc)
    ref.forward_scale(x.size, dx.ptr, 0, F, ds.ptr, 1, B)
    ref.forward_bias(x.size, dx.ptr, 0, F, db.ptr, 1, B)
    want = x.copy()
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, bias, B)
    got = dx.read(x.shape)
    assert np.array_equal(got, want)
    for act, exact in (("RELU", True), ("LOGISTIC", False), ("TANH", False)):
        dy = Dev(cu, got)
        ref.activate_array(got.size, dy.ptr, 0, oracle.ACT[act])
        w2 = got.copy()
        oracle.activate_array(w2, oracle.ACT[act])
        g2 = dy.read(got.shape)
        if exact:
            assert np.array_equal(g2, w2), act
        else:  # --use_fast_math __expf; tanh additionally uses 2/(1+e^-2x)-1 on the GPU (cuda_sgemm.cu:369)
            assert rel_elem(g2, w2) < 5e-6, act
    # leaky: the kernel multiplies by 0.1f in fp32, the Pascal CPU code by the double constant: one rounding apart
    dy = Dev(cu, got)
    ref.activate_array(got.size, dy.ptr, 0, oracle.ACT["LEAKY"])
    w2 = got.copy()
    oracle.activate_array(w2, oracle.ACT["LEAKY"])
    assert rel_elem(dy.read(got.shape), w2) < 1.2e-7
    # gradient (on outputs): leaky / relu are selections
    delta = rng.uniform(-1, 1, got.shape).astype(np.float32)
    for act in ("RELU", "LEAKY"):
        dd = Dev(cu, delta)
        dout = Dev(cu, got)
        ref.derive_array(got.size, dout.ptr, 0, oracle.ACT[act], dd.ptr)
        w3 = delta.copy()
        oracle.gradient_array(got, oracle.ACT[act], w3)
        assert rel_elem(dd.read(delta.shape), w3) < 1.2e-7, act


def test_reference_maxpool_kernel(cuda_backend, oracle, ref):
    cu = cuda_backend
    rng = np.random.default_rng(4)
    B, C, H, W, k, s = 2, 3, 8, 10, 2, 2
    x = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    want, widx = oracle.maxpool_forward(x, k, s, s, 0)
    OH, OW = want.shape[2], want.shape[3]
    dx, dout = Dev(cu, x), Dev(cu, n=want.size)
    didx = cu.createDeviceBuffer(want.size * 8)
    ref.forward_maxpool(B, C, OH, OW, dx.ptr, C, H, W, s, s, 0, k, didx, dout.ptr)
    assert np.array_equal(dout.read(want.shape), want)
    idx = np.empty(want.size, np.int64)
    cu.readBuffer(didx, idx.nbytes, idx)
    assert np.array_equal(idx.reshape(want.shape), np.asarray(widx).reshape(want.shape))
    cu.freeDeviceBuffer(didx)


def test_reference_batchnorm_statistics_and_the_documented_divergences(cuda_backend, oracle, ref):
    """means: same value to summation-order rounding.  vars: the GPU kernel divides by N (cuda_sgemm.cu:1743), the CPU
    by N-1 (ntensors.pas:9165-9166) -> exactly the factor (N-1)/N.  normblkvv: sqrt(max(var, eps)) vs the CPU's
    max(sqrt(var), eps) -> equal away from the clamp, fast-math division apart."""
    cu = cuda_backend
    rng = np.random.default_rng(6)
    B, F, P = 4, 6, 50
    x = rng.uniform(-1, 1, (B, F, P)).astype(np.float32)
    mean, var = oracle.means_and_vars(x, B, F)
    dx, dm, dv = Dev(cu, x), Dev(cu, n=F), Dev(cu, n=F)
    ref.means(x.size, F, B, dx.ptr, 0, dm.ptr)
    ref.variances(x.size, F, B, dx.ptr, 0, dm.ptr, dv.ptr)
    gm, gv = dm.read(), dv.read()
    N = B * P
    assert rel_max(gm, mean) < 1e-6
    assert rel_max(gv * (N / (N - 1.0)), var) < 2e-6          # the documented divergence, and nothing else
    # the product follows the CPU
    pm, pv = Dev(cu, n=F), Dev(cu, n=F)
    cu.meansAndVars(x.size, F, B, dx.ptr, 0, pm.ptr, pv.ptr)
    assert rel_max(pm.read(), mean) < 1e-6 and rel_max(pv.read(), var) < 2e-6
    # normalise with the CPU's statistics on both sides
    dm.write(mean), dv.write(var)
    ref.normalize(F, x.size, B, dm.ptr, 1, dv.ptr, 1, dx.ptr, 0)
    want = x.copy()
    oracle.block_normalize(want, mean, var, B)
    assert rel_elem(dx.read(x.shape), want) < 2e-6
    # reductions of the backward pass
    delta = rng.uniform(-1, 1, x.shape).astype(np.float32)
    dd, dacc, dacc2 = Dev(cu, delta), Dev(cu, np.zeros(F, np.float32)), Dev(cu, np.zeros(F, np.float32))
    ref.backward_bias(F, dacc.ptr, delta.size, dd.ptr, 0, 1, B)
    w = np.zeros(F, np.float32)
    oracle.add_sums(w, delta, B)
    assert rel_max(dacc.read(), w) < 2e-6
    dx.write(x)
    ref.add_dots(x.size, F, B, dx.ptr, dd.ptr, 0, dacc2.ptr)
    w = np.zeros(F, np.float32)
    oracle.add_dots(w, x, delta, B)
    assert rel_max(dacc2.read(), w) < 2e-6


def test_reference_axpy_scale_kernels(cuda_backend, oracle, ref):
    cu = cuda_backend
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, 1000).astype(np.float32)
    y = rng.uniform(-1, 1, 1000).astype(np.float32)
    dx, dy, py = Dev(cu, x), Dev(cu, y), Dev(cu, y)
    ref.axpy(x.size, 0.3, dx.ptr, 0, 1, dy.ptr, 0, 1)
    cu.axpy(x.size, 0.3, dx.ptr, 0, 1, py.ptr, 0, 1)
    a, b = dy.read(), py.read()
    # --use_fast_math contracts a*x + y into one FMA on the GPU; the CPU's saxpy_pas rounds the product first, its AVX2
    # build uses vfmadd231ps: the kernel must carry the bits of one of the two CPU flavours
    a32 = np.float32(0.3)
    fma = (a32.astype(np.float64) * x.astype(np.float64) + y.astype(np.float64)).astype(np.float32)  # single rounding
    two = (a32 * x).astype(np.float32) + y
    assert np.array_equal(a, fma) or np.array_equal(a, two)
    want = y.copy()
    oracle.axpy(0.3, x, want)
    assert rel_max(a, want) < 2e-7 and rel_max(b, want) < 2e-7   # (cancelling elements: absolute, not relative, error)
    ref.scale(x.size, 0.7, dx.ptr, 1)
    w = x.copy()
    oracle.scal(0.7, w)
    assert np.array_equal(dx.read(), w)
