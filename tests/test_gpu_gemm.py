"""GPU parity of SGEMM (tnn_gemm / gemmStridedBatched) against the oracle's cblas_sgemm restatement
(ntensors.pas:2231-2286) in the three math modes."""
import numpy as np
import pytest

from util import TOL_3XTF32, TOL_TF32, Dev, rel_elem, rel_max

pytestmark = pytest.mark.gpu

FP32, X3, TF32 = 0, 1, 2


def run_gemm(cu, oracle, mode, ta, tb, M, N, K, alpha, beta, seed=0, ldpad=0, off=0):
    rng = np.random.default_rng(seed)
    lda = (M if ta else K) + ldpad
    ldb = (K if tb else N) + ldpad
    ldc = N + ldpad
    A = rng.uniform(-1, 1, ((K if ta else M), lda)).astype(np.float32)
    B = rng.uniform(-1, 1, ((N if tb else K), ldb)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (M, ldc)).astype(np.float32)
    dA, dB, dC = Dev(cu, A, pad_front=off), Dev(cu, B, pad_front=off), Dev(cu, C0, pad_front=off)
    cu.set_math_mode(mode)
    cu.gemm(ta, tb, M, N, K, alpha, dA._base, off, lda, dB._base, off, ldb, beta, dC._base, off, ldc)
    got = dC.read(C0.shape)
    want = C0.copy()
    oracle.sgemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, want, ldc)
    cu.set_math_mode(X3)
    return got, want, C0


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (130, 70, 45), (256, 256, 256), (37, 29, 53), (300, 513, 1000)])
def test_gemm_fp32_mode(cuda_backend, oracle, ta, tb, M, N, K):
    got, want, C0 = run_gemm(cuda_backend, oracle, FP32, ta, tb, M, N, K, 0.75, 0.5, seed=M + N)
    assert rel_max(got, want) < 2e-6
    # padding columns of C untouched
    assert np.array_equal(got[:, N:], C0[:, N:])


def test_gemm_fp32_nn_bit_exact_with_fma_oracle(cuda_backend, oracle):
    # k-sequential FMA accumulation == s_nn with saxpy_avx2 (ntensors.pas:2093-2099,1308-1435) when alpha=1, beta=0
    oracle.set_mode(oracle.MODE_FMA)
    try:
        got, want, _ = run_gemm(cuda_backend, oracle, FP32, 0, 0, 100, 90, 333, 1.0, 0.0, seed=3)
        assert np.array_equal(got, want)
    finally:
        oracle.set_mode(oracle.MODE_PAS)


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 256, 512), (130, 70, 200), (1000, 300, 77), (512, 1024, 2048),
                                   (132, 72, 200), (1000, 300, 76), (260, 520, 36)])  # ragged but 16-byte aligned: TMA-fed kernel
def test_gemm_tensor_core_modes(cuda_backend, oracle, mode, tol, ta, tb, M, N, K):
    got, want, C0 = run_gemm(cuda_backend, oracle, mode, ta, tb, M, N, K, 1.0, 0.0, seed=K)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)
    got, want, C0 = run_gemm(cuda_backend, oracle, mode, ta, tb, M, N, K, 0.5, 2.0, seed=K + 1, ldpad=4, off=4)
    assert rel_max(got, want) < tol
    assert np.array_equal(got[:, N:], C0[:, N:])


def test_gemm_3xtf32_is_fp32_accurate(cuda_backend, oracle):
    """3xTF32 error vs fp64 truth is of the same order as the fp32 CPU reference's own error."""
    rng = np.random.default_rng(5)
    M, N, K = 256, 256, 4608  # the deepest YOLOv3 contraction (512->1024 3x3)
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    cu = cuda_backend
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=M * N)
    cu.set_math_mode(X3)
    cu.gemm(False, False, M, N, K, 1.0, dA.ptr, 0, K, dB.ptr, 0, N, 0.0, dC.ptr, 0, N)
    got = dC.read((M, N))
    truth = oracle.dgemm_truth(0, 0, M, N, K, 1.0, A, K, B, N)
    cpu = np.zeros((M, N), np.float32)
    oracle.sgemm(0, 0, M, N, K, 1.0, A, K, B, N, 0.0, cpu, N)
    err_gpu, err_cpu = rel_max(got, truth), rel_max(cpu, truth)
    assert err_gpu < TOL_3XTF32, (err_gpu, err_cpu)
    assert rel_max(got, cpu) < TOL_3XTF32


@pytest.mark.parametrize("mode", [FP32, X3])
def test_gemm_strided_batched(cuda_backend, oracle, mode):
    cu = cuda_backend
    rng = np.random.default_rng(7)
    b, M, N, K = 5, 32, 169, 288  # conv-forward shaped: weights stride 0, per-image col and out
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (b, K, N)).astype(np.float32)
    C = np.zeros((b, M, N), np.float32)
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, C)
    cu.set_math_mode(mode)
    cu.gemmStridedBatched(False, False, M, N, K, 1.0, dA.ptr, 0, K, 0, dB.ptr, 0, N, K * N, 0.0, dC.ptr, 0, N, M * N, b)
    want = oracle.sgemm_strided_batched(0, 0, M, N, K, 1.0, A, K, 0, B, N, K * N, 0.0, C.copy(), N, M * N, b)
    assert rel_max(dC.read(C.shape), want) < (2e-6 if mode == FP32 else TOL_3XTF32)
    # wgrad shaped: NT, beta=1, strideC=0 accumulates over the batch (nConvolutionLayer.pas:1228-1232)
    for Kk in (169, 256):
        D = rng.uniform(-1, 1, (b, M, Kk)).astype(np.float32)
        Col = rng.uniform(-1, 1, (b, 27, Kk)).astype(np.float32)
        W0 = rng.uniform(-1, 1, (M, 27)).astype(np.float32)
        dD, dCol, dW = Dev(cu, D), Dev(cu, Col), Dev(cu, W0)
        cu.gemmStridedBatched(False, True, M, 27, Kk, 1.0, dD.ptr, 0, Kk, M * Kk, dCol.ptr, 0, Kk, 27 * Kk, 1.0, dW.ptr, 0,
                              27, 0, b)
        want = oracle.sgemm_strided_batched(0, 1, M, 27, Kk, 1.0, D, Kk, M * Kk, Col, Kk, 27 * Kk, 1.0, W0.copy(), 27, 0, b)
        assert rel_max(dW.read(W0.shape), want) < (5e-6 if mode == FP32 else TOL_3XTF32)
    cu.set_math_mode(X3)


def test_gemm_empty_and_degenerate(cuda_backend):
    cu = cuda_backend
    d = Dev(cu, np.ones(16, np.float32))
    cu.gemm(False, False, 0, 4, 4, 1.0, d.ptr, 0, 4, d.ptr, 0, 4, 0.0, d.ptr, 0, 4)  # M = 0: no-op
    cu.gemm(False, False, 4, 4, 0, 1.0, d.ptr, 0, 4, d.ptr, 0, 4, 1.0, d.ptr, 0, 4)  # K = 0: C unchanged (beta=1)
    assert np.array_equal(d.read(), np.ones(16, np.float32))


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32), (3, TOL_3XTF32)])   # 3 = BF16x3
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0)])
def test_gemm_8192_against_the_oracle_on_a_row_sample(cuda_backend, oracle, mode, tol, ta, tb):
    """BASELINE configs[1] at its largest size, M = N = K = 8192, NN / NT / TN: the GPU computes the whole product, the
    oracle (same cblas_sgemm routine, same transpose flags) the 24 rows of a sample — 3.2 GFLOP of CPU work."""
    cu = cuda_backend
    n = 8192
    rng = np.random.default_rng(8192 + 2 * ta + tb)
    A = rng.uniform(-1, 1, (n, n)).astype(np.float32)   # [M][K], or [K][M] when ta
    B = rng.uniform(-1, 1, (n, n)).astype(np.float32)   # [K][N], or [N][K] when tb
    dA, dB, dC = Dev(cu, A), Dev(cu, B), Dev(cu, n=n * n)
    cu.set_math_mode(mode)
    try:
        cu.gemm(ta, tb, n, n, n, 1.0, dA.ptr, 0, n, dB.ptr, 0, n, 0.0, dC.ptr, 0, n)
    finally:
        cu.set_math_mode(X3)
    got = dC.read((n, n))
    rows = np.sort(rng.choice(n, 24, replace=False))
    # the sampled rows of op(A) as a matrix with the SAME storage order (so the oracle runs the same routine)
    Asub = np.ascontiguousarray(A[:, rows]) if ta else np.ascontiguousarray(A[rows, :])
    want = np.zeros((rows.size, n), np.float32)
    oracle.sgemm(ta, tb, rows.size, n, n, 1.0, Asub, rows.size if ta else n, B, n, 0.0, want, n)
    e1, e2 = rel_max(got[rows], want), rel_elem(got[rows], want, floor=1.0)
    # north-star metric (max|g-r| / max|r|) inside the gate; the stricter element metric (error over max(|r|, rms)) is
    # reported with a 4x allowance: the tensor core's truncating fp32 accumulate gives an error that grows with K
    # (profiles/r02_accuracy_vs_k.txt: 5.7e-5 of max|r| at K = 8192), and max|r| is ~3.5 rms here
    print(f"K=8192 ta={ta} tb={tb} mode={mode}: rel_max {e1:.2e}, rel_elem {e2:.2e}")
    assert e1 < tol and e2 < 4 * tol, (e1, e2)


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32), (3, TOL_3XTF32)])
@pytest.mark.parametrize("ta,tb,alpha,beta", [(0, 0, 1.0, 0.0), (0, 1, 0.5, 1.0), (1, 0, 1.0, 1.0), (1, 1, 2.0, 0.0)])
def test_large_gemm_on_the_convolution_kernel(cuda_backend, oracle, mode, tol, ta, tb, alpha, beta):
    """Products with a pair tile for every TPC run as a 1x1 convolution on the CTA-pair kernel (conv_nhwc.cu,
    gemm_on_conv): all four transposes, alpha, beta in {0, 1}, a padded C, against the oracle's cblas_sgemm."""
    cu = cuda_backend
    M, N, K = 2304, 2112, 512       # 9 x 9 pair tiles (ragged in N: 2112 = 8 x 256 + 64)
    got, want, C0 = run_gemm(cu, oracle, mode, ta, tb, M, N, K, alpha, beta, seed=77 + ta + 2 * tb, ldpad=4)
    assert "k_conv_nhwc" in cu.lastConvKernel()
    e1, e2 = rel_max(got[:, :N], want[:, :N]), rel_elem(got[:, :N], want[:, :N], floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)
    assert np.array_equal(got[:, N:], C0[:, N:])   # the padding columns of C are untouched
