"""CPU tests of the oracle (oracle/): hand-derived vectors, fp64 cross-checks, frozen fixtures.
The reference has no golden vectors for this path — parity is unpinned by reference outputs."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLD, "oracle_fixtures.npz"))


def test_im2col_hand_vectors(oracle):
    cases = json.load(open(os.path.join(GOLD, "hand_im2col.json")))
    assert len(cases) >= 7
    for cs in cases:
        im = np.array(cs["im"], dtype=np.float32)
        col = oracle.im2col(im, cs["kh"], cs["kw"], cs["ph"], cs["pw"], cs["sy"], cs["sx"], cs["dy"], cs["dx"])
        want = np.array(cs["col"], dtype=np.float32)
        assert col.shape == want.shape
        assert np.array_equal(col, want)


def test_im2col_literal_rows(oracle):
    # 1x4x4 image 1..16, 3x3, pad 1, stride 1: rows typed by hand from ntensors.pas:11436-11465
    im = np.arange(1, 17, dtype=np.float32).reshape(1, 4, 4)
    col = oracle.im2col(im, 3, 3, 1, 1, 1, 1)
    assert col.shape == (9, 16)
    assert col[0].tolist() == [0, 0, 0, 0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11]       # (kr,kc)=(0,0)
    assert col[4].tolist() == list(range(1, 17))                                          # centre tap
    assert col[8].tolist() == [6, 7, 8, 0, 10, 11, 12, 0, 14, 15, 16, 0, 0, 0, 0, 0]     # (2,2)
    assert col[1].tolist() == [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]        # (0,1)


def test_col2im_is_adjoint_and_accumulates(oracle):
    rng = np.random.default_rng(3)
    for (c, h, w, k, p, s) in [(2, 5, 7, 3, 1, 1), (3, 8, 6, 3, 1, 2), (1, 4, 4, 1, 0, 1), (2, 7, 7, 2, 0, 2)]:
        x = rng.integers(-3, 4, (c, h, w)).astype(np.float32)
        col = oracle.im2col(x, k, k, p, p, s, s)
        y = rng.integers(-3, 4, col.shape).astype(np.float32)
        im0 = rng.integers(-3, 4, (c, h, w)).astype(np.float32)
        im = oracle.col2im(y, c, h, w, k, k, p, p, s, s, im=im0.copy())
        # small integers: sums are exact, so <im2col(x), y> == <x, col2im(y)> exactly, and col2im ADDS to im0
        assert float((col.astype(np.float64) * y).sum()) == float(((im - im0).astype(np.float64) * x).sum())


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_sgemm_against_fp64(oracle, mode, ta, tb):
    oracle.set_mode(mode)
    rng = np.random.default_rng(11)
    M, N, K = 33, 47, 129
    A = rng.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    C = C0.copy()
    oracle.sgemm(ta, tb, M, N, K, 0.5, A, A.shape[1], B, B.shape[1], 2.0, C, N)
    truth = oracle.dgemm_truth(ta, tb, M, N, K, 0.5, A, A.shape[1], B, B.shape[1], 2.0, C0, N)
    assert np.abs(C - truth).max() < 2e-5
    oracle.set_mode(0)


def test_sgemm_pas_mode_is_sequential_mul_add(oracle):
    # bit-exact against a pure numpy float32 restatement of s_nn (ntensors.pas:2093-2099)
    oracle.set_mode(0)
    rng = np.random.default_rng(5)
    M, N, K = 4, 9, 17
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    C = np.zeros((M, N), np.float32)
    oracle.sgemm(0, 0, M, N, K, 1.0, A, K, B, N, 0.0, C, N)
    want = np.zeros((M, N), np.float32)
    for i in range(M):
        for kk in range(K):
            want[i] = (np.float32(A[i, kk]) * B[kk]).astype(np.float32) + want[i]
    assert np.array_equal(C, want)


def test_strided_batched_accumulates_when_stride_c_is_zero(oracle):
    rng = np.random.default_rng(6)
    b, M, N, K = 3, 5, 7, 11
    A = rng.uniform(-1, 1, (b, M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (b, N, K)).astype(np.float32)
    C = np.zeros((M, N), np.float32)
    oracle.sgemm_strided_batched(0, 1, M, N, K, 1.0, A, K, M * K, B, K, N * K, 1.0, C, N, 0, b)
    want = sum(A[i].astype(np.float64) @ B[i].astype(np.float64).T for i in range(b))
    assert np.abs(C - want).max() < 1e-5


def test_conv2d_matches_direct_definition(oracle):
    rng = np.random.default_rng(8)
    for (b, c, h, w, f, k, p, s) in [(2, 3, 9, 7, 4, 3, 1, 1), (1, 4, 8, 8, 5, 3, 1, 2), (2, 6, 5, 5, 3, 1, 0, 1)]:
        x = rng.uniform(-1, 1, (b, c, h, w)).astype(np.float32)
        wt = rng.uniform(-1, 1, (f, c * k * k)).astype(np.float32)
        y = oracle.conv2d(x, wt, p, s, s, 1, k, k)
        xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (p, p), (p, p)))
        ww = wt.reshape(f, c, k, k).astype(np.float64)
        oh, ow = y.shape[2:]
        ref = np.zeros(y.shape)
        for i in range(oh):
            for j in range(ow):
                ref[:, :, i, j] = np.einsum("bchw,fchw->bf", xp[:, :, s * i:s * i + k, s * j:s * j + k], ww)
        assert np.abs(y - ref).max() < 1e-5


def test_batchnorm_statistics_are_unbiased_two_pass(oracle):
    rng = np.random.default_rng(9)
    x = rng.uniform(-2, 3, (4, 5, 6, 7)).astype(np.float32)
    m, v = oracle.means_and_vars(x, 4, 5)
    xd = x.astype(np.float64)
    assert np.allclose(m, xd.mean((0, 2, 3)), atol=1e-6)
    assert np.allclose(v, xd.var((0, 2, 3), ddof=1), rtol=1e-5)  # /(N-1)  ntensors.pas:9136
    y = x.copy()
    oracle.block_normalize(y, m, v, 4)
    assert np.allclose(y, (xd - xd.mean((0, 2, 3), keepdims=True)) / np.sqrt(xd.var((0, 2, 3), ddof=1, keepdims=True)),
                       atol=1e-5)


def test_batchnorm_backward_formulas(oracle):
    rng = np.random.default_rng(10)
    B, C, H, W = 3, 4, 5, 6
    x = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    d = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    m, v = oracle.means_and_vars(x, B, C)
    md, vd = oracle.mean_and_variance_delta(d, x, m, v, B)
    xd, dd = x.astype(np.float64), d.astype(np.float64)
    mm, vv = m.astype(np.float64)[None, :, None, None], v.astype(np.float64)[None, :, None, None]
    assert np.allclose(md, (dd.sum((0, 2, 3)) * (-1 / np.sqrt(v.astype(np.float64)))), rtol=1e-5, atol=1e-6)
    assert np.allclose(vd, ((dd * (xd - mm)).sum((0, 2, 3)) * -0.5 * v.astype(np.float64) ** -1.5), rtol=1e-5, atol=1e-6)
    out = oracle.normalize_delta(d.copy(), x, m, v, md, vd, B)
    n = B * H * W
    want = dd / np.sqrt(vv) + (2 * vd[None, :, None, None] * (xd - mm) + md[None, :, None, None]) / n
    assert np.allclose(out, want, rtol=1e-5, atol=1e-6)


def test_activation_quirks(oracle):
    x = np.array([-1.0, 0.0, 3.0, 6.0, 6.5], np.float32)
    y = oracle.activate_array(x.copy(), oracle.ACT["RELU6"])
    assert y.tolist() == [0.0, 0.0, 3.0, 6.0, 0.0]  # 0 above 6 — reference behaviour (nactivation.pas:316)
    y = oracle.activate_array(np.array([-2.0, 2.0], np.float32), oracle.ACT["LEAKY"])
    assert y[1] == 2.0 and abs(y[0] + 0.2) < 1e-7
    y = oracle.activate_array(np.array([0.0], np.float32), oracle.ACT["LOGISTIC"])
    assert y[0] == 0.5


def test_frozen_fixtures_reproduce(oracle, fixtures):
    """The committed fixtures are what the current oracle produces (guards both against drift)."""
    from oracle.ref_layers import RefConvLayer
    f = fixtures
    for mode, tag in ((0, "pas"), (1, "fma")):
        oracle.set_mode(mode)
        A, B, C0 = f[f"{tag}_sgemm_A"], f[f"{tag}_sgemm_B"], f[f"{tag}_sgemm_C0"]
        for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
            a = np.ascontiguousarray(A.T) if ta else A
            b = np.ascontiguousarray(B.T) if tb else B
            c = C0.copy()
            oracle.sgemm(ta, tb, 37, 29, 53, 0.75, a, a.shape[1], b, b.shape[1], 0.5, c, 29)
            assert np.array_equal(c, f[f"{tag}_sgemm_{ta}{tb}"])
        for bn in (0, 1):
            p = f"{tag}_conv_bn{bn}_"
            L = RefConvLayer(3, 5, 9, 8, 6, 3, 2, 1, oracle.ACT["LEAKY"], bn, f[p + "w"], f[p + "bias"],
                             f[p + "scales"], f[p + "rm"], f[p + "rv"])
            assert np.array_equal(L.forward(f[p + "x"], training=False), f[p + "y_inf"])
            assert np.array_equal(L.forward(f[p + "x"], training=True), f[p + "y_tr"])
            L.delta[...] = f[p + "g"]
            up = np.zeros_like(f[p + "x"])
            L.backward(up)
            assert np.array_equal(up, f[p + "dx"])
            assert np.array_equal(L.weight_updates, f[p + "dw"])
    oracle.set_mode(0)


def test_fuse_batchnorm_equals_inference_batchnorm(oracle, fixtures):
    from oracle.ref_layers import RefConvLayer
    f, p = fixtures, "pas_conv_bn1_"
    L = RefConvLayer(3, 5, 9, 8, 6, 3, 2, 1, oracle.ACT["LEAKY"], 1, f[p + "w"], f[p + "bias"], f[p + "scales"],
                     f[p + "rm"], f[p + "rv"])
    y0 = L.forward(f[p + "x"], training=False).copy()
    L.fuse_batchnorm()
    y1 = L.forward(f[p + "x"], training=False)
    assert np.abs(y0 - y1).max() < 1e-5
