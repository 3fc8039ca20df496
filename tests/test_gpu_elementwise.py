"""GPU parity of the HBM-bound ops, through the C ABI, against the oracle (bit-exact where the arithmetic is
a single IEEE operation; 1e-6 relative for transcendental activations)."""
import os

import numpy as np
import pytest

from util import Dev, rel_elem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EXACT_ACTS = ["RELU", "RELU6", "RELIE", "LINEAR", "RAMP", "PLSE", "REVLEAKY", "LEAKY", "HARDTAN", "LHTAN", "STAIR"]
APPROX_ACTS = ["LOGISTIC", "TANH", "ELU", "LOGGY", "SELU", "GELU", "SWISH"]


def test_activations_against_fixture(cuda_backend, oracle):
    cu = cuda_backend
    f = np.load(os.path.join(GOLD, "oracle_fixtures.npz"))
    x = f["act_x"]
    for name in EXACT_ACTS + APPROX_ACTS:
        for pad in (0, 1, 3):  # misaligned offsets exercise the scalar head/tail
            d = Dev(cu, x, pad_front=pad)
            cu.ActivateArray(x.size, d._base, pad, oracle.ACT[name])
            y = d.read()
            want = f[f"act_{name}"]
            if name in EXACT_ACTS:
                assert np.array_equal(y, want), name
            else:
                assert rel_elem(y, want) < 2e-6, (name, rel_elem(y, want))
        if name != "SWISH":
            out = Dev(cu, f[f"act_{name}"])
            dl = Dev(cu, np.ones_like(x))
            cu.DeriveArray(x.size, out.ptr, 0, oracle.ACT[name], dl.ptr)
            g = dl.read()
            assert rel_elem(g, f[f"grad_{name}"]) < 2e-6, name


@pytest.mark.parametrize("B,C,HW", [(2, 3, 169), (3, 5, 1024), (1, 7, 13), (4, 32, 4096), (2, 4, 1), (1, 1, 5000),
                                    (64, 40, 1), (3, 10, 1), (300, 1024, 1)])  # HW = 1: the connected layers' row-broadcast path
def test_bias_scale_normalize(cuda_backend, oracle, B, C, HW):
    cu = cuda_backend
    rng = np.random.default_rng(B * 100 + C)
    x = rng.uniform(-1, 1, (B, C, HW)).astype(np.float32)
    b = rng.uniform(-1, 1, C).astype(np.float32)
    s = rng.uniform(0.5, 1.5, C).astype(np.float32)
    m = rng.uniform(-0.5, 0.5, C).astype(np.float32)
    v = rng.uniform(0.0, 1.5, C).astype(np.float32)
    v[0] = 0.0  # exercises the epsilon clamp
    for pad in (0, 2):
        db, ds, dm, dv = Dev(cu, b), Dev(cu, s), Dev(cu, m), Dev(cu, v)
        d = Dev(cu, x, pad_front=pad)
        cu.forwardBias(x.size, d._base, pad, C, db.ptr, 1, B)
        assert np.array_equal(d.read(x.shape), oracle.forward_bias(x.copy(), b, B))
        d.write(x)
        cu.forwardScale(x.size, d._base, pad, C, ds.ptr, 1, B)
        assert np.array_equal(d.read(x.shape), oracle.forward_scale(x.copy(), s, B))
        d.write(x)
        cu.forwardScaleAdd(x.size, d._base, pad, C, ds.ptr, db.ptr, 1, B)
        assert np.array_equal(d.read(x.shape), oracle.forward_bias(oracle.forward_scale(x.copy(), s, B), b, B))
        d.write(x)
        cu.normalize(C, x.size, B, dm.ptr, 1, dv.ptr, 1, d._base, pad)
        want = oracle.block_normalize(x.copy(), m, v, B)
        assert rel_elem(d.read(x.shape), want) < 1e-6


def test_blas1_ops(cuda_backend, oracle):
    cu = cuda_backend
    rng = np.random.default_rng(1)
    n = 10007
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = rng.uniform(-1, 1, n).astype(np.float32)
    dx, dy = Dev(cu, x), Dev(cu, y)
    cu.axpy(n, 0.37, dx.ptr, 0, 1, dy.ptr, 0, 1)
    assert np.array_equal(dy.read(), oracle.axpy(np.float32(0.37), x, y.copy()))  # saxpy_pas: mul then add
    # strided + offset
    dy.write(y)
    cu.axpy(1000, -1.5, dx.ptr, 3, 2, dy.ptr, 5, 3)
    want = y.copy()
    want[5:5 + 3000:3] = (np.float32(-1.5) * x[3:3 + 2000:2]).astype(np.float32) + want[5:5 + 3000:3]
    assert np.array_equal(dy.read(), want)
    dy.write(y)
    cu.scale(n, 0.9, dy.ptr, 1)
    assert np.array_equal(dy.read(), (np.float32(0.9) * y).astype(np.float32))
    cu.scale(n, 0.0, dy.ptr, 1)
    assert not dy.read().any()
    cu.fill(100, dy.ptr, 7, 2.5, 2)
    r = dy.read()
    assert (r[7:207:2] == 2.5).all() and r[8] == 0 and r[:7].sum() == 0
    dz = Dev(cu, n=n)
    cu.copy(n, dx.ptr, 0, 1, dz.ptr, 0, 1)
    assert np.array_equal(dz.read(), x)
    cu.copy(50, dx.ptr, 1, 2, dz.ptr, 10, 1)
    assert np.array_equal(dz.read()[10:60], x[1:101:2])
    dy.write(y)
    for op, fn in ((cu.addvv, np.add), (cu.subvv, np.subtract), (cu.mulvv, np.multiply)):
        op(n, dx.ptr, 0, 1, dy.ptr, 0, 1, dz.ptr, 0, 1)
        assert np.array_equal(dz.read(), fn(x, y))
    op = cu.addvv  # in place on src1 (shortcut layers do this)
    op(n - 1, dx.ptr, 1, 1, dy.ptr, 0, 1, dx.ptr, 1, 1)
    assert np.array_equal(dx.read()[1:], x[1:] + y[:-1])
    dx.write(x)
    cu.fmavss(n, dx.ptr, 0, 2.0, -1.0, dz.ptr)
    assert np.array_equal(dz.read(), x * np.float32(2) - np.float32(1))
    cu.clamp(n, 0.25, dx.ptr, dz.ptr)
    assert np.array_equal(dz.read(), np.clip(x, -0.25, 0.25))
    cu.inverseSqrt(n, 0.0, dx.ptr, dz.ptr)
    assert rel_elem(dz.read(), 1 / np.sqrt(np.maximum(x, np.float32(1e-6)))) < 1e-6
    d3 = Dev(cu, y)
    cu.fmavv(n, dx.ptr, 0, 1, dy.ptr, 0, 1, d3.ptr, 0, 1, dz.ptr, 0, 1)
    assert np.array_equal(dz.read(), (x * y).astype(np.float32) + y)


def test_sgd_update_matches_reference_sequence(cuda_backend, oracle):
    cu = cuda_backend
    rng = np.random.default_rng(2)
    for n in (864, 147456 + 3, 7):
        w = rng.uniform(-1, 1, n).astype(np.float32)
        u = rng.uniform(-1, 1, n).astype(np.float32)
        dw, du = Dev(cu, w), Dev(cu, u)
        rate, mom, dec = np.float32(0.001 / 128), np.float32(0.9), np.float32(-0.0001 * 128)
        cu.sgdUpdate(n, dw.ptr, du.ptr, float(rate), float(mom), float(dec))
        uu = oracle.axpy(dec, w, u.copy())          # weight_updates.axpy(-decay*batch, weights)
        ww = oracle.axpy(rate, uu, w.copy())        # weights.axpy(lr/batch, weight_updates)
        uu = oracle.scal(mom, uu)                   # weight_updates.Multiply(momentum)
        assert np.array_equal(dw.read(), ww) and np.array_equal(du.read(), uu)


def test_losses_softmax_dropout(cuda_backend):
    cu = cuda_backend
    rng = np.random.default_rng(3)
    n = 1000
    p = rng.uniform(0.01, 0.99, n).astype(np.float32)
    t = (rng.uniform(0, 1, n) > 0.7).astype(np.float32)
    dp, dt, dd, de = Dev(cu, p), Dev(cu, t), Dev(cu, n=n), Dev(cu, n=n)
    cu.costL2(n, dp.ptr, dt.ptr, dd.ptr, de.ptr)
    assert np.array_equal(dd.read(), t - p) and np.array_equal(de.read(), (t - p) * (t - p))
    cu.crossEntropySoftmax(n, dp.ptr, dt.ptr, dd.ptr, de.ptr)
    assert np.allclose(de.read(), np.where(t != 0, -np.log(p), 0), rtol=1e-6)
    cu.crossEntropyLogistic(n, dp.ptr, dt.ptr, dd.ptr, de.ptr)
    assert np.allclose(de.read(), -t * np.log(p) - (1 - t) * np.log(1 - p), rtol=1e-5, atol=1e-7)
    x = rng.uniform(-3, 3, (4, 10)).astype(np.float32)
    dx, dy = Dev(cu, x), Dev(cu, n=x.size)
    cu.softmaxBatch(10, dx.ptr, 0, 4, 10, 1, 10, 1, 1.0, dy.ptr, 0)
    e = np.exp(x - x.max(1, keepdims=True))
    assert np.allclose(dy.read(x.shape), e / e.sum(1, keepdims=True), rtol=1e-5)
    src = rng.uniform(-1, 1, n).astype(np.float32)
    ds, dr, do = Dev(cu, src), Dev(cu, n=n), Dev(cu, n=n)
    cu.forwardDropout(n, ds.ptr, 0.2, 1.25, dr.ptr, do.ptr, seed=1234)
    r, o = dr.read(), do.read()
    assert (r >= 0).all() and (r < 1).all() and 0.1 < (r < 0.2).mean() < 0.3
    assert np.array_equal(o, np.where(r < 0.2, 0, src * np.float32(1.25)))
    cu.backwardDropout(n, ds.ptr, 0.2, 1.25, dr.ptr, do.ptr)
    assert np.array_equal(do.read(), o)


def test_maxpool_upsample_concat(cuda_backend):
    cu = cuda_backend
    rng = np.random.default_rng(4)
    B, C, H, W = 2, 3, 8, 8
    x = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    dx = Dev(cu, x)
    oh = ow = 4
    dout = Dev(cu, n=B * C * oh * ow)
    didx = cu.createDeviceBuffer(B * C * oh * ow * 8)
    cu.forwardMaxPool(B, C, oh, ow, dx.ptr, C, H, W, 2, 2, 0, 2, didx, dout.ptr)
    want = x.reshape(B, C, 4, 2, 4, 2).max((3, 5))
    assert np.array_equal(dout.read(want.shape), want)
    g = rng.uniform(-1, 1, want.shape).astype(np.float32)
    dg, dd = Dev(cu, g), Dev(cu, n=x.size)
    cu.backwardMaxPool(B, C, oh, ow, dd.ptr, didx, dg.ptr)
    back = dd.read(x.shape)
    mask = (x == np.repeat(np.repeat(want, 2, 2), 2, 3))
    assert np.allclose(back, mask * np.repeat(np.repeat(g, 2, 2), 2, 3))
    cu.freeDeviceBuffer(didx)
    # upsample x2 (nupsamplelayer.pas:196-201: outHeight/outWidth are the low-res sizes)
    lo = rng.uniform(-1, 1, (B, C, 4, 4)).astype(np.float32)
    dlo, dhi = Dev(cu, lo), Dev(cu, n=B * C * 64)
    cu.upSample(B, C, 4, 4, dlo.ptr, 2, 1, 1.0, dhi.ptr)
    assert np.array_equal(dhi.read((B, C, 8, 8)), np.repeat(np.repeat(lo, 2, 2), 2, 3))
    dacc = Dev(cu, np.ones_like(lo))
    cu.upSample(B, C, 4, 4, dacc.ptr, 2, 0, 1.0, dx.ptr)
    assert np.allclose(dacc.read(lo.shape), 1 + x.reshape(B, C, 4, 2, 4, 2).sum((3, 5)), atol=1e-6)
    # per-image channel concat (route)
    a = rng.uniform(-1, 1, (B, 2, 5, 5)).astype(np.float32)
    b = rng.uniform(-1, 1, (B, 3, 5, 5)).astype(np.float32)
    da, db, dc = Dev(cu, a), Dev(cu, b), Dev(cu, n=B * 5 * 25)
    cu.concatChannels(B, 2, 25, da.ptr, 5, 0, dc.ptr)
    cu.concatChannels(B, 3, 25, db.ptr, 5, 2, dc.ptr)
    assert np.array_equal(dc.read((B, 5, 5, 5)), np.concatenate([a, b], 1))
