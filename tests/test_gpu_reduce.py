"""GPU parity of the per-channel reductions and the batch-norm passes against the oracle."""
import numpy as np
import pytest

from util import Dev, rel_elem, rel_max

pytestmark = pytest.mark.gpu

SHAPES = [(2, 3, 169), (4, 32, 1024), (128, 6, 64), (1, 5, 7), (3, 4, 1), (2, 2, 70000)]


@pytest.mark.parametrize("B,C,HW", SHAPES)
def test_reductions(cuda_backend, oracle, B, C, HW):
    cu = cuda_backend
    rng = np.random.default_rng(B + C + HW)
    x = rng.uniform(-1, 1, (B, C, HW)).astype(np.float32)
    d = rng.uniform(-1, 1, (B, C, HW)).astype(np.float32)
    acc0 = rng.uniform(-1, 1, C).astype(np.float32)
    dx, dd = Dev(cu, x, pad_front=1), Dev(cu, d, pad_front=1)
    dacc = Dev(cu, acc0)
    cu.backwardBias(C, dacc.ptr, x.size, dx._base, 1, 1, B)
    want = oracle.add_sums(acc0.copy(), x, B)
    assert rel_max(dacc.read(), want) < 1e-5
    dacc.write(acc0)
    cu.addDots(x.size, C, B, dx._base, dd._base, 1, dacc.ptr)
    assert rel_max(dacc.read(), oracle.add_dots(acc0.copy(), x, d, B)) < 1e-5
    if B * HW > 1:
        dm, dv = Dev(cu, n=C), Dev(cu, n=C)
        cu.meansAndVars(x.size, C, B, dx._base, 1, dm.ptr, dv.ptr)
        m, v = oracle.means_and_vars(x, B, C)
        assert rel_max(dm.read(), m) < 1e-5 and rel_max(dv.read(), v) < 1e-5
        dm.write(m)
        dv.write(v)
        dmd, dvd = Dev(cu, n=C), Dev(cu, n=C)
        cu.meansAndVarsDelta(x.size, C, B, dd._base, dx._base, 1, dm.ptr, dv.ptr, dmd.ptr, dvd.ptr)
        md, vd = oracle.mean_and_variance_delta(d, x, m, v, B)
        assert rel_max(dmd.read(), md) < 1e-5 and rel_max(dvd.read(), vd) < 1e-5
        dmd.write(md)
        dvd.write(vd)
        cu.normalizeDelta(x.size, C, B, dd._base, dx._base, 1, dm.ptr, dv.ptr, dmd.ptr, dvd.ptr)
        want = oracle.normalize_delta(d.copy(), x, m, v, md, vd, B)
        assert rel_max(dd.read(x.shape), want) < 2e-6 and rel_elem(dd.read(x.shape), want) < 1e-3


@pytest.mark.parametrize("B,C,HW", [(3, 6, 20), (8, 32, 256), (2, 5, 169)])
@pytest.mark.parametrize("act", ["LEAKY", "LINEAR", "LOGISTIC"])
def test_fused_batchnorm(cuda_backend, oracle, B, C, HW, act):
    cu = cuda_backend
    a = oracle.ACT[act]
    rng = np.random.default_rng(B * C)
    x = rng.uniform(-1, 2, (B, C, HW)).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, C).astype(np.float32)
    bi = rng.uniform(-0.2, 0.2, C).astype(np.float32)
    rm = rng.uniform(-0.1, 0.1, C).astype(np.float32)
    rv = rng.uniform(0.5, 1.5, C).astype(np.float32)
    # oracle: nbaselayer.pas:336-370 sequence
    m, v = oracle.means_and_vars(x, B, C)
    xn = oracle.block_normalize(x.copy(), m, v, B)
    y = oracle.activate_array(oracle.forward_bias(oracle.forward_scale(xn.copy(), sc, B), bi, B), a)
    rm2 = oracle.axpy(np.float32(0.1), m, oracle.scal(np.float32(0.9), rm.copy()))
    rv2 = oracle.axpy(np.float32(0.1), v, oracle.scal(np.float32(0.9), rv.copy()))
    dout, dxc, dxn = Dev(cu, x), Dev(cu, n=x.size), Dev(cu, n=x.size)
    dm, dv, drm, drv, dsc, dbi = Dev(cu, n=C), Dev(cu, n=C), Dev(cu, rm), Dev(cu, rv), Dev(cu, sc), Dev(cu, bi)
    cu.batchNormForwardTraining(B, C, HW, dout.ptr, dxc.ptr, dxn.ptr, dm.ptr, dv.ptr, drm.ptr, drv.ptr, 0.1, dsc.ptr,
                                dbi.ptr, a)
    assert np.array_equal(dxc.read(x.shape), x)
    assert rel_max(dm.read(), m) < 2e-6 and rel_max(dv.read(), v) < 1e-5
    assert rel_max(drm.read(), rm2) < 1e-6 and rel_max(drv.read(), rv2) < 1e-5
    # elementwise metric floors |r| at 1e-3*rms, so 1e-3 there means 1e-6*rms absolute
    assert rel_max(dxn.read(x.shape), xn) < 2e-6 and rel_elem(dxn.read(x.shape), xn) < 1e-3
    assert rel_max(dout.read(x.shape), y) < 2e-6 and rel_elem(dout.read(x.shape), y) < 1e-3
    # inference
    d2 = Dev(cu, x)
    drm.write(rm)
    drv.write(rv)
    cu.batchNormForwardInference(B, C, HW, d2.ptr, drm.ptr, drv.ptr, dsc.ptr, dbi.ptr, a)
    yi = oracle.activate_array(
        oracle.forward_bias(oracle.forward_scale(oracle.block_normalize(x.copy(), rm, rv, B), sc, B), bi, B), a)
    assert rel_max(d2.read(x.shape), yi) < 2e-6 and rel_elem(d2.read(x.shape), yi) < 1e-3
    # backward: DeriveArray + batchNormBack (nbaselayer.pas:372-396)
    g = rng.uniform(-1, 1, x.shape).astype(np.float32)
    gd = oracle.gradient_array(y, a, g.copy())
    su0 = rng.uniform(-1, 1, C).astype(np.float32)
    su = oracle.add_dots(su0.copy(), xn, gd, B)
    gs = oracle.forward_scale(gd.copy(), sc, B)
    md, vd = oracle.mean_and_variance_delta(gs, x, m, v, B)
    want = oracle.normalize_delta(gs.copy(), x, m, v, md, vd, B)
    dy, dg, dsu, dmd, dvd = Dev(cu, y), Dev(cu, g), Dev(cu, su0), Dev(cu, n=C), Dev(cu, n=C)
    dm.write(m)
    dv.write(v)
    dxc.write(x)
    dxn.write(xn)
    cu.batchNormBackward(B, C, HW, dy.ptr, a, dg.ptr, dxc.ptr, dxn.ptr, dm.ptr, dv.ptr, dsc.ptr, dsu.ptr, dmd.ptr,
                         dvd.ptr)
    assert rel_max(dsu.read(), su) < 1e-5
    assert rel_max(dmd.read(), md) < 1e-5 and rel_max(dvd.read(), vd) < 1e-5
    assert rel_max(dg.read(x.shape), want) < 1e-5
