"""Parity of the channels-innermost (NHWC) convolution path — TMA im2col-mode activations, TMA weights, TMA store —
against the oracle (the reference's CPU im2col + sgemm, source/ntensors.pas:8252-8349), through the C ABI.
The oracle works on NCHW tensors; the tests convert at the edges (host side, numpy transposes) so the kernel is
checked on exactly the reference's inputs and outputs."""
import numpy as np
import pytest

from util import TOL_3XTF32, TOL_TF32, Dev, rel_elem, rel_max

pytestmark = pytest.mark.gpu
FP32, X3, TF32, BF3 = 0, 1, 2, 3
NCHW, NHWC = 0, 1


def to_nhwc(x, pitch=None):
    B, C, H, W = x.shape
    pitch = pitch or C
    out = np.full((B, H, W, pitch), 555.0, np.float32)  # the pad lanes must never be read as data
    out[..., :C] = x.transpose(0, 2, 3, 1)
    return out


def from_nhwc(y, C):
    return np.ascontiguousarray(y[..., :C].transpose(0, 3, 1, 2))


def conv_case(rng, B, C, H, W, F, k):
    x = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    s = np.sqrt(2.0 / (k * k * C))
    w = rng.uniform(-s, s, (F, C * k * k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, F).astype(np.float32)
    return x, w, b


# (B, C, H, W, F, k, pad, stride, inPitch, outPitch)   pitch 0 = dense
GEOMS = [
    (2, 32, 16, 16, 32, 3, 1, 1, 0, 0),      # one 32-channel block per tap, 512 pixels = 4 tiles
    (2, 64, 13, 13, 128, 3, 1, 1, 0, 0),     # 13-wide rows: tiles cross rows and images, ragged last tile
    (3, 32, 26, 26, 64, 3, 1, 2, 0, 0),      # stride 2 (traversal stride of the TMA walk)
    (1, 128, 52, 52, 64, 1, 0, 1, 0, 0),     # 1x1 bottleneck
    (2, 48, 12, 10, 40, 3, 1, 1, 0, 0),      # C = 48: second 32-block half empty (TMA zero fill), F = 40 (BN 48)
    (2, 36, 9, 9, 24, 3, 1, 1, 40, 28),      # C = 36 inside pitch-40 rows, output into pitch-28 rows
    (1, 16, 20, 24, 272, 3, 1, 1, 0, 0),     # two filter tiles (256 + 16)
    (2, 64, 8, 8, 96, 5, 2, 1, 0, 0),        # 5x5
    (1, 32, 17, 15, 32, 3, 0, 1, 0, 0),      # no padding
    (2, 32, 14, 14, 64, 3, 2, 1, 0, 0),      # padding 2 with a 3x3 kernel (output larger than input)
    (1, 32, 31, 29, 48, 3, 1, 2, 0, 64),     # stride 2 on odd sizes, output slice of a 64-wide concat buffer
    (4, 256, 13, 13, 512, 3, 1, 1, 0, 0),    # the YOLO 13x13 shape (K = 2304), 6 tiles x 2 filter tiles
    # enough 128-pixel tiles for every TPC: these run on CTA PAIRS (tcgen05.mma.cta_group::2, half a weight tile per CTA)
    (8, 64, 52, 52, 128, 3, 1, 1, 0, 0),     # 169 tiles (odd: the last pair has one empty half), BN = 128
    (7, 32, 52, 52, 64, 3, 1, 2, 0, 0),      # stride 2, BN = 64 (32 weight rows per CTA), 37 pair tiles x ... single n tile
    (10, 96, 40, 40, 288, 1, 0, 1, 0, 320),  # 1x1, two filter tiles (256 + 32), output slice of a 320-wide buffer
]


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32), (BF3, TOL_3XTF32)])
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_forward_nhwc(cuda_backend, oracle, mode, tol, geom):
    cu = cuda_backend
    B, C, H, W, F, k, p, s, ip, op = geom
    ip, op = ip or C, op or F
    rng = np.random.default_rng(sum(geom) + 3)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    OH, OW = want.shape[2], want.shape[3]
    dx, dw, db, ds = Dev(cu, to_nhwc(x, ip)), Dev(cu, w), Dev(cu, b), Dev(cu, sc)
    dout = Dev(cu, n=B * OH * OW * op)
    cu.fill(dout.n, dout.ptr, 0, 777.0, 1)
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    cu.set_math_mode(mode)
    try:
        cu.conv2dForwardNhwc(desc, dx.ptr, 0, ip, NHWC, dw.ptr, ds.ptr, db.ptr, oracle.ACT["LEAKY"], dout.ptr, 0, op, NHWC)
        assert "k_conv_nhwc" in cu.lastConvKernel()
    finally:
        cu.set_math_mode(X3)
    raw = dout.read((B, OH, OW, op))
    got = from_nhwc(raw, F)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)
    if op > F:
        assert np.all(raw[..., F:] == 777.0)  # the neighbouring slice of the concat buffer is untouched


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32), (BF3, TOL_3XTF32)])
@pytest.mark.parametrize("B,H", [(3, 13), (10, 52)])   # (10, 52): 212 tiles -> CTA pairs
def test_conv_forward_nhwc_planar_output(cuda_backend, oracle, mode, tol, B, H):
    """Head layers: channels-innermost input, NCHW output planes (255 filters, linear, bias only)."""
    cu = cuda_backend
    C, W, F, k, p, s = 64, H, 255, 1, 0, 1
    rng = np.random.default_rng(77)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_bias(want, b, B)
    dx, dw, db, dout = Dev(cu, to_nhwc(x)), Dev(cu, w), Dev(cu, b), Dev(cu, n=want.size)
    cu.set_math_mode(mode)
    try:
        cu.conv2dForwardNhwc(cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s), dx.ptr, 0, C, NHWC, dw.ptr, None, db.ptr,
                             oracle.ACT["LINEAR"], dout.ptr, 0, 0, NCHW)
    finally:
        cu.set_math_mode(X3)
    got = dout.read(want.shape)
    assert rel_max(got, want) < tol and rel_elem(got, want, floor=1.0) < tol


def test_first_layer_planar_input_to_nhwc(cuda_backend, oracle):
    """The image layer: NCHW 3-channel input, direct fp32 kernel, channels-innermost output — exact fp32 FMA sums."""
    cu = cuda_backend
    B, C, H, W, F, k, p, s = 2, 3, 20, 24, 32, 3, 1, 1
    rng = np.random.default_rng(5)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    oracle.set_mode(oracle.MODE_FMA)
    try:
        want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    finally:
        oracle.set_mode(oracle.MODE_PAS)
    oracle.forward_bias(want, b, B)
    dx, dw, db = Dev(cu, x), Dev(cu, w), Dev(cu, b)
    dout = Dev(cu, n=B * H * W * F)
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    cu.conv2dForwardNhwc(desc, dx.ptr, 0, 0, NCHW, dw.ptr, None, db.ptr, oracle.ACT["LINEAR"], dout.ptr, 0, F, NHWC)
    assert "NHWC" in cu.lastConvKernel()
    assert np.array_equal(from_nhwc(dout.read((B, H, W, F)), F), want)
    # leaky: the kernel multiplies by 0.1f in fp32, the Pascal expression by the double 0.1 (one rounding apart)
    cu.conv2dForwardNhwc(desc, dx.ptr, 0, 0, NCHW, dw.ptr, None, db.ptr, oracle.ACT["LEAKY"], dout.ptr, 0, F, NHWC)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    assert rel_elem(from_nhwc(dout.read((B, H, W, F)), F), want) < 2e-7


def test_layout_companions(cuda_backend):
    """nchw<->nhwc, pitched row copy / add (route, shortcut) and the channels-innermost upsample: bit-exact."""
    cu = cuda_backend
    rng = np.random.default_rng(11)
    B, C, H, W = 3, 44, 7, 9
    x = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    dx = Dev(cu, x)
    pitch = 48
    dn = Dev(cu, n=B * H * W * pitch)
    cu.fill(dn.n, dn.ptr, 0, 9.0, 1)
    cu.nchwToNhwc(B, C, H * W, dx.ptr, dn.ptr, pitch)
    got = dn.read((B, H, W, pitch))
    assert np.array_equal(got[..., :C], x.transpose(0, 2, 3, 1)) and np.all(got[..., C:] == 9.0)
    db = Dev(cu, n=x.size)
    cu.nhwcToNchw(B, C, H * W, dn.ptr, pitch, db.ptr)
    assert np.array_equal(db.read(x.shape), x)
    # route: two sources into slices of one concat buffer; shortcut: a + b
    y = rng.uniform(-1, 1, (B, H, W, 20)).astype(np.float32)
    dy = Dev(cu, y)
    dc = Dev(cu, n=B * H * W * 64)
    cu.rowsCopy(B * H * W, C, dn.ptr, pitch, dc.ptr, 64)
    cu.rowsCopy(B * H * W, 20, dy.ptr, 20, dc.ptr + 44 * 4, 64)
    cat = dc.read((B, H, W, 64))
    assert np.array_equal(cat[..., :44], got[..., :44]) and np.array_equal(cat[..., 44:], y)
    ds = Dev(cu, n=B * H * W * 20)
    cu.rowsAdd(B * H * W, 20, dy.ptr, 20, dc.ptr + 44 * 4, 64, ds.ptr, 20)
    assert np.array_equal(ds.read(y.shape), y + y)
    # upsample x2
    du = Dev(cu, n=B * 2 * H * 2 * W * 20)
    cu.upSampleNhwc(B, 20, H, W, 2, 1.0, dy.ptr, 20, du.ptr, 20)
    assert np.array_equal(du.read((B, 2 * H, 2 * W, 20)), y.repeat(2, axis=1).repeat(2, axis=2))


# ------------------------------------------------------------------ pre-split (bf16 hi / lo planes) activations
SPLIT = 2


def bf16_split(x):
    """hi = bf16_rn(x), lo = bf16_rn(x - hi) as uint16 bit patterns (numpy restatement of the definition in the header)."""
    def rn(v):
        u = np.ascontiguousarray(v, np.float32).view(np.uint32).astype(np.uint64)
        return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)
    hi = rn(x)
    hif = (hi.astype(np.uint32) << 16).view(np.float32)
    return hi, rn(np.asarray(x, np.float32) - hif)


def planes_to_f32(hi, lo):
    return (hi.astype(np.uint32) << 16).view(np.float32) + (lo.astype(np.uint32) << 16).view(np.float32)


def test_split_planes_round_trip(cuda_backend):
    """tnn_nhwc_to_split / tnn_split_to_nhwc against the definition, bit for bit, with padded pitches."""
    cu = cuda_backend
    rng = np.random.default_rng(3)
    rows, C, ip, op = 1000, 40, 44, 48
    x = (rng.standard_normal((rows, ip)) * np.exp(rng.uniform(-20, 20, (rows, ip)))).astype(np.float32)
    dx, dsp, dback = Dev(cu, x), Dev(cu, n=rows * op), Dev(cu, n=rows * C)
    cu.nhwcToSplit(rows, C, dx.ptr, ip, dsp.ptr, op, rows * op)
    planes = dsp.read().view(np.uint16).reshape(2, rows, op)
    hi, lo = bf16_split(x[:, :C])
    assert np.array_equal(planes[0, :, :C], hi) and np.array_equal(planes[1, :, :C], lo)
    cu.splitToNhwc(rows, C, dsp.ptr, op, rows * op, dback.ptr, C)
    back = dback.read((rows, C))
    assert np.array_equal(back, planes_to_f32(hi, lo))
    assert np.abs(back / x[:, :C] - 1).max() < 2.0 ** -16


SPLIT_GEOMS = [
    # (B, C, H, W, F, k, pad, stride, residual)
    (2, 32, 16, 16, 32, 3, 1, 1, False),
    (2, 64, 13, 13, 128, 3, 1, 1, True),      # ragged last tile, shortcut operand
    (3, 32, 26, 26, 64, 3, 1, 2, False),      # stride 2
    (1, 128, 52, 52, 64, 1, 0, 1, False),     # 1x1
    (2, 40, 12, 10, 40, 3, 1, 1, False),      # C = 40: rest of the 32-block zero-filled by the TMA unit; F = 40
    (4, 256, 13, 13, 512, 3, 1, 1, True),     # K = 2304, two filter tiles
    (8, 64, 52, 52, 128, 3, 1, 1, True),      # CTA pairs, odd tile count
    (7, 32, 52, 52, 64, 3, 1, 2, False),      # CTA pairs, stride 2, BN = 64
    (10, 96, 40, 40, 288, 1, 0, 1, False),    # CTA pairs, 1x1, two filter tiles (256 + 32)
]


@pytest.mark.parametrize("geom", SPLIT_GEOMS)
def test_conv_forward_presplit(cuda_backend, oracle, geom):
    """tnn_conv2d_forward_split: bf16 hi/lo planes in, fp32 rows AND planes out (with and without the fused shortcut).
    (1) the fp32 result is within the 1e-4 gate of the oracle; (2) it is BIT-IDENTICAL to the fp32-input BF16x3 kernel,
    whose converter warps derive the same planes in shared memory; (3) the output planes are the split of the fp32 output;
    (4) planes-only and fp32-only outputs write the same bits."""
    cu = cuda_backend
    B, C, H, W, F, k, p, s, with_res = geom
    rng = np.random.default_rng(sum(geom[:8]) + 11)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    OH, OW = want.shape[2], want.shape[3]
    M = B * OH * OW
    res = rng.uniform(-1, 1, (B, F, OH, OW)).astype(np.float32) if with_res else None
    if with_res:
        want = want + res
    dx, dw, db, ds = Dev(cu, to_nhwc(x)), Dev(cu, w), Dev(cu, b), Dev(cu, sc)
    dres = Dev(cu, to_nhwc(res)) if with_res else None
    dxs = Dev(cu, n=x.size)
    n_in = B * H * W * C
    cu.nhwcToSplit(B * H * W, C, dx.ptr, C, dxs.ptr, C, n_in)
    outs = {}
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    act = oracle.ACT["LEAKY"]
    cu.set_math_mode(BF3)
    try:
        for name, (src, lay, plane, f32, spl) in {"f32in": (dx, NHWC, 0, True, False), "both": (dxs, SPLIT, n_in, True, True),
                                                  "planes": (dxs, SPLIT, n_in, False, True),
                                                  "rows": (dxs, SPLIT, n_in, True, False)}.items():
            do, dsp = Dev(cu, n=M * F), Dev(cu, n=M * F)
            cu.fill(do.n, do.ptr, 0, 777.0, 1)
            cu.conv2dForwardSplit(desc, src.ptr, 0, C, lay, plane, dw.ptr, ds.ptr, db.ptr, act,
                                  dres.ptr if with_res else None, 0, F if with_res else 0,
                                  do.ptr if f32 else None, 0, F, NHWC, dsp.ptr if spl else None, 0, F, M * F)
            assert "k_conv_nhwc" in cu.lastConvKernel()
            outs[name] = (do.read((B, OH, OW, F)) if f32 else None,
                          dsp.read().view(np.uint16).reshape(2, M, F) if spl else None)
    finally:
        cu.set_math_mode(X3)
    got = from_nhwc(outs["both"][0], F)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < TOL_3XTF32 and e2 < TOL_3XTF32, (e1, e2)
    assert np.array_equal(outs["both"][0], outs["f32in"][0])
    assert np.array_equal(outs["rows"][0], outs["both"][0])
    hi, lo = bf16_split(outs["both"][0].reshape(M, F))
    assert np.array_equal(outs["both"][1][0], hi) and np.array_equal(outs["both"][1][1], lo)
    assert np.array_equal(outs["planes"][1], outs["both"][1])


@pytest.mark.parametrize("geom", [g for g in SPLIT_GEOMS if g[8]] + [(3, 64, 26, 26, 32, 1, 0, 1, True)])
def test_conv_forward_presplit_residual_from_planes(cuda_backend, oracle, geom):
    """tnn_conv2d_forward_split_residual with the shortcut operand as pre-split planes: the operand the kernel adds is
    hi + lo, so the result is BIT-IDENTICAL to the fp32-residual form fed hi + lo (both outputs), and within the 1e-4 gate
    of the oracle computed with the ORIGINAL fp32 operand (the planes hold it to 2^-18 relative)."""
    cu = cuda_backend
    B, C, H, W, F, k, p, s, _ = geom
    rng = np.random.default_rng(sum(geom[:8]) + 23)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    OH, OW = want.shape[2], want.shape[3]
    M = B * OH * OW
    res = rng.uniform(-1, 1, (B, F, OH, OW)).astype(np.float32)
    want = want + res
    dx, dw, db, ds = Dev(cu, to_nhwc(x)), Dev(cu, w), Dev(cu, b), Dev(cu, sc)
    dres = Dev(cu, to_nhwc(res))
    dxs, drs, drr = Dev(cu, n=x.size), Dev(cu, n=M * F), Dev(cu, n=M * F)
    n_in = B * H * W * C
    cu.nhwcToSplit(B * H * W, C, dx.ptr, C, dxs.ptr, C, n_in)
    cu.nhwcToSplit(M, F, dres.ptr, F, drs.ptr, F, M * F)
    cu.splitToNhwc(M, F, drs.ptr, F, M * F, drr.ptr, F)       # hi + lo as fp32 rows
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    act = oracle.ACT["LEAKY"]
    outs = {}
    cu.set_math_mode(BF3)
    try:
        for name, (rp, lay, plane) in {"planes": (drs, SPLIT, M * F), "rows": (drr, NHWC, 0)}.items():
            do, dsp = Dev(cu, n=M * F), Dev(cu, n=M * F)
            cu.conv2dForwardSplitResidual(desc, dxs.ptr, 0, C, SPLIT, n_in, dw.ptr, ds.ptr, db.ptr, act, rp.ptr, 0, F, lay,
                                          plane, do.ptr, 0, F, NHWC, dsp.ptr, 0, F, M * F)
            assert "k_conv_nhwc" in cu.lastConvKernel()
            outs[name] = (do.read((B, OH, OW, F)), dsp.read().view(np.uint16).reshape(2, M, F))
    finally:
        cu.set_math_mode(X3)
    assert np.array_equal(outs["planes"][0], outs["rows"][0])
    assert np.array_equal(outs["planes"][1], outs["rows"][1])
    got = from_nhwc(outs["planes"][0], F)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < TOL_3XTF32 and e2 < TOL_3XTF32, (e1, e2)


# (B, C, H, W, F, k, pad, stride)
IMAGE_GEOMS = [
    (2, 3, 20, 24, 32, 3, 1, 1),     # the YOLOv3 / CIFAR image layer shape: K = 27
    (3, 3, 33, 36, 32, 3, 1, 2),     # stride 2, ragged patches in both directions
    (2, 1, 28, 28, 16, 5, 2, 1),     # one plane, 5x5: K = 25, 16 filters
    (1, 3, 40, 40, 48, 3, 0, 1),     # no padding, 48 filters (two 32-filter chunks, the second half empty)
    (40, 3, 32, 32, 32, 3, 1, 1),    # 320 tiles: more tiles than SMs, every accumulator slot and stage reused
    (2, 2, 9, 12, 24, 4, 1, 1),      # K = 32 exactly, even kernel
    (5, 3, 416, 416, 32, 3, 1, 1),   # the YOLOv3 image layer itself (5 images: 6760 patches)
]


@pytest.mark.parametrize("mode,tol", [(BF3, TOL_3XTF32), (TF32, TOL_3XTF32)])
@pytest.mark.parametrize("geom", IMAGE_GEOMS)
def test_image_layer_on_tensor_cores(cuda_backend, oracle, mode, tol, geom):
    """nhwc::k_conv_image: planar image in, BF16x3 on the tensor cores (also in TF32 mode), fp32 rows and pre-split planes
    out — against the oracle within the 1e-4 gate; the planes are the split of the fp32 rows."""
    cu = cuda_backend
    B, C, H, W, F, k, p, s = geom
    rng = np.random.default_rng(sum(geom) + 1)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    OH, OW = want.shape[2], want.shape[3]
    M = B * OH * OW
    dx, dw, db, ds = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, sc)
    do, dsp = Dev(cu, n=M * F), Dev(cu, n=M * F)
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    cu.set_math_mode(mode)
    try:
        cu.conv2dForwardSplit(desc, dx.ptr, 0, 0, NCHW, 0, dw.ptr, ds.ptr, db.ptr, oracle.ACT["LEAKY"], None, 0, 0,
                              do.ptr, 0, F, NHWC, dsp.ptr if F % 8 == 0 else None, 0, F, M * F)
        assert "k_conv_image" in cu.lastConvKernel()
    finally:
        cu.set_math_mode(X3)
    rows = do.read((B, OH, OW, F))
    got = from_nhwc(rows, F)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)
    hi, lo = bf16_split(rows.reshape(M, F))
    planes = dsp.read().view(np.uint16).reshape(2, M, F)
    assert np.array_equal(planes[0], hi) and np.array_equal(planes[1], lo)
