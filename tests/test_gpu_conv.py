"""GPU parity of the convolution path: im2col/col2im index mapping (bit-exact), implicit-GEMM forward /
backward-filter / backward-data (tolerance of BASELINE.json's north_star), and the layer-level sequences."""
import json
import os

import numpy as np
import pytest

from util import TOL_3XTF32, TOL_TF32, Dev, rel_elem, rel_max

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP32, X3, TF32 = 0, 1, 2

GEOMS = [  # (B, C, H, W, F, k, pad, stride)
    (2, 3, 16, 16, 32, 3, 1, 1),      # YOLO/CIFAR first layer shape (K = 27)
    (2, 32, 16, 16, 64, 3, 1, 2),     # stride-2 downsampler
    (3, 64, 13, 13, 32, 1, 0, 1),     # 1x1 bottleneck on a 13x13 map
    (2, 32, 13, 13, 255, 1, 0, 1),    # YOLO head: 255 filters
    (1, 128, 26, 26, 256, 3, 1, 1),   # deep 3x3
    (2, 5, 9, 8, 6, 3, 1, 2),         # ragged everything
    (1, 16, 7, 7, 300, 3, 1, 1),      # more than one filter tile
    (4, 8, 5, 5, 16, 5, 2, 1),        # 5x5 kernel (LeNet-like)
]


def test_im2col_bit_exact(cuda_backend, oracle):
    cu = cuda_backend
    for cs in json.load(open(os.path.join(GOLD, "hand_im2col.json"))):
        im = np.array(cs["im"], dtype=np.float32)
        want = np.array(cs["col"], dtype=np.float32)
        dim, dcol = Dev(cu, im, pad_front=3), Dev(cu, n=want.size, pad_front=5)
        c, h, w = im.shape
        cu.im2col(c, h, w, cs["kh"], cs["kw"], cs["ph"], cs["pw"], cs["sy"], cs["sx"], cs["dy"], cs["dx"], dim._base, 3,
                  dcol._base, 5)
        assert np.array_equal(dcol.read(want.shape), want)
    rng = np.random.default_rng(0)
    for (c, h, w, k, p, s, d) in [(3, 17, 13, 3, 1, 1, 1), (4, 16, 16, 3, 1, 2, 1), (2, 9, 9, 3, 2, 1, 2), (5, 8, 6, 1, 0, 1, 1)]:
        x = rng.uniform(-1, 1, (c, h, w)).astype(np.float32)
        want = oracle.im2col(x, k, k, p, p, s, s, d, d)
        dim, dcol = Dev(cu, x), Dev(cu, n=want.size)
        cu.im2col(c, h, w, k, k, p, p, s, s, d, d, dim.ptr, 0, dcol.ptr, 0)
        assert np.array_equal(dcol.read(want.shape), want)  # pure gather: bytes must match


def test_col2im_bit_exact(cuda_backend, oracle):
    cu = cuda_backend
    rng = np.random.default_rng(1)
    for (c, h, w, k, p, s, d) in [(3, 17, 13, 3, 1, 1, 1), (4, 16, 16, 3, 1, 2, 1), (2, 9, 9, 3, 2, 1, 2), (5, 8, 6, 1, 0, 1, 1),
                                  (2, 7, 7, 2, 0, 2, 1)]:
        oh, ow = oracle.out_dim(h, p, k, s, d), oracle.out_dim(w, p, k, s, d)
        # (a) small integers: exact sums regardless of order
        col = rng.integers(-4, 5, (c * k * k, oh * ow)).astype(np.float32)
        im0 = rng.integers(-4, 5, (c, h, w)).astype(np.float32)
        want = oracle.col2im(col, c, h, w, k, k, p, p, s, s, d, d, im=im0.copy())
        dcol, dim = Dev(cu, col), Dev(cu, im0)
        cu.col2im(c, h, w, k, k, p, p, s, s, d, d, dcol.ptr, 0, dim.ptr, 0)
        assert np.array_equal(dim.read(im0.shape), want)
        # (b) real values: the gather adds in the CPU scatter's (kernel_row, kernel_col) order -> same bits
        col = rng.uniform(-1, 1, col.shape).astype(np.float32)
        want = oracle.col2im(col, c, h, w, k, k, p, p, s, s, d, d, im=im0.copy())
        dcol.write(col)
        dim.write(im0)
        cu.col2im(c, h, w, k, k, p, p, s, s, d, d, dcol.ptr, 0, dim.ptr, 0)
        assert np.array_equal(dim.read(im0.shape), want)


def conv_case(rng, B, C, H, W, F, k):
    x = rng.uniform(0, 1, (B, C, H, W)).astype(np.float32)
    s = np.sqrt(2.0 / (k * k * C))
    w = rng.uniform(-s, s, (F, C * k * k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, F).astype(np.float32)
    return x, w, b


@pytest.mark.parametrize("mode,tol", [(FP32, 2e-6), (X3, TOL_3XTF32), (TF32, TOL_TF32)])
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_forward(cuda_backend, oracle, mode, tol, geom):
    cu = cuda_backend
    B, C, H, W, F, k, p, s = geom
    rng = np.random.default_rng(sum(geom))
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    dx, dw, db = Dev(cu, x, pad_front=2), Dev(cu, w), Dev(cu, b)
    dout = Dev(cu, n=want.size, pad_front=1)
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    cu.set_math_mode(mode)
    cu.conv2dForward(desc, dx._base, 2, dw.ptr, None, db.ptr, oracle.ACT["LEAKY"], dout._base, 1)
    cu.set_math_mode(X3)
    got = dout.read(want.shape)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)


def test_conv_forward_fp32_bit_exact_with_fma_oracle(cuda_backend, oracle):
    cu = cuda_backend
    oracle.set_mode(oracle.MODE_FMA)
    try:
        B, C, H, W, F, k, p, s = 2, 5, 9, 8, 6, 3, 1, 2
        x, w, b = conv_case(np.random.default_rng(3), B, C, H, W, F, k)
        want = oracle.conv2d(x, w, p, s, s, 1, k, k)
        dx, dw, dout = Dev(cu, x), Dev(cu, w), Dev(cu, n=want.size)
        cu.set_math_mode(FP32)
        cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s), dx.ptr, 0, dw.ptr, None, None, 4, dout.ptr, 0)
        assert np.array_equal(dout.read(want.shape), want)
    finally:
        oracle.set_mode(oracle.MODE_PAS)
        cu.set_math_mode(X3)


@pytest.mark.parametrize("act", ["LINEAR", "LOGISTIC", "RELU"])
def test_conv_forward_scale_bias_activation_epilogue(cuda_backend, oracle, act):
    cu = cuda_backend
    B, C, H, W, F, k, p, s = 2, 16, 12, 12, 48, 3, 1, 1
    rng = np.random.default_rng(9)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    sc = rng.uniform(0.5, 1.5, F).astype(np.float32)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_scale(want, sc, B)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT[act])
    dx, dw, db, ds, dout = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, sc), Dev(cu, n=want.size)
    cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s), dx.ptr, 0, dw.ptr, ds.ptr, db.ptr, oracle.ACT[act],
                     dout.ptr, 0)
    assert rel_max(dout.read(want.shape), want) < TOL_3XTF32


TMA_GEOMS = [  # stride-1 geometries whose rows are 16-byte aligned: activations arrive by TMA (patch atoms)
    (2, 32, 32, 32, 32, 3, 1, 1),    # deepCIFAR10 c2: 4x... 32-wide rows, one row per atom
    (2, 64, 16, 16, 64, 3, 1, 1),    # 16-wide: 2x16 patches
    (3, 128, 8, 8, 128, 3, 1, 1),    # 8-wide: 4x8 patches, atoms straddle nothing; 3 images = 6 atoms (ragged tile)
    (1, 64, 52, 52, 32, 1, 0, 1),    # YOLO 1x1 bottleneck: flattened 2704-pixel rows, last atom ragged
    (2, 32, 20, 24, 48, 3, 1, 1),    # W=24: 8-wide patches, H=20 = 5 patch rows
    (1, 48, 12, 16, 40, 5, 2, 1),    # 5x5 kernel, C=48 (second channel block half empty)
    (2, 36, 8, 8, 16, 3, 1, 1),      # C=36: channel quads past the end are zero-filled
    (1, 16, 40, 36, 24, 3, 0, 1),    # no padding: output 38x34, patches overhang the right/bottom edge
    (2, 512, 4, 4, 64, 1, 0, 1),     # tiny maps, deep: flattened 16-pixel rows (atom half empty)
]


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32)])
@pytest.mark.parametrize("geom", TMA_GEOMS)
def test_conv_forward_tma_activations(cuda_backend, oracle, mode, tol, geom):
    cu = cuda_backend
    B, C, H, W, F, k, p, s = geom
    rng = np.random.default_rng(sum(geom) + 7)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    dx, dw, db, dout = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, n=want.size)
    cu.fill(want.size, dout.ptr, 0, 777.0, 1)  # rows the kernel must not skip
    cu.set_math_mode(mode)
    try:
        cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s), dx.ptr, 0, dw.ptr, None, db.ptr,
                         oracle.ACT["LEAKY"], dout.ptr, 0)
    finally:
        cu.set_math_mode(X3)
    got = dout.read(want.shape)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)


HALO_GEOMS = [  # k >= 2 on maps that fill the 16x8 patches: the halo-tile kernel (stride 1, and 2 via parity planes)
    (2, 32, 32, 32, 32, 3, 1, 1),    # deepCIFAR10 c2
    (2, 64, 16, 16, 64, 3, 1, 1),    # deepCIFAR10 c4: one patch row per image
    (1, 32, 48, 40, 64, 3, 1, 1),    # YOLO-like, 3 x 5 patches
    (2, 64, 64, 64, 128, 3, 1, 2),   # stride 2: four parity planes, 16-channel halo stages (SWIZZLE_64B weights)
    (2, 48, 64, 48, 40, 3, 1, 2),    # stride 2, C = 48 (three 16-channel stages), ragged filter tile
    (1, 32, 64, 64, 24, 5, 2, 2),    # 5x5 stride 2: taps shift by -1, 0, +1 inside the planes
    (1, 20, 32, 32, 16, 3, 1, 1),    # C = 20: the channel block is zero padded to 32
    (1, 16, 34, 34, 32, 3, 0, 1),    # no padding: output 32 x 32
    (1, 32, 32, 24, 160, 3, 1, 1),   # two filter tiles (128 + 32): the halo is staged once per filter tile
]


@pytest.mark.parametrize("mode,tol", [(X3, TOL_3XTF32), (TF32, TOL_TF32)])
@pytest.mark.parametrize("geom", HALO_GEOMS)
def test_conv_forward_halo_kernel(cuda_backend, oracle, mode, tol, geom):
    cu = cuda_backend
    B, C, H, W, F, k, p, s = geom
    rng = np.random.default_rng(sum(geom) + 11)
    x, w, b = conv_case(rng, B, C, H, W, F, k)
    want = oracle.conv2d(x, w, p, s, s, 1, k, k)
    oracle.forward_bias(want, b, B)
    oracle.activate_array(want, oracle.ACT["LEAKY"])
    dx, dw, db, dout = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, n=want.size)
    cu.fill(want.size, dout.ptr, 0, 777.0, 1)
    cu.set_math_mode(mode)
    try:
        cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s), dx.ptr, 0, dw.ptr, None, db.ptr,
                         oracle.ACT["LEAKY"], dout.ptr, 0)
        assert "halo" in cu.lastConvKernel()  # the geometry really exercises the halo kernel
    finally:
        cu.set_math_mode(X3)
    got = dout.read(want.shape)
    e1, e2 = rel_max(got, want), rel_elem(got, want, floor=1.0)
    assert e1 < tol and e2 < tol, (e1, e2)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_halo_index_exact(cuda_backend, stride):
    """Index-encoded input against one-hot filters on the halo kernel: every output is EXACTLY the input element the
    im2col mapping names (shifted descriptor views, parity planes, zero-filled borders)."""
    cu = cuda_backend
    B, C, H, W, k, p = 2, 32, 32, 16, 3, 1
    c_, h_, w_ = np.meshgrid(np.arange(C), np.arange(H), np.arange(W), indexing="ij")
    x = np.broadcast_to((c_ * 4096 + h_ * 64 + w_ + 1).astype(np.float32), (B, C, H, W)).copy()
    x[1] += 131072.0
    taps = [(c, kr, kc) for c in (0, 5, 31) for kr in range(k) for kc in range(k)]
    F = len(taps)
    w = np.zeros((F, C, k, k), np.float32)
    for f, (c, kr, kc) in enumerate(taps):
        w[f, c, kr, kc] = 1
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
    OH, OW = (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1
    want = np.stack([xp[:, c, kr:kr + stride * (OH - 1) + 1:stride, kc:kc + stride * (OW - 1) + 1:stride]
                     for (c, kr, kc) in taps], axis=1)
    dx, dw, dout = Dev(cu, x), Dev(cu, w), Dev(cu, n=want.size)
    cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, stride, stride), dx.ptr, 0, dw.ptr, None, None, 4,
                     dout.ptr, 0)
    assert "halo" in cu.lastConvKernel()
    assert np.array_equal(dout.read(want.shape), want)


def test_conv_tma_activations_index_exact(cuda_backend):
    """The same for the TMA-fed 1x1 path: out[f] must be EXACTLY channel c_f of the input (atoms in the right place,
    ragged last atom zero filled)."""
    cu = cuda_backend
    B, C, H, W = 2, 64, 10, 12   # 120 pixels per image: 3 full atoms + one of 24
    x = (np.arange(B * C * H * W) % 4093).astype(np.float32).reshape(B, C, H, W)
    chans = [0, 5, 31, 32, 63]
    w = np.zeros((len(chans), C, 1, 1), np.float32)
    for f, c in enumerate(chans):
        w[f, c] = 1
    want = x[:, chans]
    dx, dw, dout = Dev(cu, x), Dev(cu, w), Dev(cu, n=want.size)
    cu.conv2dForward(cu.conv_desc(B, C, H, W, len(chans), 1, 1, 0, 0, 1, 1), dx.ptr, 0, dw.ptr, None, None, 4, dout.ptr, 0)
    assert "TmaPatchLoader" in cu.lastConvKernel()
    assert np.array_equal(dout.read(want.shape), want)


BWD_GEOMS = GEOMS[:6] + [
    (8, 32, 32, 32, 64, 3, 1, 1),    # deepCIFAR10 c3 shape (P = 1024: tcgen05 wgrad with split-K, tcgen05 dgrad)
    (4, 64, 16, 16, 64, 3, 1, 1),    # deepCIFAR10 c4 shape
    (2, 128, 8, 8, 128, 3, 1, 1),    # deepCIFAR10 c6 shape (P = 64)
    (4, 16, 12, 12, 40, 1, 0, 1),    # 1x1, P = 144 (not a multiple of 32: CUDA-core wgrad)
    (2, 48, 16, 16, 160, 3, 1, 1),   # tile wgrad: two channel blocks (32 + 16), two filter tiles (128 + 32)
    (1, 16, 12, 16, 24, 3, 1, 1),    # tile wgrad: 2 x 16 pixel tiles, half-empty channel block
    (2, 32, 8, 8, 16, 1, 0, 1),      # tile wgrad of a 1x1 convolution (one tap, 4 x 8 tiles)
    (1, 24, 16, 32, 40, 2, 0, 1),    # 2x2 kernel, no padding: output 15 x 31 (not tileable: split-K GEMM wgrad)
]


@pytest.mark.parametrize("mode,tol", [(FP32, 5e-6), (X3, TOL_3XTF32)])
@pytest.mark.parametrize("geom", BWD_GEOMS)
def test_conv_backward(cuda_backend, oracle, geom, mode, tol):
    cu = cuda_backend
    B, C, H, W, F, k, p, s = geom
    rng = np.random.default_rng(sum(geom) + 1)
    x, w, _ = conv_case(rng, B, C, H, W, F, k)
    oh, ow = oracle.out_dim(H, p, k, s), oracle.out_dim(W, p, k, s)
    delta = rng.uniform(-1, 1, (B, F, oh, ow)).astype(np.float32)
    wu0 = rng.uniform(-1, 1, w.shape).astype(np.float32)
    dx0 = rng.uniform(-1, 1, x.shape).astype(np.float32)
    # oracle: nConvolutionLayer.pas:628-653
    col = oracle.im2col_batched(x, k, k, p, p, s, s)
    wu = wu0.copy()
    for b in range(B):
        oracle.sgemm(False, True, F, C * k * k, oh * ow, 1.0, delta[b], oh * ow, col[b], oh * ow, 1.0, wu, C * k * k)
    ws = np.zeros_like(col)
    oracle.sgemm_strided_batched(True, False, C * k * k, oh * ow, F, 1.0, w, C * k * k, 0, delta, oh * ow, F * oh * ow, 0.0,
                                 ws, oh * ow, ws[0].size, B)
    dxw = oracle.col2im_batched(ws, B, C, H, W, k, k, p, p, s, s, im=dx0.copy())
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    din, dw, dd, dwu, ddx = Dev(cu, x), Dev(cu, w), Dev(cu, delta), Dev(cu, wu0), Dev(cu, dx0)
    cu.set_math_mode(mode)
    try:
        cu.conv2dBackwardFilter(desc, din.ptr, 0, dd.ptr, 0, dwu.ptr)
        cu.conv2dBackwardData(desc, dw.ptr, dd.ptr, 0, ddx.ptr, 0)
    finally:
        cu.set_math_mode(X3)
    assert rel_max(dwu.read(w.shape), wu) < tol
    assert rel_max(ddx.read(x.shape), dxw) < tol


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("bn", [0, 1])
def test_layer_forward_backward_update_against_golden(cuda_backend, oracle, fused, bn):
    """TConvolutionalLayer.forwardGPU/backwardGPU/updateGPU (fused entry points and the reference's own call
    sequence) against the frozen fixtures of tests/golden/ (oracle, pure-Pascal mode)."""
    from tensorium_b200 import layers as L
    from tensorium_b200.tensor import TSingleTensor
    f = np.load(os.path.join(GOLD, "oracle_fixtures.npz"))
    p = f"pas_conv_bn{bn}_"
    lay = L.TConvolutionalLayer(3, 9, 8, 5, 6, 3, 2, 1, oracle.ACT["LEAKY"], bn, train=True, fused=fused)
    lay.weights.Data[...] = f[p + "w"]
    lay.biases.Data[...] = f[p + "bias"]
    if bn:
        lay.scales.Data[...] = f[p + "scales"]
        lay.rolling_mean.Data[...] = f[p + "rm"]
        lay.rolling_variance.Data[...] = f[p + "rv"]
    x = TSingleTensor(f[p + "x"].shape, 3, f[p + "x"])
    ws = TSingleTensor((lay.getWorkspaceSize(),), 1)
    st = L.TNNetState(input=x, workspace=ws, isTraining=False)
    lay.forwardGPU(st)
    assert rel_max(lay.output.pullFromDevice(), f[p + "y_inf"]) < TOL_3XTF32
    st.isTraining = True
    lay.forwardGPU(st)
    assert rel_max(lay.output.pullFromDevice(), f[p + "y_tr"]) < TOL_3XTF32
    lay.output.setCUDA()
    if bn:
        assert rel_max(lay.rolling_mean.pullFromDevice(), f[p + "rm_after"]) < 1e-5
        assert rel_max(lay.mean.pullFromDevice(), f[p + "mean"]) < 1e-5
        for t in (lay.rolling_mean, lay.mean):
            t.setCUDA()
    lay.delta.Data[...] = f[p + "g"]
    lay.delta.setCPU()
    st.delta = TSingleTensor(f[p + "x"].shape, 3)
    st.delta.pushToDevice()
    lay.backwardGPU(st)
    assert rel_max(st.delta.pullFromDevice(), f[p + "dx"]) < TOL_3XTF32
    assert rel_max(lay.weight_updates.pullFromDevice(), f[p + "dw"]) < TOL_3XTF32
    lay.weight_updates.setCUDA()
    if bn:
        assert rel_max(lay.scale_updates.pullFromDevice(), f[p + "dscale"]) < TOL_3XTF32
        lay.scale_updates.setCUDA()
    else:
        assert rel_max(lay.bias_updates.pullFromDevice(), f[p + "db"]) < 1e-5
        lay.bias_updates.setCUDA()
    lay.updateGPU(L.TUpdateArgs(batch=3, learningRate=0.001, momentum=0.9, decay=0.0001))
    assert rel_max(lay.weights.pullFromDevice(), f[p + "w_after"]) < 1e-6
    assert rel_max(lay.weight_updates.pullFromDevice(), f[p + "wu_after"]) < TOL_3XTF32


def test_conv_roundtrip_linearity_large(cuda_backend):
    """Size-independent property at a bench-sized layer: conv(a*x1 + x2) == a*conv(x1) + conv(x2)."""
    cu = cuda_backend
    B, C, H, W, F, k, p, s = 4, 64, 104, 104, 128, 3, 1, 1
    rng = np.random.default_rng(12)
    x1 = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, (F, C * k * k)).astype(np.float32)
    desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
    dw = Dev(cu, w)
    outs = []
    for x in (x1, x2, (np.float32(0.5) * x1 + x2).astype(np.float32)):
        dx, do = Dev(cu, x), Dev(cu, n=B * F * H * W)
        cu.conv2dForward(desc, dx.ptr, 0, dw.ptr, None, None, 4, do.ptr, 0)
        outs.append(do.read())
    assert rel_max(outs[2], 0.5 * outs[0].astype(np.float64) + outs[1]) < 1e-5


def test_static_weights_cache(cuda_backend, oracle):
    """tnn_set_weights_static: same results as the per-call preparation, and re-declaring after a change picks up the
    new contents (on each of the three tensor-core kernels)."""
    cu = cuda_backend
    rng = np.random.default_rng(77)
    for geom in [(2, 32, 32, 32, 64, 3, 1, 1), (2, 64, 26, 26, 48, 3, 1, 1), (2, 64, 16, 16, 32, 1, 0, 1)]:
        B, C, H, W, F, k, p, s = geom
        x, w, b = conv_case(rng, B, C, H, W, F, k)
        desc = cu.conv_desc(B, C, H, W, F, k, k, p, p, s, s)
        dx, dw, db, dout = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, n=B * F * H * W)

        def run():
            cu.conv2dForward(desc, dx.ptr, 0, dw.ptr, None, db.ptr, oracle.ACT["LEAKY"], dout.ptr, 0)
            return dout.read((B, F, H, W))
        base = run()
        cu.setWeightsStatic(dw.ptr, True)
        assert np.array_equal(run(), base) and np.array_equal(run(), base)  # first call prepares, second reuses
        w2 = (w * 0.5).astype(np.float32)
        dw.write(w2)
        cu.setWeightsStatic(dw.ptr, True)  # contents changed: declare again
        want = oracle.conv2d(x, w2, p, s, s, 1, k, k)
        oracle.forward_bias(want, b, B)
        oracle.activate_array(want, oracle.ACT["LEAKY"])
        assert rel_max(run(), want) < TOL_3XTF32
        cu.setWeightsStatic(dw.ptr, False)
        assert rel_max(run(), want) < TOL_3XTF32
