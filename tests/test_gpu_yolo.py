"""End-to-end YOLOv3 graph on the GPU backend vs the CPU oracle: per-layer outputs and post-NMS boxes
(BASELINE.json: same count, same class, IoU >= 0.99)."""
import numpy as np
import pytest

from util import TOL_3XTF32, rel_max

pytestmark = pytest.mark.gpu


def run_net(cu, size, batch, fused, params, x):
    from tensorium_b200.nnet import build_yolov3
    net = build_yolov3(batch, size, params=params, fused=fused)
    cu.writeBuffer(net.input.devData, x.nbytes, x)
    net.predict()
    return net


@pytest.mark.parametrize("fused", [True, False])
def test_yolov3_small_against_oracle(cuda_backend, oracle, fused):
    from oracle import ref_net
    from tensorium_b200.nnet import box_iou, doNMSSort, init_yolov3_params, yolov3_spec
    cu = cuda_backend
    size, batch = 96, 2
    spec = yolov3_spec()
    params = init_yolov3_params(spec, seed=5, objectness_bias=-1.5)
    x = np.random.default_rng(6).random((batch, 3, size, size), dtype=np.float32)
    probe = {i: None for i in range(len(spec))}
    oracle.set_mode(oracle.MODE_FMA)
    try:
        want = ref_net.forward(spec, params, x, collect=probe)
    finally:
        oracle.set_mode(oracle.MODE_PAS)
    net = run_net(cu, size, batch, fused, params, x)
    # (1) per-layer parity: every conv layer is fed the ORACLE's input of that layer and must reproduce the
    #     oracle's output within the north-star tolerance (no accumulated drift in this comparison)
    from tensorium_b200.layers import TConvolutionalLayer, TNNetState
    from tensorium_b200.tensor import TSingleTensor
    worst = 0.0
    for i, l in enumerate(net.layers):
        if not isinstance(l, TConvolutionalLayer):
            continue
        src = x if i == 0 else probe[i - 1]
        t = TSingleTensor(src.shape, batch, src)
        st = TNNetState(input=t, workspace=net.workspace)
        l.forwardGPU(st)
        e = rel_max(l.output.pullFromDevice(), probe[i])
        worst = max(worst, e)
        assert e < TOL_3XTF32, (i, e)
    print("worst per-layer rel err", worst)
    # (2) end to end: 75 convolutions of drift stay well inside the detection tolerance
    net.predict()
    for i in (4, 36, 61, 79, 98, 105):
        got = net.layers[i].output.pullFromDevice()
        assert got.shape == probe[i].shape and rel_max(got, probe[i]) < 2e-3, (i, rel_max(got, probe[i]))
    outs = [l.output.pullFromDevice() for l in net.yolo_layers()]
    for g, r in zip(outs, want):
        assert rel_max(g, r) < 2e-3
    # detections + NMS on both sides (getDetections nyololayer.pas:399-441, doNMSSort ntypes.pas:1612-1646).
    # The threshold is placed in a gap of the oracle's objectness values near their 90% quantile so that a few dozen
    # boxes pass and none sits within the arithmetic tolerance of the threshold.
    objs = np.sort(np.concatenate([w.reshape(batch, 3, 85, -1)[:, :, 4, :].ravel() for w in want]))
    idx = int(0.9 * objs.size)
    while objs[idx] - objs[idx - 1] < 1e-3:
        idx += 1
    thresh = float((objs[idx] + objs[idx - 1]) / 2)
    total = 0
    for b in range(batch):
        db, do, dp = net.Detections(thresh, 0.45, b, host_outputs=outs)
        rb, ro, rp = net.Detections(thresh, 0.45, b, host_outputs=want)
        assert len(db) == len(rb)
        total += len(db)
        for i in range(len(db)):
            assert np.array_equal(dp[i] > 0, rp[i] > 0)  # same surviving classes
            assert box_iou(db[i], rb[i]) >= 0.99
    assert total > 0  # the comparison is not vacuous


def test_shortcut_fusion_is_bit_identical(cuda_backend):
    """tnn_conv2d_forward_residual: conv epilogue + the following shortcut's add == the two separate layers."""
    from tensorium_b200.nnet import build_yolov3, init_yolov3_params, yolov3_spec
    cu = cuda_backend
    size, batch = 64, 2
    params = init_yolov3_params(yolov3_spec(), seed=8)
    x = np.random.default_rng(2).random((batch, 3, size, size), dtype=np.float32)
    outs = []
    for fuse in (False, True):
        net = build_yolov3(batch, size, params=params, fuse_shortcuts=fuse)
        if fuse:
            assert sum(getattr(l, "_fused_into_conv", False) for l in net.layers) == 23
        cu.writeBuffer(net.input.devData, x.nbytes, x)
        net.predict()
        outs.append([net.layers[i].output.pullFromDevice().copy() for i in (4, 36, 61, 74)] +
                    [l.output.pullFromDevice().copy() for l in net.yolo_layers()])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_cfg_and_weights_file_build_the_same_net(cuda_backend, tmp_path):
    """yolov3.cfg text + a .weights file (tensorium_b200/darknet.py; nparser.pas:1140-1182, 1275-1330) give the net that
    build_yolov3 assembles from the same parameters: every head output bit-identical."""
    from tensorium_b200 import darknet as dn
    from tensorium_b200.nnet import build_yolov3, init_yolov3_params, yolov3_spec
    cu = cuda_backend
    size, batch = 96, 2
    spec = yolov3_spec()
    params = init_yolov3_params(spec, seed=8, objectness_bias=-1.5)
    path = tmp_path / "yolov3.weights"
    path.write_bytes(dn.write_weights(params, spec, 3, size, size, seen=640))
    x = np.random.default_rng(9).random((batch, 3, size, size), dtype=np.float32)
    a = build_yolov3(batch, size, params=params)
    b = dn.load_cfg(dn.yolov3_cfg(size=size, batch=batch), weights=str(path), fuse_bn=True)
    assert b.seen == 640 and len(b.layers) == 107
    for net in (a, b):
        cu.writeBuffer(net.input.devData, x.nbytes, x)
        net.predict()
    heads = [i for i, it in enumerate(spec) if it[0] == "yolo"]
    for i in heads:
        assert np.array_equal(a.layers[i].output.pullFromDevice(), b.layers[i].output.pullFromDevice())
    # and the parameters read back from the built net reproduce the file (fused batch norm excluded: build unfused)
    c = dn.load_cfg(dn.yolov3_cfg(size=size, batch=batch), weights=str(path), fuse_bn=False)
    _, spec_c = dn.cfg_to_spec(dn.parse_cfg(dn.yolov3_cfg(size=size, batch=batch)))
    assert dn.write_weights(dn.net_params(c, spec_c), spec_c, 3, size, size, seen=640) == path.read_bytes()


def test_cfg_built_training_net_steps(cuda_backend):
    """A [convolutional]/[maxpool]/[dropout]/[connected]/[softmax] cfg builds a trainable net whose loss goes down and
    whose saved weights reload to the same predictions."""
    from tensorium_b200 import darknet as dn
    cfg = ("[net]\nbatch=16\nwidth=16\nheight=16\nchannels=3\nlearning_rate=0.02\nmomentum=0.9\ndecay=0.0001\n"
           "[convolutional]\nbatch_normalize=1\nfilters=16\nsize=3\nstride=1\npad=1\nactivation=relu\n"
           "[maxpool]\nsize=2\nstride=2\npadding=0\n"
           "[convolutional]\nbatch_normalize=1\nfilters=32\nsize=3\nstride=1\npad=1\nactivation=relu\n"
           "[maxpool]\nsize=2\nstride=2\npadding=0\n[dropout]\nprobability=0.1\n"
           "[connected]\noutput=10\nactivation=linear\n[softmax]\n")
    net = dn.load_cfg(cfg, train=True)
    assert (net.learningRate, net.batch) == (0.02, 16)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (16, 3, 16, 16)).astype(np.float32)
    t = np.zeros((16, 10), np.float32)
    t[np.arange(16), rng.integers(0, 10, 16)] = 1
    costs = []
    for _ in range(25):
        net.trainBatch(x, t)
        costs.append(net.cost())
    assert costs[-1] < 0.3 * costs[0], costs
    _, spec = dn.cfg_to_spec(dn.parse_cfg(cfg))
    blob = dn.write_weights(dn.net_params(net, spec), spec, 3, 16, 16, seen=25 * 16)
    again, hdr = dn.read_weights(blob, spec, 3, 16, 16)
    assert hdr["seen"] == 400
    for p, l in zip(again, net.layers):
        if p is not None:
            assert np.array_equal(p["w"].reshape(-1), l.weights.pullFromDevice().reshape(-1))
