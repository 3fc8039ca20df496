"""Host-side graph plumbing of the YOLOv3 inference graph (tensorium_b200/nnet.py: setChannelsLast, fuseShortcuts,
setSplitActivations(planes_only), fuseYoloHeads, elideRoutes) checked WITHOUT a GPU: a recording stand-in for the
backend hands out fake device addresses and logs every operator call of one forward pass; the assertions are about
which entry point each layer calls and with which layouts, pitches and offsets (the numbers themselves are the GPU
tests' business, tests/test_gpu_yolo.py)."""
import numpy as np
import pytest

from tensorium_b200 import nncuda as nn
from tensorium_b200 import tensor


class RecordingBackend:
    """Every TNNCuda method becomes a logger; device buffers are fake, 256-byte aligned addresses."""
    deviceIndex = 0
    LAYOUT_NCHW, LAYOUT_NHWC, LAYOUT_NHWC_SPLIT = 0, 1, 2

    def __init__(self):
        self.calls, self._next = [], 1 << 20

    def createDeviceBuffer(self, nbytes):
        p = self._next
        self._next += (int(nbytes) + 255) // 256 * 256
        return p

    def freeDeviceBuffer(self, p):
        pass

    def conv_desc(self, *a):
        return a

    def __getattr__(self, name):
        def op(*args, **kw):
            self.calls.append((name, args))
        return op


@pytest.fixture()
def backend(monkeypatch):
    b = RecordingBackend()
    monkeypatch.setattr(tensor, "_cuda", b)
    yield b


def build(planes=True, fuse_yolo=True, elide=True, size=64, batch=2):
    from tensorium_b200.nnet import build_yolov3, init_yolov3_params, yolov3_spec
    params = init_yolov3_params(yolov3_spec(), seed=1)
    return build_yolov3(batch, size, params=params, channels_last=True, fuse_shortcuts=True,
                        split_activations="planes" if planes else True, fuse_yolo=fuse_yolo, elide_routes=elide)


def test_planes_only_graph_calls(backend):
    net = build()
    backend.calls.clear()
    net.predict()
    names = [c[0] for c in backend.calls]
    # 75 convolutions: 72 through the split entry point (23 of them with a folded shortcut read from planes), 3 yolo heads
    conv = [c for c in backend.calls if c[0] == "conv2dForwardSplitResidual"]
    heads = [c for c in backend.calls if c[0] == "conv2dForwardYoloHead"]
    assert len(conv) == 72 and len(heads) == 3
    with_res = [c for c in conv if c[1][10] is not None]
    assert len(with_res) == 23 and all(c[1][13] == nn.TNNCuda.LAYOUT_NHWC_SPLIT for c in with_res)
    # every tensor between convolutions exists only as planes: no fp32 output anywhere but the image layer's none either
    assert all(c[1][15] is None for c in conv)            # `out` (fp32 rows) never given
    assert all(c[1][19] is not None for c in conv)        # `outSplit` always given
    # the first convolution reads the planar image, every other one planes
    assert conv[0][1][4] == nn.TNNCuda.LAYOUT_NCHW and all(c[1][4] == nn.TNNCuda.LAYOUT_NHWC_SPLIT for c in conv[1:])
    assert all(h[1][4] == nn.TNNCuda.LAYOUT_NHWC_SPLIT for h in heads)
    # no route copies, no shortcut adds, no yolo copy / activation passes
    for gone in ("rowsCopy", "rowsAdd", "addvv", "yoloActivate", "copy", "ActivateArray"):
        assert gone not in names, gone
    assert names.count("upSampleNhwc") == 4               # two upsample layers x two planes


def test_elided_routes_use_pitch_and_offset(backend):
    from tensorium_b200.nnet import TConcatLayer, TUpsampleLayer
    net = build()
    routes = [(i, l) for i, l in enumerate(net.layers) if isinstance(l, TConcatLayer)]
    assert [len(l.inputLayers) for _, l in routes] == [1, 2, 1, 2] and all(l._elided for _, l in routes)
    for i, l in routes:
        if len(l.inputLayers) == 1:
            assert l.output is net.layers[l.inputLayers[0]].output          # an alias
            continue
        cat = l.output
        rows = net.batch * l.h * l.w
        off = 0
        for li, c in zip(l.inputLayers, l._cs):
            ptr, pitch, plane = tensor.split_view(net.layers[li].output)
            assert (ptr, pitch, plane) == (cat.split.devData + 2 * off, l.outC, rows * l.outC)
            off += c
    backend.calls.clear()
    net.predict()
    conv = [c for c in backend.calls if c[0] == "conv2dForwardSplitResidual"]
    # the skip tensors (layers 36 and 61) are WRITTEN with the route's pitch by the convolution that folds their shortcut
    # and READ with it by the stride-2 convolution behind them
    for li, route_i in ((36, 98), (61, 86)):
        cat = net.layers[route_i]
        ptr, pitch, plane = tensor.split_view(net.layers[li].output)
        writers = [c for c in conv if c[1][19] == ptr]
        readers = [c for c in conv if c[1][1] == ptr]
        assert len(writers) == 1 and writers[0][1][21] == cat.outC and writers[0][1][22] == plane
        assert len(readers) == 1 and readers[0][1][3] == cat.outC and readers[0][1][5] == plane
    # the upsample layers write into their slice (column 0) of the route buffers
    ups = [c for c in backend.calls if c[0] == "upSampleNhwc"]
    for route_i in (86, 98):
        cat = net.layers[route_i]
        assert sum(1 for c in ups if c[1][8] == cat.output.split.devData and c[1][9] == cat.outC // 2) == 1   # hi plane


def test_route_elision_is_reversible_and_fp32_graph_keeps_copying(backend):
    net = build()
    net.setSplitActivations(False)                       # back to fp32 rows (what bench.py does for the tf32 modes)
    assert not any(getattr(l, "_elided", False) for l in net.layers)
    assert all(getattr(l.output, "split_at", None) is None for l in net.layers)
    backend.calls.clear()
    net.predict()
    names = [c[0] for c in backend.calls]
    assert names.count("rowsCopy") == 6 and "conv2dForwardSplitResidual" not in names
    assert names.count("conv2dForwardNhwcResidual") == 23 and names.count("conv2dForwardYoloHead") == 3


def test_fp32_shortcut_graph_reads_rows(backend):
    """setSplitActivations() without planes_only: shortcut operands and 1x1 inputs stay fp32 rows."""
    net = build(planes=False, fuse_yolo=False, elide=False)
    backend.calls.clear()
    net.predict()
    conv = [c for c in backend.calls if c[0] == "conv2dForwardSplitResidual"]
    with_res = [c for c in conv if c[1][10] is not None]
    assert len(with_res) == 23 and all(c[1][13] == nn.TNNCuda.LAYOUT_NHWC for c in with_res)
    names = [c[0] for c in backend.calls]
    assert names.count("yoloActivate") == 3 and names.count("rowsCopy") > 0


def test_deepcifar10_training_step_call_sequence(backend):
    """One fused training step of deepCIFAR10 (models.py; nmodels.pas:82-111) as the backend sees it: six convolutions
    with training batch norm forward, the mirrored backward (batch-norm backward, wgrad, dgrad except into the image),
    ONE multi-tensor SGD launch — and, data parallel, ONE all-reduce of the update bucket between the last backward
    call and the update (dataparallel.NativeBucket; nnet.pas:378-403)."""
    from tensorium_b200 import dataparallel, models
    net = models.deepCIFAR10(8, fused=True, seed=0, learningRate=0.01)
    # a 2-replica bucket without a process group: the native bucket only needs a communicator handle and a world size
    grads = net.gradients()
    bucket = dataparallel.NativeBucket(backend, [g.Size() for g in grads], comm=object(), world=2)
    for i, g in enumerate(grads):
        g.rebind(bucket.ptr(i))
        g.setCUDA()
    net.bucket, net.replicas = bucket, 2
    x = np.zeros((8, 3, 32, 32), np.float32)
    t = np.zeros((8, 10), np.float32)
    t[:, 0] = 1
    net.input.Data[...] = x
    net.input.setCPU()
    net.truth = tensor.TSingleTensor(net.layers[-1].output.FShape, groups=8)
    net.truth.Data[...] = t
    backend.calls.clear()
    net.Propagate()
    net.update()
    names = [c[0] for c in backend.calls]
    assert names.count("conv2dForward") == 6 and names.count("batchNormForwardTraining") == 6
    assert names.count("batchNormBackward") == 6 and names.count("conv2dBackwardFilter") == 6
    assert names.count("conv2dBackwardData") == 5          # nothing is back-propagated into the image
    assert names.count("sgdUpdateMulti") == 1 and names.count("sgdUpdate") == 0
    assert names.count("allreduceSum") == 1
    ar, sgd = names.index("allreduceSum"), names.index("sgdUpdateMulti")
    last_bwd = max(i for i, n in enumerate(names) if n.startswith("conv2dBackward") or n == "batchNormBackward")
    assert last_bwd < ar < sgd
    _, (comm, base, off, n) = backend.calls[ar]
    assert base == bucket.base and off == 0 and n == bucket.total
    # every replica keeps momentum / replicas in its buffer: the SUM of the next step restores momentum * summed update
    items, mom = backend.calls[sgd][1]
    assert mom == pytest.approx(net.momentum / 2) and len(items) == len(grads)
