"""Darknet .cfg / .weights formats (host-only part of tensorium_b200/darknet.py; reference: source/nconfig.pas:18-56,
source/nparser.pas:154-268, 520-561, 1140-1182, 1275-1330)."""
import struct

import numpy as np
import pytest

from tensorium_b200 import darknet as dn
from tensorium_b200.nnet import YOLOV3_ANCHORS, init_yolov3_params, yolov3_spec

CFG = """
# a comment
[net]
batch=4
width=32
height=32
channels=3
learning_rate=0.01
momentum = 0.9
decay=0.0005

[convolutional]
batch_normalize=1
filters=8
size=3
stride=1
pad=1
activation=relu
; another comment style
[maxpool]
size=2
stride=2
padding=0

[convolutional]
filters=16
size=1
activation=linear

[dropout]
probability=.25

[connected]
output=10
activation=linear

[softmax]
"""


def test_parse_cfg_and_reference_defaults():
    net, spec = dn.cfg_to_spec(dn.parse_cfg(CFG))
    assert net["batch"] == 4 and net["width"] == 32 and abs(net["learning_rate"] - 0.01) < 1e-12 and net["subdivisions"] == 1
    assert spec == [("conv", 8, 3, 1, 1, "relu", 1), ("maxpool", 2, 2, 0), ("conv", 16, 1, 1, 0, "linear", 0),
                    ("dropout", 0.25), ("connected", 10, "linear"), ("softmax",)]
    # nparser.pas:190 / :256-257: activation defaults to logistic, max-pool size to stride and padding to size-1
    _, spec = dn.cfg_to_spec(dn.parse_cfg("[net]\nwidth=8\nheight=8\nchannels=1\n[convolutional]\nfilters=2\nsize=3\n[maxpool]\nstride=2\n"))
    assert spec == [("conv", 2, 3, 1, 0, "logistic", 0), ("maxpool", 2, 2, 1)]
    assert dn.infer_shapes(spec, 1, 8, 8) == [(2, 6, 6), (2, 3, 3)]
    with pytest.raises(ValueError):
        dn.cfg_to_spec(dn.parse_cfg("[net]\n[lstm]\noutput=3\n"))


def test_yolov3_cfg_round_trip_and_known_file_size():
    text = dn.yolov3_cfg(size=416, batch=64)
    net, spec = dn.cfg_to_spec(dn.parse_cfg(text))
    ref = yolov3_spec()
    assert len(spec) == 107 and (net["width"], net["height"], net["channels"], net["batch"]) == (416, 416, 3, 64)
    for a, b in zip(spec, ref):
        if b[0] == "conv":
            assert a == b + (b[2] // 2,)
        elif b[0] == "yolo":
            assert a == ("yolo", b[1], YOLOV3_ANCHORS, 80)
        else:
            assert a == b
    shapes = dn.infer_shapes(spec, 3, 416, 416)
    assert [shapes[i] for i, it in enumerate(spec) if it[0] == "yolo"] == [(255, 13, 13), (255, 26, 26), (255, 52, 52)]
    # the published yolov3.weights is 248 007 048 bytes = 20-byte header + 62 001 757 floats
    n = sum(cnt for _, fields in dn.param_layout(spec, 3, 416, 416) for _, cnt in fields)
    assert n == 62001757 and 20 + 4 * n == 248007048


def test_weights_round_trip_headers_and_truncation():
    net, spec = dn.cfg_to_spec(dn.parse_cfg(CFG))
    c, h, w = net["channels"], net["height"], net["width"]
    rng = np.random.default_rng(0)
    params = [None] * len(spec)
    for i, fields in dn.param_layout(spec, c, h, w):
        params[i] = {name: rng.standard_normal(n).astype(np.float32) for name, n in fields}
    assert [i for i, p in enumerate(params) if p] == [0, 2, 4]
    assert params[4]["w"].size == 10 * 16 * 16 * 16 and set(params[0]) == {"b", "scales", "rm", "rv", "w"} and set(params[2]) == {"b", "w"}
    blob = dn.write_weights(params, spec, c, h, w, seen=12345678901)
    assert struct.unpack("<3i", blob[:12]) == (0, 2, 5) and struct.unpack("<Q", blob[12:20])[0] == 12345678901
    # file order of a batch-normalised convolution: biases, scales, rolling mean, rolling variance, weights
    first = np.frombuffer(blob[20:20 + 4 * 8 * 4], "<f4").reshape(4, 8)
    assert all(np.array_equal(first[j], params[0][k]) for j, k in enumerate(("b", "scales", "rm", "rv")))
    back, hdr = dn.read_weights(blob, spec, c, h, w)
    assert hdr["seen"] == 12345678901 and hdr["major"] == 0 and hdr["minor"] == 2
    for p, q in zip(params, back):
        assert (p is None) == (q is None)
        if p:
            assert all(np.array_equal(p[k].reshape(-1), q[k].reshape(-1)) for k in p)
    assert back[0]["w"].shape == (8, 27) and back[4]["w"].shape == (10, 4096)
    # pre-0.2 files carry a 32-bit `seen` (nparser.pas:1294-1307)
    old = dn.write_weights(params, spec, c, h, w, seen=77, version=(0, 1, 0))
    assert len(old) == len(blob) - 4
    back2, hdr2 = dn.read_weights(old, spec, c, h, w)
    assert hdr2["seen"] == 77 and np.array_equal(back2[4]["w"], back[4]["w"])
    with pytest.raises(EOFError):
        dn.read_weights(blob[:-8], spec, c, h, w)


def test_yolov3_weights_file_round_trip_small_input():
    spec = yolov3_spec()
    params = init_yolov3_params(spec, seed=3)
    blob = dn.write_weights(params, spec, 3, 64, 64)
    assert len(blob) == 248007048  # parameter count does not depend on the input size
    back, _ = dn.read_weights(blob, spec, 3, 64, 64)
    for i in [j for j, it in enumerate(spec) if it[0] == "conv"][::9]:
        assert all(np.array_equal(params[i][k].reshape(-1), back[i][k].reshape(-1)) for k in params[i])


def test_transposed_connected_weights_round_trip():
    """Files whose header carries major or minor > 1000 store connected weights transposed (nparser.pas:1140-1182):
    write_weights honours the flag like read_weights does."""
    from tensorium_b200 import darknet as dn
    spec = [("conv", 4, 3, 1, 1, "relu", 1), ("connected", 5, "linear"), ("softmax",)]
    rng = np.random.default_rng(1)
    params = [dict(w=rng.standard_normal((4, 27)).astype(np.float32), b=rng.standard_normal(4).astype(np.float32),
                   scales=rng.standard_normal(4).astype(np.float32), rm=rng.standard_normal(4).astype(np.float32),
                   rv=rng.uniform(0.5, 1.5, 4).astype(np.float32)),
              dict(w=rng.standard_normal((5, 4 * 6 * 6)).astype(np.float32), b=rng.standard_normal(5).astype(np.float32)),
              None]
    plain = dn.write_weights(params, spec, 3, 6, 6, seen=7, version=(0, 2, 5))
    transposed = dn.write_weights(params, spec, 3, 6, 6, seen=7, version=(0, 1001, 0))
    assert len(plain) == len(transposed) and plain[20:] != transposed[20:]
    for blob in (plain, transposed):
        back, hdr = dn.read_weights(blob, spec, 3, 6, 6)
        assert hdr["seen"] == 7
        assert np.array_equal(back[1]["w"], params[1]["w"]) and np.array_equal(back[0]["w"].reshape(4, 27), params[0]["w"])


@pytest.mark.parametrize("bad", ["[net]\nwidth=8\nheight=8\nchannels=3\npolicy=steps\nsteps=100\nscales=.1\n[softmax]\n",
                                 "[net]\nwidth=8\nheight=8\nchannels=3\n[yolo]\nmask=0\nanchors=1,2\nclasses=1\nscale_x_y=1.05\n",
                                 "[net]\nwidth=8\nheight=8\nchannels=3\n[upsample]\nstride=2\nscale=2\n",
                                 "[net]\nwidth=8\nheight=8\nchannels=3\n[convolutional]\nfilters=3\n[shortcut]\nfrom=-1\nactivation=leaky\n"])
def test_cfg_keys_that_change_the_numerics_are_refused(bad):
    from tensorium_b200 import darknet as dn
    with pytest.raises(ValueError):
        dn.cfg_to_spec(dn.parse_cfg(bad))


def test_training_a_graph_net_is_refused_up_front():
    from tensorium_b200 import darknet as dn
    spec = [("conv", 8, 3, 1, 1, "leaky", 1), ("conv", 8, 3, 1, 1, "leaky", 1), ("shortcut", -2)]
    with pytest.raises(ValueError, match="sequential nets only"):
        dn.build_from_spec(spec, 2, 3, 8, 8, train=True)
