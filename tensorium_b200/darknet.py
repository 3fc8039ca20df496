"""Darknet on-disk formats either side of the convolution path:

  .cfg      source/nconfig.pas:18-56 (TCFGSection / TCFGList), source/nparser.pas:154-247 (parseConvolutional),
            :249-268 (parseMaxPool), :520-561 (parseYolo): `[section]` headers, `key=value` lines, `#`/`;` comments;
            defaults follow the reference (convolution: activation `logistic`, `pad=1` means padding = size div 2;
            max-pool: size = stride, padding = size-1).
  .weights  source/nparser.pas:1140-1182 (per-layer order) and :1275-1330 (header): int32 major, minor, revision, then
            `seen` as uint64 when major*10+minor >= 2 else uint32; per convolution biases[n], then (batch-normalised)
            scales[n], rolling_mean[n], rolling_variance[n], then weights[n*c*k*k]; per connected layer biases, weights.

Everything up to `build_from_spec` is host-only (numpy): a spec is a list of tuples in the vocabulary of
`nnet.yolov3_spec`, parameters are the dict lists the model builders take.  `build_from_spec` instantiates the layers on
the current `cuda` backend.
"""
from __future__ import annotations

import io
import struct

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------ .cfg
def parse_cfg(text):
    """-> [(section name, {key: value string})] in file order."""
    sections = []
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line[0] in "#;":
            continue
        if line[0] == "[":
            assert line[-1] == "]", f"malformed section header: {raw!r}"
            sections.append((line[1:-1].strip().lower(), {}))
            continue
        assert sections, f"option before the first section: {raw!r}"
        key, sep, val = line.partition("=")
        assert sep, f"malformed option line: {raw!r}"
        sections[-1][1][key.strip()] = val.strip()
    return sections


def _int(o, k, d):
    return int(o.get(k, d))


def _float(o, k, d):
    return float(o.get(k, d))


def _ints(s):
    return [int(x) for x in s.replace(" ", "").split(",") if x]


def cfg_to_spec(sections):
    """-> (net options dict, spec).  Spec entries:
    ("conv", filters, size, stride, batch_normalize, activation, padding) | ("maxpool", size, stride, padding) |
    ("connected", outputs, activation) | ("dropout", probability) | ("softmax",) | ("shortcut", from) |
    ("route", [layers]) | ("upsample", stride) | ("yolo", mask, anchors, classes)"""
    assert sections and sections[0][0] in ("net", "network"), "the first section must be [net]"
    o = sections[0][1]
    net = dict(batch=_int(o, "batch", 1), subdivisions=_int(o, "subdivisions", 1), width=_int(o, "width", 0),
               height=_int(o, "height", 0), channels=_int(o, "channels", 0), momentum=_float(o, "momentum", 0.9),
               decay=_float(o, "decay", 0.0001), learning_rate=_float(o, "learning_rate", 0.001))
    # keys that change the numerics and that this mirror does not implement are refused, not ignored
    policy = o.get("policy", "constant").strip().lower()
    if policy != "constant":
        raise ValueError(f"[net] policy={policy}: only the constant learning-rate policy is mirrored (steps / scales / "
                         "burn_in schedules are outside the path)")
    spec = []
    for name, o in sections[1:]:
        if name == "convolutional":
            size = _int(o, "size", 1)
            padding = size // 2 if _int(o, "pad", 0) else _int(o, "padding", 0)
            assert _int(o, "groups", 1) == 1, "grouped convolutions are outside the path"
            assert size == 1 or _int(o, "dilation", 1) == 1, "dilated convolutions are not expressible in a spec tuple"
            spec.append(("conv", _int(o, "filters", 1), size, _int(o, "stride", 1), 1 if _int(o, "batch_normalize", 0) else 0,
                         o.get("activation", "logistic"), padding))
        elif name == "maxpool":
            stride = _int(o, "stride", 1)
            size = _int(o, "size", stride)
            spec.append(("maxpool", size, stride, _int(o, "padding", size - 1)))
        elif name == "connected":
            spec.append(("connected", _int(o, "output", 1), o.get("activation", "logistic")))
        elif name == "dropout":
            spec.append(("dropout", _float(o, "probability", 0.5)))
        elif name == "softmax":
            spec.append(("softmax",))
        elif name == "shortcut":
            if o.get("activation", "linear").strip().lower() != "linear":
                raise ValueError("[shortcut] activation other than linear is not mirrored")
            spec.append(("shortcut", _ints(o["from"])[0]))
        elif name == "route":
            spec.append(("route", _ints(o["layers"])))
        elif name == "upsample":
            if _float(o, "scale", 1.0) != 1.0:
                raise ValueError("[upsample] scale other than 1 is not mirrored")
            spec.append(("upsample", _int(o, "stride", 2)))
        elif name == "yolo":
            if _float(o, "scale_x_y", 1.0) != 1.0:
                raise ValueError("[yolo] scale_x_y other than 1 (YOLOv4-style heads) is not mirrored")
            spec.append(("yolo", _ints(o.get("mask", "")), _ints(o["anchors"]), _int(o, "classes", 20)))
        else:
            raise ValueError(f"[{name}] is outside the convolution path")
    return net, spec


def spec_to_cfg(net, spec):
    """The inverse of cfg_to_spec (what TCFGList.saveToFile would write for these sections)."""
    out = ["[net]"] + [f"{k}={net[k]}" for k in ("batch", "subdivisions", "width", "height", "channels", "momentum",
                                                "decay", "learning_rate") if k in net]
    for item in spec:
        kind = item[0]
        out.append("")
        if kind == "conv":
            _, f, k, s, bn, act = item[:6]
            pad = item[6] if len(item) > 6 else k // 2
            out += ["[convolutional]"] + (["batch_normalize=1"] if bn else []) + [f"filters={f}", f"size={k}", f"stride={s}"]
            out += ["pad=1"] if pad == k // 2 else [f"padding={pad}"]
            out += [f"activation={act}"]
        elif kind == "maxpool":
            out += ["[maxpool]", f"size={item[1]}", f"stride={item[2]}", f"padding={item[3]}"]
        elif kind == "connected":
            out += ["[connected]", f"output={item[1]}", f"activation={item[2]}"]
        elif kind == "dropout":
            out += ["[dropout]", f"probability={item[1]}"]
        elif kind == "softmax":
            out += ["[softmax]"]
        elif kind == "shortcut":
            out += ["[shortcut]", f"from={item[1]}", "activation=linear"]
        elif kind == "route":
            out += ["[route]", "layers=" + ",".join(str(j) for j in item[1])]
        elif kind == "upsample":
            out += ["[upsample]", f"stride={item[1]}"]
        elif kind == "yolo":
            anchors = item[2] if len(item) > 2 else None
            classes = item[3] if len(item) > 3 else 80
            if anchors is None:
                from .nnet import YOLOV3_ANCHORS
                anchors = YOLOV3_ANCHORS
            out += ["[yolo]", "mask=" + ",".join(str(m) for m in item[1]), "anchors=" + ",".join(str(a) for a in anchors),
                    f"classes={classes}", f"num={len(anchors) // 2}"]
    return "\n".join(out) + "\n"


def yolov3_cfg(size=416, batch=64):
    """The standard yolov3.cfg (not in the reference tree, MSCOCOYolo.pas:31) regenerated from nnet.yolov3_spec."""
    from .nnet import yolov3_spec
    return spec_to_cfg(dict(batch=batch, subdivisions=1, width=size, height=size, channels=3, momentum=0.9, decay=0.0005,
                            learning_rate=0.001), yolov3_spec())


# ------------------------------------------------------------------ shapes and parameters
def infer_shapes(spec, c, h, w):
    """Output (c, h, w) of every layer; connected / softmax / dropout outputs are (n, 1, 1)."""
    shapes = []
    for i, item in enumerate(spec):
        kind = item[0]
        if kind == "conv":
            _, f, k, s = item[:4]
            pad = item[6] if len(item) > 6 else k // 2
            c, h, w = f, (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        elif kind == "maxpool":
            _, k, s, pad = item
            h, w = (h + pad - k) // s + 1, (w + pad - k) // s + 1
        elif kind == "connected":
            c, h, w = item[1], 1, 1
        elif kind == "route":
            src = [shapes[(i + j) if j < 0 else j] for j in item[1]]
            c, h, w = sum(s_[0] for s_ in src), src[0][1], src[0][2]
        elif kind == "shortcut":
            c, h, w = shapes[i + item[1] if item[1] < 0 else item[1]]
        elif kind == "upsample":
            h, w = h * item[1], w * item[1]
        shapes.append((c, h, w))
    return shapes


def param_layout(spec, c, h, w):
    """[(layer index, [(field, count)...])] in the order loadWeights walks the file (nparser.pas:1140-1182, 1307-1330)."""
    shapes = infer_shapes(spec, c, h, w)
    out = []
    for i, item in enumerate(spec):
        ci, hi, wi = shapes[i - 1] if i else (c, h, w)
        if item[0] == "conv":
            _, f, k, s, bn = item[:5]
            fields = [("b", f)] + ([("scales", f), ("rm", f), ("rv", f)] if bn else []) + [("w", f * ci * k * k)]
            out.append((i, fields))
        elif item[0] == "connected":
            out.append((i, [("b", item[1]), ("w", item[1] * ci * hi * wi)]))
    return out


def read_weights(data, spec, c, h, w):
    """-> (params list aligned with spec (None for parameter-free layers), header dict)."""
    fp = io.BytesIO(data) if isinstance(data, (bytes, bytearray)) else open(data, "rb")
    try:
        major, minor, revision = struct.unpack("<3i", fp.read(12))
        seen = struct.unpack("<Q", fp.read(8))[0] if major * 10 + minor >= 2 else struct.unpack("<I", fp.read(4))[0]
        transpose = major > 1000 or minor > 1000
        params = [None] * len(spec)
        for i, fields in param_layout(spec, c, h, w):
            p = {}
            for name, n in fields:
                buf = fp.read(4 * n)
                if len(buf) != 4 * n:
                    raise EOFError(f"unexpected end of the weights file in layer {i} ({name}: wanted {n} floats)")
                p[name] = np.frombuffer(buf, dtype="<f4").astype(f32)
            rows = p["b"].size
            p["w"] = p["w"].reshape(-1, rows).T.copy() if (transpose and spec[i][0] == "connected") else p["w"].reshape(rows, -1)
            params[i] = p
        return params, dict(major=major, minor=minor, revision=revision, seen=seen)
    finally:
        fp.close()


def write_weights(params, spec, c, h, w, seen=0, version=(0, 2, 5)):
    """-> bytes of a .weights file holding `params` (saveWeights, nparser.pas:1185-1273, writes the same order)."""
    out = io.BytesIO()
    out.write(struct.pack("<3i", *version))
    out.write(struct.pack("<Q", seen) if version[0] * 10 + version[1] >= 2 else struct.pack("<I", seen))
    transpose = version[0] > 1000 or version[1] > 1000   # the same flag read_weights honours (nparser.pas:1140-1182)
    for i, fields in param_layout(spec, c, h, w):
        for name, n in fields:
            a = np.asarray(params[i][name], dtype="<f4")
            if transpose and spec[i][0] == "connected" and name == "w":
                a = a.reshape(spec[i][1], -1).T   # stored [inputs][outputs] in transposed files
            a = np.ascontiguousarray(a).reshape(-1)
            assert a.size == n, (i, name, a.size, n)
            out.write(a.tobytes())
    return out.getvalue()


# ------------------------------------------------------------------ layers on the device
def build_from_spec(spec, batch, c, h, w, params=None, train=False, fused=True, fuse_bn=False, seed=0, **hyper):
    """Instantiates the layers of a spec on the current backend (what TDarknetParser.Create does after parsing)."""
    from . import nncuda as nn
    from .layers import TConvolutionalLayer
    from .nnet import YOLOV3_ANCHORS, TConcatLayer, TNNet, TShortcutLayer, TUpsampleLayer, TYoloLayer
    from .train_layers import TConnectedLayer, TDropoutLayer, TMaxPoolLayer, TSoftmaxLayer
    act = {"leaky": nn.acLEAKY, "linear": nn.acLINEAR, "relu": nn.acRELU, "logistic": nn.acLOGISTIC}
    if train:
        graph = sorted({it[0] for it in spec if it[0] in ("shortcut", "route", "upsample", "yolo")})
        if graph:
            raise ValueError(f"training is mirrored for sequential nets only; this spec has {', '.join(graph)} layers "
                             "(their backward passes are outside the path, DESIGN.md section 7)")
    rng = np.random.default_rng(seed)
    layers = []
    c0, h0, w0 = c, h, w
    for i, item in enumerate(spec):
        kind = item[0]
        p = params[i] if params is not None else None
        if kind == "conv":
            _, f, k, s, bn, a = item[:6]
            pad = item[6] if len(item) > 6 else k // 2
            l = TConvolutionalLayer(batch, h, w, c, f, k, s, pad, act[a], bn, train=train, fused=fused, rng=rng,
                                    host_output=False)
            if p is not None:
                l.weights.Data[...] = p["w"].reshape(l.weights.FShape)
                l.biases.Data[...] = p["b"]
                if bn:
                    l.scales.Data[...] = p["scales"]
                    l.rolling_mean.Data[...] = p["rm"]
                    l.rolling_variance.Data[...] = p["rv"]
            if bn and fuse_bn:
                l.fuseBatchNorm()
        elif kind == "maxpool":
            l = TMaxPoolLayer(batch, h, w, c, item[1], item[2], item[2], item[3], train=train)
        elif kind == "connected":
            l = TConnectedLayer(batch, 1, c * h * w, item[1], act[item[2]], train=train, rng=rng)
            if p is not None:
                l.weights.Data[...] = p["w"].reshape(l.weights.FShape)
                l.biases.Data[...] = p["b"]
        elif kind == "dropout":
            l = TDropoutLayer(batch, item[1], c * h * w, rng=rng)
        elif kind == "softmax":
            l = TSoftmaxLayer(batch, c * h * w)
        elif kind == "shortcut":
            j = i + item[1] if item[1] < 0 else item[1]
            l = TShortcutLayer(batch, j, layers[j].outC, layers[j].outH, layers[j].outW)
        elif kind == "route":
            idx = [(i + j) if j < 0 else j for j in item[1]]
            l = TConcatLayer(batch, idx, [(layers[j].outC, layers[j].outH, layers[j].outW) for j in idx])
        elif kind == "upsample":
            l = TUpsampleLayer(batch, c, h, w, item[1])
        elif kind == "yolo":
            anchors = item[2] if len(item) > 2 else YOLOV3_ANCHORS
            l = TYoloLayer(batch, c, h, w, item[1], anchors, item[3] if len(item) > 3 else 80, fused=fused)
        else:
            raise ValueError(kind)
        layers.append(l)
        if kind in ("connected", "softmax", "dropout"):
            c, h, w = getattr(l, "outputs"), 1, 1
        else:
            c, h, w = l.outC, l.outH, l.outW
    return TNNet(layers, batch, c0, h0, w0, **hyper)


def load_cfg(text, batch=None, weights=None, **kw):
    """cfg text (+ optional .weights path or bytes) -> TNNet; `batch` overrides the [net] batch."""
    net, spec = cfg_to_spec(parse_cfg(text))
    c, h, w = net["channels"], net["height"], net["width"]
    params = None
    seen = 0
    if weights is not None:
        params, hdr = read_weights(weights, spec, c, h, w)
        seen = hdr["seen"]
    hyper = dict(learningRate=net["learning_rate"], momentum=net["momentum"], decay=net["decay"],
                 subdivisions=net["subdivisions"])
    hyper.update(kw.pop("hyper", {}))
    n = build_from_spec(spec, batch or net["batch"], c, h, w, params=params, **kw, **hyper)
    n.seen = seen
    return n


def net_params(net, spec):
    """Current parameters of a built net, aligned with its spec (pulled from the device when they live there)."""
    out = [None] * len(spec)
    for i, (item, l) in enumerate(zip(spec, net.layers)):
        if item[0] not in ("conv", "connected"):
            continue

        def get(t):
            if not t.wasGPU():
                return t.Data.copy()
            a = t.pullFromDevice().copy()
            t.setCUDA()
            return a

        p = dict(w=get(l.weights), b=get(l.biases))
        if item[0] == "conv" and item[4]:
            p.update(scales=get(l.scales), rm=get(l.rolling_mean), rv=get(l.rolling_variance))
        out[i] = p
    return out
