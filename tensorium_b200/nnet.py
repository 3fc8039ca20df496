"""TNNet and the YOLOv3 graph layers — host-side mirror of the callers either side of the conv hot path
(SURVEY.md §8f rank 2): shortcut (source/naddlayer.pas:885-910), route (source/nconcatlayer.pas:205-222),
upsample (source/nupsamplelayer.pas:190-206), yolo (source/nyololayer.pas:97-122,399-498,1068-1112), the
sequential net driver (source/nnet.pas:275-472) and NMS (source/ntypes.pas:1059-1108,1612-1646).
"""
from __future__ import annotations

import numpy as np

from . import nncuda as nn
from .layers import TConvolutionalLayer, TNNetState, TUpdateArgs
from .tensor import TSingleTensor, cuda, split_view


class TShortcutLayer:
    """TAddLayer.forwardGPU (naddlayer.pas:885-910): output = input + layers[from].output, then activation."""
    layerType = "shortcut"

    def __init__(self, batch, index_from, c, h, w, activation=nn.acLINEAR):
        self.batch, self.c, self.h, self.w = batch, c, h, w
        self.outC, self.outH, self.outW = c, h, w
        self.inputLayers = [index_from]
        self.ActivationType = activation
        self.output = TSingleTensor((batch, c, h, w), groups=batch, host=False)

    def forwardGPU(self, state: TNNetState, net: "TNNet"):
        cu = cuda()
        self.output.setCUDA()
        if getattr(self, "_fused_into_conv", False) and not state.isTraining:
            return  # the preceding convolution's epilogue already wrote input + layers[from].output here
        other = net.layers[self.inputLayers[0]].output
        # (channels-innermost graph: both operands are dense rows of c floats — the same flat add)
        # copy + addvv of the reference collapse into one out-of-place add (same arithmetic: a single fp32 add)
        cu.addvv(self.output.Size(), state.input.devData, 0, 1, other.devData, 0, 1, self.output.devData, 0, 1)
        cu.ActivateArray(self.output.Size(), self.output.devData, 0, self.ActivationType)


class TConcatLayer:
    """Route.  The reference copies whole tensors back to back (nconcatlayer.pas:214-220), which equals a
    per-image channel concat only for batch=1; here every image is concatenated along C."""
    layerType = "route"

    def __init__(self, batch, input_layers, shapes):
        self.batch = batch
        self.inputLayers = list(input_layers)
        self.h, self.w = shapes[0][1], shapes[0][2]
        self.outC = sum(s[0] for s in shapes)
        self.outH, self.outW = self.h, self.w
        self._cs = [s[0] for s in shapes]
        self.output = TSingleTensor((batch, self.outC, self.h, self.w), groups=batch, host=False)

    def forwardGPU(self, state: TNNetState, net: "TNNet"):
        cu = cuda()
        self.output.setCUDA()
        if getattr(self, "_elided", False):
            return  # the sources already live in this buffer (TNNet.elideRoutes)
        off = 0
        for li, c in zip(self.inputLayers, self._cs):
            src = net.layers[li].output
            if getattr(self, "layout", "NCHW") == "NHWC":
                # channels-innermost: every source is a column slice of the output rows
                rows = self.batch * self.h * self.w
                if getattr(self, "f32_rows", True):
                    cu.rowsCopy(rows, c, src.devData, getattr(src, "pitch", 0) or c, self.output.devData + 4 * off, self.outC)
                if getattr(self, "split_planes", False):
                    # pre-split tensors (TNNet.setSplitActivations): the same slice copy on the hi and the lo plane —
                    # rows of c bf16 = c/2 floats, planes rows*C bf16 elements apart
                    sp, spitch, splane = split_view(src)
                    for plane in (0, 1):
                        cu.rowsCopy(rows, c // 2, sp + 2 * plane * splane, spitch // 2,
                                    self.output.split.devData + 2 * (plane * rows * self.outC + off), self.outC // 2)
            elif self.batch == 1 or len(self.inputLayers) == 1:
                cu.copy(src.Size(), src.devData, 0, 1, self.output.devData, off * self.h * self.w, 1)
            else:
                cu.concatChannels(self.batch, c, self.h * self.w, src.devData, self.outC, off, self.output.devData)
            off += c


class TUpsampleLayer:
    """nupsamplelayer.pas:190-206 (forward, non-reverse)."""
    layerType = "upsample"

    def __init__(self, batch, c, h, w, stride=2, scale=1.0):
        self.batch, self.c, self.h, self.w, self.stride, self.scale = batch, c, h, w, stride, scale
        self.outC, self.outH, self.outW = c, h * stride, w * stride
        self.output = TSingleTensor((batch, c, self.outH, self.outW), groups=batch, host=False)

    def forwardGPU(self, state: TNNetState, net: "TNNet"):
        self.output.setCUDA()
        if getattr(self, "layout", "NCHW") == "NHWC":
            if getattr(self, "f32_rows", True):
                cuda().upSampleNhwc(self.batch, self.c, self.h, self.w, self.stride, self.scale, state.input.devData,
                                    getattr(state.input, "pitch", 0) or self.c, self.output.devData, self.c)
            if getattr(self, "split_planes", False):
                assert self.scale == 1.0, "pre-split planes are repeated, not rescaled"
                ip, ipitch, iplane = split_view(state.input)
                op, opitch, oplane = split_view(self.output)   # (its own planes, or a slice of an elided route's buffer)
                for plane in (0, 1):  # bf16 rows seen as rows of c/2 floats
                    cuda().upSampleNhwc(self.batch, self.c // 2, self.h, self.w, self.stride, 1.0,
                                        ip + 2 * plane * iplane, ipitch // 2, op + 2 * plane * oplane, opitch // 2)
            return
        cuda().upSample(self.batch, self.c, self.h, self.w, state.input.devData, self.stride, 1, self.scale,
                        self.output.devData)


class TYoloLayer:
    """nyololayer.pas: forwardGPU :1068-1112, entryIndex :97-103, getBox :105-122, getDetections :399-441."""
    layerType = "yolo"

    def __init__(self, batch, c, h, w, mask, anchors, classes=80, scale_xy=1.0, fused=True):
        self.batch, self.c, self.h, self.w = batch, c, h, w
        self.outC, self.outH, self.outW = c, h, w
        self.mask, self.n, self.classes = list(mask), len(mask), classes
        self.biases = np.asarray(anchors, dtype=np.float32)
        self.scaleXY = scale_xy
        self.outputs = h * w * self.n * (classes + 4 + 1)
        assert c == self.n * (classes + 5)
        self.fused = fused
        self.output = TSingleTensor((batch, c, h, w), groups=batch, host=False)

    def entryIndex(self, abatch, location, entry):
        wh = self.w * self.h
        _n, loc = divmod(location, wh)
        return abatch * self.outputs + _n * (self.classes + 5) * wh + entry * wh + loc

    def forwardGPU(self, state: TNNetState, net: "TNNet"):
        cu = cuda()
        self.output.setCUDA()
        if getattr(self, "_fused_into_conv", False):
            return  # the head convolution's epilogue already wrote the activated output here (TNNet.fuseYoloHeads)
        wh = self.w * self.h
        cu.copy(state.input.Size(), state.input.devData, 0, 1, self.output.devData, 0, 1)
        if self.fused:
            cu.yoloActivate(self.batch, self.n, self.classes, wh, self.output.devData, self.scaleXY)
            return
        for b in range(self.batch):  # the reference's 2 launches per (image, anchor)
            for _n in range(self.n):
                bbox_index = self.entryIndex(b, _n * wh, 0)
                cu.ActivateArray(2 * wh, self.output.devData, bbox_index, nn.acLOGISTIC)
                obj_index = self.entryIndex(b, _n * wh, 4)
                cu.ActivateArray((1 + self.classes) * wh, self.output.devData, obj_index, nn.acLOGISTIC)
                if self.scaleXY != 1:
                    cu.fmavss(2 * wh, self.output.devData, bbox_index, self.scaleXY, 0.5 * (1 - self.scaleXY),
                              self.output.devData)

    def getDetections(self, pred: np.ndarray, netw, neth, thresh, abatch=0, img_w=None, img_h=None, relative=True,
                      letter=False):
        """pred: host copy of this layer's output [batch, c, h, w].  Returns (boxes[n,4] xywh, objectness[n],
        prob[n,classes]) following nyololayer.pas:399-441 (+ correctBoxes :443-498; the MSCOCO sample calls
        Detections(img.w, img.h, thresh, relative=true, letter=false), MSCOCOYolo.pas:364)."""
        wh = self.w * self.h
        p = pred[abatch].reshape(self.n, self.classes + 5, wh)
        cells = []
        for i in range(wh):
            for _n in range(self.n):
                if p[_n, 4, i] > thresh:
                    cells.append((i, _n, p[_n, :, i]))
        return self.decodeCells(cells, netw, neth, thresh, img_w, img_h, relative, letter)

    def decodeCells(self, cells, netw, neth, thresh, img_w=None, img_h=None, relative=True, letter=False):
        """The body of getDetections' loop for the cells that passed `objectness > thresh`, in the reference's visiting
        order: cells = [(cell index, anchor, the classes + 5 entries of that cell)] — from the full output
        (getDetections) or from the device-side candidate list (tnn_yolo_candidates, TNNet.collectCandidates)."""
        boxes, objs, probs = [], [], []
        for i, _n, e in cells:
            row, col = divmod(i, self.w)
            objectness = e[4]
            a = self.mask[_n]
            bx = (col + e[0]) / self.w
            by = (row + e[1]) / self.h
            bw = np.exp(e[2]) * self.biases[2 * a] / netw
            bh = np.exp(e[3]) * self.biases[2 * a + 1] / neth
            pr = objectness * e[5:]
            pr = np.where(pr > thresh, pr, 0).astype(np.float32)
            boxes.append((bx, by, bw, bh))
            objs.append(objectness)
            probs.append(pr)
        if not boxes:
            return np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros((0, self.classes), np.float32)
        boxes = correctBoxes(np.asarray(boxes, np.float32), img_w or netw, img_h or neth, netw, neth, relative, letter)
        return boxes, np.asarray(objs, np.float32), np.stack(probs)


def correctBoxes(boxes, w, h, netw, neth, relative=True, letter=False):
    """TYoloLayer.correctBoxes nyololayer.pas:443-498: undo the letter-box embedding of a (w x h) image in the
    (netw x neth) network input; relative=False scales to image pixels."""
    if letter:
        if (netw / w) < (neth / h):
            new_w, new_h = netw, (h * netw) // w
        else:
            new_h, new_w = neth, (w * neth) // h
    else:
        new_w, new_h = netw, neth
    deltaw, deltah = np.float32(netw - new_w), np.float32(neth - new_h)
    ratiow, ratioh = np.float32(new_w / netw), np.float32(new_h / neth)
    b = boxes.astype(np.float64)
    b[:, 0] = (b[:, 0] - deltaw / 2.0 / netw) / ratiow
    b[:, 1] = (b[:, 1] - deltah / 2.0 / neth) / ratioh
    b[:, 2] *= 1 / np.float64(ratiow)
    b[:, 3] *= 1 / np.float64(ratioh)
    b = b.astype(np.float32)
    if not relative:
        b[:, 0] *= np.float32(w); b[:, 2] *= np.float32(w)
        b[:, 1] *= np.float32(h); b[:, 3] *= np.float32(h)
    return b


def box_iou(a, b):
    """ntypes.pas:1059-1108 (xywh boxes)."""
    def overlap(x1, w1, x2, w2):
        return min(x1 + w1 / 2, x2 + w2 / 2) - max(x1 - w1 / 2, x2 - w2 / 2)
    w, h = overlap(a[0], a[2], b[0], b[2]), overlap(a[1], a[3], b[1], b[3])
    if w < 0 or h < 0:
        return 0.0
    i = w * h
    u = a[2] * a[3] + b[2] * b[3] - i
    return 0.0 if i == 0 or u == 0 else float(i / u)


def doNMSSort(boxes, objs, probs, thresh):
    """TDetectionsHelper.doNMSSort ntypes.pas:1612-1646: per class, sort by class probability, suppress IoU > thresh."""
    keep = objs != 0
    boxes, objs, probs = boxes[keep], objs[keep], probs[keep].copy()
    for k in range(probs.shape[1] if probs.size else 0):
        order = np.argsort(-probs[:, k], kind="stable")
        for ii, i in enumerate(order):
            if probs[i, k] == 0:
                continue
            for j in order[ii + 1:]:
                if probs[j, k] != 0 and box_iou(boxes[i], boxes[j]) > thresh:
                    probs[j, k] = 0
    return boxes, objs, probs


class TNNet:
    """Sequential driver, GPU path: TNNet.forward (nnet.pas:275-317), backward (:319-376), update (:378-403),
    Propagate (:404-432), predict (:452-472).  Learning-rate policy is lrpCONSTANT (the CIFAR sample's,
    CIFAR10Conv.pas:150-160), for which computeCurrentLearningRate returns learningRate unchanged."""

    def __init__(self, layers, batch, c, h, w, learningRate=0.001, momentum=0.9, decay=0.0001, subdivisions=1):
        self.layers = layers
        self.batch, self.c, self.h, self.w = batch, c, h, w
        self.learningRate, self.momentum, self.decay, self.subdivisions = learningRate, momentum, decay, subdivisions
        self.seen = 0
        self.replicas = 1  # set by dataparallel.GradientBucket
        self.bucket = None
        self.input = TSingleTensor((batch, c, h, w), groups=batch, host=False)
        self.truth = None
        # state.workspace (im2col columns) is only needed by the reference's unfused call sequence
        ws = max([l.getWorkspaceSize() for l in layers if isinstance(l, TConvolutionalLayer) and not l.fused] + [0])
        self.workspace = TSingleTensor((ws,), groups=1, host=False) if ws else None

    def forward(self, state: TNNetState):
        state.workspace = self.workspace
        cu = cuda()
        for i, l in enumerate(self.layers):
            state.index = i
            delta = getattr(l, "delta", None)
            if state.isTraining and delta is not None and getattr(l, "train", False):
                cu.scale(delta.Size(), 0.0, delta.devData, 1)  # nnet.pas:288-293
                delta.setCUDA()
            l.forwardGPU(state, self)
            state.input = l.output
        return state.input

    def backward(self, state: TNNetState):
        original_input, original_delta = self.input, state.delta
        state.workspace = self.workspace
        for i in range(len(self.layers) - 1, -1, -1):
            state.index = i
            if i == 0:
                state.input, state.delta = original_input, original_delta
            else:
                prev = self.layers[i - 1]
                state.input, state.delta = prev.output, prev.delta
            self.layers[i].backwardGPU(state)
        self.seen += self.batch

    def update(self):
        """nnet.pas:378-403: every layer but the last; a data-parallel run first sums the update buffers."""
        if self.bucket is not None:
            self.bucket.all_reduce()
        self._update_layers()

    def _update_layers(self):
        arg = TUpdateArgs(batch=self.batch * self.subdivisions * self.replicas, learningRate=self.learningRate,
                          momentum=self.momentum, decay=self.decay, replicas=self.replicas)
        todo = [l for l in self.layers[:-1] if getattr(l, "trainable", False) and l.train]
        if todo and all(getattr(l, "fused", True) and hasattr(l, "sgdItems") for l in todo):
            # fused entry points: every parameter tensor of the net in one launch (tnn_sgd_update_multi) instead of two
            # or three launches per layer; same arithmetic per tensor as updateGPU
            items = [it for l in todo for it in l.sgdItems(arg)]
            cuda().sgdUpdateMulti(items, arg.momentum / arg.replicas)
            return
        for l in todo:
            l.updateGPU(arg)

    def Propagate(self):
        """nnet.pas:404-432; `input` and `truth` must already hold the batch (host or device side)."""
        state = TNNetState(input=self.input, isTraining=True, truth=self.truth)
        self.forward(state)
        state.delta = None
        self.backward(state)

    def trainBatch(self, x: np.ndarray | None = None, truth: np.ndarray | None = None):
        """One iteration of trainEpoch's loop body (nnet.pas:497-520): load, Propagate, update when due."""
        cu = cuda()
        if self.truth is None:
            out = self.layers[-1].output
            self.truth = TSingleTensor(out.FShape, groups=self.batch, host=False)
        if x is not None and truth is not None:
            # The reference copies the batch synchronously (input.setCPU -> pushToDevice).  Here the batch goes through
            # two page-locked staging buffers and asynchronous copies, so the host keeps enqueuing the next step
            # while the GPU runs this one; an event per buffer keeps the host from overwriting a buffer whose copy
            # has not run yet.
            if getattr(self, "_stage", None) is None:
                self._stage = [(cu.allocHost(self.input.FShape), cu.allocHost(self.truth.FShape), cu.createEvent())
                               for _ in range(2)]
                for _, _, ev in self._stage:
                    cu.recordEvent(ev)
                self._stage_i = 0
            sx, st, ev = self._stage[self._stage_i & 1]
            self._stage_i += 1
            cu.syncEvent(ev)
            sx[...] = x.reshape(self.input.FShape)
            st[...] = truth.reshape(self.truth.FShape)
            cu.writeBufferAsync(self.input.devData, sx.nbytes, sx.ctypes.data)
            cu.writeBufferAsync(self.truth.devData, st.nbytes, st.ctypes.data)
            cu.recordEvent(ev)
            self.input.setCUDA()
            self.truth.setCUDA()
        else:
            if x is not None:
                self.input.Data[...] = x.reshape(self.input.FShape)
                self.input.setCPU()
            if truth is not None:
                self.truth.Data[...] = truth.reshape(self.truth.FShape)
                self.truth.setCPU()
        if getattr(self, "_graphs", None) is not None:
            return self._replay()
        self.Propagate()
        if self.seen % (self.batch * self.subdivisions) == 0:
            self.update()

    # ---- CUDA-graph replay of the training step (the step is ~110 small launches: launch-bound from any host)
    GRAPH_SEED_STEP = 0x9E3779B97F4A7C15  # the dropout seed counter advances by this much per step (masks decorrelate)

    def enableGraph(self, seed: int = 0x1234567):
        """Capture forward+backward (and update) of ONE training step into CUDA graphs and replay them from then on.
        Call after at least one eager trainBatch (scratch buffers must have their final size, every parameter must be
        on the device).  Dropout seeds move to a device counter: layer i, step t uses seed + t*GRAPH_SEED_STEP +
        (i+1)*2^40.  A native gradient bucket (dataparallel.NativeBucket: the library's NCCL entry point) is captured
        with the step — forward + backward + all-reduce + update are ONE graph; with a torch bucket (gloo tests) the
        all-reduce stays eager between a forward/backward graph and an update graph."""
        assert self.subdivisions == 1, "graph mode updates every step"
        cu = cuda()
        self._seed_ctr = TSingleTensor((1,), groups=1, host=True, dtype=np.int64)
        self._seed_ctr.Data[0] = np.int64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF).astype(np.int64))
        self._seed_ctr.pushToDevice()
        self._seed_host = seed & 0xFFFFFFFFFFFFFFFF
        for i, l in enumerate(self.layers):
            if hasattr(l, "seed_dev"):
                l.seed_dev = (self._seed_ctr.devData, ((i + 1) << 40) & 0xFFFFFFFFFFFFFFFF)
        cu.finish()
        cu.beginCapture()
        self.Propagate()
        cu.counterAdd(self._seed_ctr.devData, self.GRAPH_SEED_STEP)
        if self.bucket is None or getattr(self.bucket, "capturable", False):
            # (a native bucket's all-reduce is an operator on the backend's stream: it is captured with the step)
            self.update()
            self._graphs = (cu.endCapture(), None)
        else:
            g1 = cu.endCapture()
            cu.beginCapture()
            self._update_layers()
            self._graphs = (g1, cu.endCapture())
        self.seen -= self.batch  # the capture pass did not run anything

    def dropout_seed(self, layer_index: int, step: int) -> int:
        """The seed dropout layer `layer_index` used at graph step `step` (0-based) — for parity checks."""
        return (self._seed_host + step * self.GRAPH_SEED_STEP + ((layer_index + 1) << 40)) & 0xFFFFFFFFFFFFFFFF

    def _replay(self):
        cu = cuda()
        g1, g2 = self._graphs
        cu.launchGraph(g1)
        self.seen += self.batch
        for l in self.layers:  # the loss tensor is new after every replay
            if hasattr(l, "_cost_stale"):
                l._cost_stale = True
            if hasattr(l, "_bn_fold"):
                l._bn_fold = None  # scales / rolling statistics / biases moved on the device: a later predict() folds anew
        if g2 is not None:
            self.bucket.all_reduce()
            cu.launchGraph(g2)

    def cost(self) -> float:
        """nnet.pas:551-563: mean of the layers' cost[0]."""
        cs = [l.cost for l in self.layers if hasattr(l, "cost")]
        return float(sum(cs) / len(cs)) if cs else 0.0

    def gradients(self):
        out = []
        for l in self.layers[:-1]:
            if getattr(l, "trainable", False) and l.train:
                out += l.gradients()
        return out

    def fuseShortcuts(self, only_kernels=None):
        """Inference-time graph fusion: a fused convolution whose output feeds ONLY the linear shortcut layer that
        follows adds the shortcut's other operand in its epilogue (tnn_conv2d_forward_residual) and writes the
        shortcut's output directly; the add kernel and one round trip of the tensor through HBM disappear.  Same
        arithmetic as the separate layers (one fp32 add after the activation).  Returns the number of fused pairs.

        `only_kernels`: substrings of TNNCuda.lastConvKernel() — fuse only convolutions that ran on one of those
        kernels in a PREVIOUS forward.

        OFF by default.  Measured on B200 (YOLOv3-416, batch 64): 27.7 ms un-fused, 28.0 ms with the gathering-kernel
        layers fused, 28.9 ms with all fused (residual loads batched 16 columns at a time; column by column it was
        47.6 ms).  The read costs the four epilogue warps what the separate add kernels cost; it would take a
        TMA-staged residual tile to win."""
        used = set()  # layer outputs referenced by index from elsewhere in the graph
        for l in self.layers:
            used.update(getattr(l, "inputLayers", []))
        n = 0
        for i in range(len(self.layers) - 1):
            c, sc = self.layers[i], self.layers[i + 1]
            if (isinstance(c, TConvolutionalLayer) and c.fused and isinstance(sc, TShortcutLayer)
                    and sc.ActivationType == nn.acLINEAR and i not in used and sc.inputLayers[0] < i
                    and sc.output.Size() == c.output.Size()
                    and (only_kernels is None or any(k in getattr(c, "_last_kernel", "") for k in only_kernels))):
                c._residual = self.layers[sc.inputLayers[0]].output
                c._fused_out = sc.output
                sc._fused_into_conv = True
                n += 1
        return n

    def elideRoutes(self, on: bool = True):
        """Planes-only inference graph (after setSplitActivations(planes_only=True)): route layers stop copying.
        A single-source route is an alias of its source.  A concatenating route becomes the OWNER of the tensor: every
        source's planes are a column slice of the route's buffer (tensor.split_view: pitch = the route's channels), so the
        producing epilogue / upsample writes there and every other reader of the source reads there — the pitch and
        offset arguments of the C ABI do the rest.  Returns the number of elided routes.  on=False puts the copying
        routes back (setSplitActivations does so before it re-assigns the tensor forms)."""
        if not on:
            for l in self.layers:
                if isinstance(l, TConcatLayer) and getattr(l, "_elided", False):
                    l.output = l._own_output
                    for i in l.inputLayers:
                        if getattr(self.layers[i].output, "split_at", None) is not None:
                            self.layers[i].output.split_at = None
                    l._elided = False
            return 0
        assert getattr(self, "split_activations", False), "routes are elided in the pre-split graph"
        n = 0
        for j, l in enumerate(self.layers):
            if not isinstance(l, TConcatLayer) or not getattr(l, "split_planes", False) or getattr(l, "f32_rows", True):
                continue
            srcs = [self.layers[i].output for i in l.inputLayers]
            if any(getattr(t, "split_at", None) is not None or getattr(t, "f32_valid", False) for t in srcs):
                continue   # a source that already lives in another route's buffer, or that something reads as fp32
            l._own_output = l.output
            if len(srcs) == 1:
                l.output = srcs[0]
            else:
                rows = self.batch * l.h * l.w
                off = 0
                for t, c in zip(srcs, l._cs):
                    t.split_at = (l.output.split, off, l.outC, rows * l.outC)
                    off += c
            l._elided = True
            n += 1
        return n

    def fuseYoloHeads(self):
        """Channels-innermost inference graph: a linear head convolution whose output feeds ONLY the yolo layer that
        follows applies that layer's activation in its epilogue (tnn_conv2d_forward_yolo_head) and writes the yolo
        layer's output directly — the copy and the activation pass of TYoloLayer.forwardGPU (nyololayer.pas:1068-1112)
        disappear.  The head convolution's own output tensor is then never written.  Same expressions, same bits
        (tests/test_gpu_yolo.py).  Call after setChannelsLast(); returns the number of fused heads."""
        assert getattr(self, "channels_last", False), "yolo heads are folded in the channels-innermost graph"
        used = set()
        for l in self.layers:
            used.update(getattr(l, "inputLayers", []))
        n = 0
        for i in range(len(self.layers) - 1):
            c, y = self.layers[i], self.layers[i + 1]
            if (isinstance(c, TConvolutionalLayer) and isinstance(y, TYoloLayer) and c.fused and y.fused and i not in used
                    and c.ActivationType == nn.acLINEAR and not c.isBatchNormalized and c._residual is None):
                c._yolo = y
                y._fused_into_conv = True
                n += 1
        return n

    def setChannelsLast(self, on: bool = True):
        """Inference graph with channels-innermost (NHWC) INTERNAL activations (include/tensorium_b200.h, last section):
        the first convolution reads the planar image, every convolution that feeds a yolo layer writes planes, and all
        tensors in between are rows of channels — what the TMA im2col-mode convolution kernel reads and writes.
        Shortcut adds are layout-blind, routes copy column slices, upsample repeats rows.  Same arithmetic per layer as
        the planar graph."""
        NCHW, NHWC = nn.TNNCuda.LAYOUT_NCHW, nn.TNNCuda.LAYOUT_NHWC
        n = len(self.layers)
        feeds_yolo = [i + 1 < n and isinstance(self.layers[i + 1], TYoloLayer) for i in range(n)]
        for i, l in enumerate(self.layers):
            if isinstance(l, TConvolutionalLayer):
                assert l.fused, "channels-last needs the fused entry points"
                l.in_layout = NHWC if (on and i > 0) else NCHW
                l.out_layout = NHWC if (on and not feeds_yolo[i]) else NCHW
                l.output.layout = "NHWC" if l.out_layout == NHWC else "NCHW"
            elif isinstance(l, (TShortcutLayer, TConcatLayer, TUpsampleLayer)):
                l.layout = "NHWC" if on else "NCHW"
                l.output.layout = l.layout
            elif isinstance(l, TYoloLayer):
                pass
            else:
                assert not on, f"layer {i} ({l.layerType}) has no channels-innermost path"
        self.channels_last = bool(on)
        return self

    def setSplitActivations(self, on: bool = True, split_into_1x1: bool = False, planes_only: bool = False):
        """BF16x3 graphs: keep every activation that a convolution reads as PRE-SPLIT bf16 planes
        (TNN_LAYOUT_NHWC_SPLIT, include/tensorium_b200.h): the producing epilogue writes hi = bf16(x), lo = bf16(x - hi),
        the consuming convolution's TMA lands them as tensor-core operands and nothing converts — the same products
        and fp32 accumulation as the fp32-input kernel, bit for bit.  fp32 rows are written only where something other
        than a convolution reads the tensor: the operand of a later shortcut, a yolo layer, the graph's last layer.
        split_into_1x1=False: a 1x1 convolution reads fp32 rows and splits them in shared memory itself — it touches
        every element once (no 9x redundancy as in a 3x3 layer) and is bound by HBM, while its input is typically a
        shortcut output that a later shortcut needs as fp32 anyway: the tensor is then written once instead of twice.
        planes_only=True: every tensor between convolutions exists ONLY as planes.  The folded shortcuts read their other
        operand from its planes too (tnn_conv2d_forward_split_residual: residual = hi + lo, the fp32 value to 2^-18
        relative), so no tensor is written twice AND the 1x1 convolutions take their operands by TMA with no conversion
        in shared memory.  fp32 is written only for the yolo layers and the graph's last layer.
        Call after setChannelsLast() and fuseShortcuts()."""
        n = len(self.layers)
        split_into_1x1 = split_into_1x1 or planes_only
        self.elideRoutes(False)
        if not on:
            for l in self.layers:
                for a in ("in_split", "out_split", "split_planes", "res_split"):
                    if hasattr(l, a):
                        setattr(l, a, False)
                for a in ("out_f32", "f32_rows"):
                    if hasattr(l, a):
                        setattr(l, a, True)
                l.output.f32_valid = True
            self.split_activations = False
            return self
        assert getattr(self, "channels_last", False), "pre-split activations live in the channels-innermost graph"
        need_split, need_f32 = [False] * n, [False] * n
        need_f32[n - 1] = True
        for j in range(n - 1, -1, -1):
            l = self.layers[j]
            if isinstance(l, TConvolutionalLayer):
                if j > 0:
                    if l.kernelSize == 1 and not split_into_1x1:
                        need_f32[j - 1] = True
                    else:
                        need_split[j - 1] = True
            elif isinstance(l, TShortcutLayer):
                assert getattr(l, "_fused_into_conv", False), "pre-split graphs need the shortcuts folded into the convolutions"
                if planes_only:
                    need_split[l.inputLayers[0]] = True
                else:
                    need_f32[l.inputLayers[0]] = True
            elif isinstance(l, TConcatLayer):
                for idx in l.inputLayers:
                    need_split[idx] |= need_split[j]
                    need_f32[idx] |= need_f32[j]
            elif isinstance(l, TUpsampleLayer):
                need_split[j - 1] |= need_split[j]
                need_f32[j - 1] |= need_f32[j]
            elif isinstance(l, TYoloLayer):
                need_f32[j - 1] = True
        for j, l in enumerate(self.layers):
            if isinstance(l, TConvolutionalLayer):
                t = j + 1 if l._residual is not None else j     # the tensor this convolution writes
                tgt = self.layers[t].output
                l.in_split = j > 0 and (l.kernelSize != 1 or split_into_1x1)
                l.out_split = need_split[t]
                l.res_split = planes_only and l._residual is not None
                # (an image layer too deep for the tensor-core image kernel runs the direct kernel, which writes rows first)
                l.out_f32 = need_f32[t] or not need_split[t] or (j == 0 and l.c * l.kernelSize ** 2 > 32)
                assert not (l.out_split and l.out_layout != nn.TNNCuda.LAYOUT_NHWC), "planes are fp32 only"
                assert l.c % 8 == 0 or j == 0, "pre-split rows need a channel count that is a multiple of 8"
            elif isinstance(l, (TConcatLayer, TUpsampleLayer)):
                t, tgt = j, l.output
                l.split_planes = need_split[j]
                l.f32_rows = need_f32[j] or not need_split[j]
            else:
                continue
            tgt.f32_valid = getattr(l, "out_f32", False) or getattr(l, "f32_rows", False)
            if need_split[t] and getattr(tgt, "split", None) is None:
                tgt.split = TSingleTensor(tgt.FShape, groups=self.batch, host=False)   # hi plane | lo plane, 4 B/element
        self.split_activations = True
        return self

    def output_nchw(self, i: int) -> np.ndarray:
        """Host copy of layer i's output in the reference's planar order, whatever the internal layout."""
        l = self.layers[i]
        if not getattr(l.output, "f32_valid", True):
            # only the pre-split planes were written: fp32 = hi + lo (exact) into the tensor's idle fp32 twin
            b, c, h, w = l.output.FShape
            sp, spitch, splane = split_view(l.output)
            cuda().splitToNhwc(b * h * w, c, sp, spitch, splane, l.output.devData, c)
        data = l.output.pullFromDevice()
        l.output.setCUDA()
        if getattr(l.output, "layout", "NCHW") == "NHWC":
            b, c, h, w = l.output.FShape
            return np.ascontiguousarray(data.reshape(b, h, w, c).transpose(0, 3, 1, 2))
        return data.copy()

    def predict(self, x_dev_tensor: TSingleTensor | None = None):
        st = TNNetState(input=x_dev_tensor or self.input, isTraining=False)
        return self.forward(st)

    def yolo_layers(self):
        return [l for l in self.layers if isinstance(l, TYoloLayer)]

    def conv_flops(self) -> int:
        """Algorithmic FLOPs of one forward pass over the whole batch: 2*B*F*OH*OW*C*kh*kw per conv."""
        return sum(2 * l.batch * l.filters * l.outH * l.outW * l.c * l.kernelSize ** 2 for l in self.layers
                   if isinstance(l, TConvolutionalLayer))

    def Detections(self, thresh=0.45, nms=0.45, abatch=0, host_outputs=None, img_w=None, img_h=None, relative=True,
                   letter=False):
        """nnet.pas:584-677 + MSCOCOYolo.pas:252-253,364: gather yolo detections of one image, then NMS."""
        bs, os_, ps = [], [], []
        for li, l in enumerate(self.yolo_layers()):
            pred = host_outputs[li] if host_outputs is not None else l.output.pullFromDevice()
            b, o, p = l.getDetections(pred, self.w, self.h, thresh, abatch, img_w, img_h, relative, letter)
            bs.append(b), os_.append(o), ps.append(p)
        boxes, objs, probs = np.concatenate(bs), np.concatenate(os_), np.concatenate(ps)
        return doNMSSort(boxes, objs, probs, nms)


    # ---- device-side candidate test (tnn_yolo_candidates): the host reads a short list, not the output tensors
    def collectCandidates(self, thresh=0.45, max_cand=8192, cand: TSingleTensor | None = None) -> TSingleTensor:
        """After predict(): every (image, anchor, cell) of the yolo layers whose objectness exceeds `thresh` is appended
        to one compact list on the device (layer = tag).  Returns the device tensor: [count, -, -, -] then records of
        4 + classes + 5 floats."""
        yl = self.yolo_layers()
        rec = yl[0].classes + 9
        if cand is None:
            cand = getattr(self, "_cand", None)
            if cand is None or cand.Size() != 4 + max_cand * rec:
                cand = self._cand = TSingleTensor((4 + max_cand * rec,), groups=1, host=False)
        cu = cuda()
        cand.setCUDA()
        cu.fill(4, cand.devData, 0, 0.0, 1)
        for tag, l in enumerate(yl):
            cu.yoloCandidates(l.batch, l.n, l.classes, l.w * l.h, l.output.devData, thresh, tag, max_cand, cand.devData)
        return cand

    def DetectionsFromCandidates(self, cand_host: np.ndarray, thresh=0.45, nms=0.45, abatch=0, img_w=None, img_h=None,
                                 relative=True, letter=False):
        """Detections() from the host copy of collectCandidates()' list: the records of image `abatch` are put back into
        the order in which getDetections visits them (layer, cell, anchor) and decoded by the same code."""
        yl = self.yolo_layers()
        rec = yl[0].classes + 9
        count = int(cand_host[:1].view(np.uint32)[0])
        cap = (cand_host.size - 4) // rec
        if count > cap:
            raise OverflowError(f"{count} candidates, the list holds {cap}: read the full outputs (Detections)")
        r = cand_host[4:4 + count * rec].reshape(count, rec)
        key = r[:, :4].view(np.int32)
        bs, os_, ps = [], [], []
        for tag, l in enumerate(yl):
            sel = np.nonzero((key[:, 0] == tag) & (key[:, 1] == abatch))[0]
            sel = sel[np.lexsort((key[sel, 2], key[sel, 3]))]   # by cell, then anchor
            cells = [(int(key[j, 3]), int(key[j, 2]), r[j, 4:]) for j in sel]
            b, o, p = l.decodeCells(cells, self.w, self.h, thresh, img_w, img_h, relative, letter)
            bs.append(b), os_.append(o), ps.append(p)
        boxes, objs, probs = np.concatenate(bs), np.concatenate(os_), np.concatenate(ps)
        return doNMSSort(boxes, objs, probs, nms)


YOLOV3_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]


def yolov3_spec():
    """The standard Darknet-53 / YOLOv3 layer list (the .cfg is not in the reference tree: MSCOCOYolo.pas:31 points
    outside it).  107 layers: 75 conv, 23 shortcut, 4 route, 2 upsample, 3 yolo."""
    L = []
    conv = lambda f, k, s=1, bn=1, act="leaky": L.append(("conv", f, k, s, bn, act))

    def res(n, f):
        for _ in range(n):
            conv(f // 2, 1)
            conv(f, 3)
            L.append(("shortcut", -3))

    conv(32, 3)
    conv(64, 3, 2); res(1, 64)
    conv(128, 3, 2); res(2, 128)
    conv(256, 3, 2); res(8, 256)
    conv(512, 3, 2); res(8, 512)
    conv(1024, 3, 2); res(4, 1024)
    for _ in range(3):
        conv(512, 1); conv(1024, 3)
    conv(255, 1, 1, 0, "linear"); L.append(("yolo", [6, 7, 8]))
    L.append(("route", [-4])); conv(256, 1); L.append(("upsample", 2)); L.append(("route", [-1, 61]))
    for _ in range(3):
        conv(256, 1); conv(512, 3)
    conv(255, 1, 1, 0, "linear"); L.append(("yolo", [3, 4, 5]))
    L.append(("route", [-4])); conv(128, 1); L.append(("upsample", 2)); L.append(("route", [-1, 36]))
    for _ in range(3):
        conv(128, 1); conv(256, 3)
    conv(255, 1, 1, 0, "linear"); L.append(("yolo", [0, 1, 2]))
    assert len(L) == 107
    return L


def layer_channels(spec, c_in=3):
    """Output channel count of every layer of a spec (routes sum their sources)."""
    outc, c = [], c_in
    for i, item in enumerate(spec):
        if item[0] == "conv":
            c = item[1]
        elif item[0] == "route":
            c = sum(outc[(i + j) if j < 0 else j] for j in item[1])
        outc.append(c)
    return outc


def init_yolov3_params(spec, c_in=3, seed=1234, objectness_bias=-4.0):
    """Random-init parameters of BASELINE config 4: weights U(-s,s), s=sqrt(2/(k*k*C))
    (nConvolutionLayer.pas:216-220); BN scales=1, rolling_mean U(-0.1,0.1), rolling_variance U(0.5,1.5); the
    objectness biases of the three head convs = -4 so that only a few hundred boxes pass the 0.45 threshold."""
    rng = np.random.default_rng(seed)
    outc = layer_channels(spec, c_in)
    params = []
    for i, item in enumerate(spec):
        if item[0] != "conv":
            params.append(None)
            continue
        _, f, k, s, bn, act = item
        c = outc[i - 1] if i else c_in
        sc = np.sqrt(2.0 / (k * k * c))
        p = dict(w=rng.uniform(-sc, sc, (f, c * k * k)).astype(np.float32), b=np.zeros(f, np.float32))
        if bn:
            p.update(scales=np.ones(f, np.float32), rm=rng.uniform(-0.1, 0.1, f).astype(np.float32),
                     rv=rng.uniform(0.5, 1.5, f).astype(np.float32))
        else:  # head conv: 3 anchors x (4 box + 1 obj + 80 classes)
            p["b"].reshape(3, 85)[:, 4] = objectness_bias
        params.append(p)
    return params


def build_yolov3(batch, size=416, params=None, fuse_bn=True, fused=True, seed=1234, fuse_shortcuts=False,
                 channels_last=False, split_activations=False, fuse_yolo=False, elide_routes=False):
    """Builds the 107-layer net on the current `cuda` backend.  fuse_bn=True mirrors MSCOCOYolo.pas:327-340
    (Neural.fuseBatchNorm before predict)."""
    spec = yolov3_spec()
    params = params or init_yolov3_params(spec, seed=seed)
    act = {"leaky": nn.acLEAKY, "linear": nn.acLINEAR}
    layers = []
    c, h, w = 3, size, size
    for i, item in enumerate(spec):
        kind = item[0]
        if kind == "conv":
            _, f, k, s, bn, a = item
            l = TConvolutionalLayer(batch, h, w, c, f, k, s, k // 2, act[a], bn, fused=fused, host_output=False)
            p = params[i]
            l.weights.Data[...] = p["w"]
            l.biases.Data[...] = p["b"]
            if bn:
                l.scales.Data[...] = p["scales"]
                l.rolling_mean.Data[...] = p["rm"]
                l.rolling_variance.Data[...] = p["rv"]
                if fuse_bn:
                    l.fuseBatchNorm()
        elif kind == "shortcut":
            src = layers[i + item[1]]
            l = TShortcutLayer(batch, i + item[1], src.outC, src.outH, src.outW)
        elif kind == "route":
            idx = [(i + j) if j < 0 else j for j in item[1]]
            l = TConcatLayer(batch, idx, [(layers[j].outC, layers[j].outH, layers[j].outW) for j in idx])
        elif kind == "upsample":
            l = TUpsampleLayer(batch, c, h, w, item[1])
        elif kind == "yolo":
            l = TYoloLayer(batch, c, h, w, item[1], YOLOV3_ANCHORS, 80, fused=fused)
        layers.append(l)
        c, h, w = l.outC, l.outH, l.outW
    net = TNNet(layers, batch, 3, size, size)
    if channels_last:
        net.setChannelsLast()
    if fuse_shortcuts and fused:
        net.fuseShortcuts()
    if split_activations:
        net.setSplitActivations(planes_only=split_activations == "planes")
    if fuse_yolo:
        net.fuseYoloHeads()
    if elide_routes:
        net.elideRoutes()
    return net
