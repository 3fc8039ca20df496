"""CPU training step of the reference's CIFAR-10 sample net, built on the oracle primitives.

TEST INFRASTRUCTURE ONLY (see oracle/tensorium_oracle.c header; parity unpinned).
Layers follow the reference's CPU bodies call for call:
  TMaxPoolLayer   nMaxPoolLayer.pas:363-513      TDropoutLayer  ndropoutlayer.pas:82-133
  TConnectedLayer nconnectedlayer.pas:157-360    TSoftmaxLayer  nsoftmaxlayer.pas:139-182
and the driver follows TNNet.forward/backward/update/Propagate (nnet.pas:275-432).
The model is deepCIFAR10 (nmodels.pas:82-111).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc
from .ref_layers import RefConvLayer

f32 = np.float32


class RefMaxPool:
    trainable = False

    def __init__(self, batch, c, h, w, ksize, stride=0, padding=0):
        stride = stride or ksize  # nMaxPoolLayer.pas:78-81
        self.batch, self.c, self.h, self.w = batch, c, h, w
        self.ksize, self.stride, self.padding = ksize, stride, padding
        self.outH = (h + padding - ksize) // stride + 1
        self.outW = (w + padding - ksize) // stride + 1
        self.output = np.zeros((batch, c, self.outH, self.outW), f32)
        self.delta = np.zeros_like(self.output)

    def forward(self, x, training=False):
        self.output, self.indexes = orc.maxpool_forward(np.ascontiguousarray(x, f32), self.ksize, self.stride,
                                                        self.stride, self.padding)
        return self.output

    def backward(self, upstream_delta=None):
        if upstream_delta is not None:
            orc.maxpool_backward(self.indexes, self.delta, upstream_delta)


class RefDropout:
    """In place on the previous layer's output / delta (ndropoutlayer.pas:90 `output := state.input^`)."""
    trainable = False

    def __init__(self, batch, inputs, probability):
        self.batch, self.inputs = batch, inputs
        self.probability = f32(probability)
        self.scale = f32(1.0 / (1.0 - probability))
        self.seed = 0
        self.delta = None

    def forward(self, x, training=False):
        self.output = x  # alias
        if training:
            self.rand = orc.dropout_rand(x.size, self.seed)
            orc.dropout_apply(x, self.rand, self.probability, self.scale)
        return self.output

    def backward(self, upstream_delta=None):
        if upstream_delta is None:
            return
        orc.dropout_apply(upstream_delta, self.rand, self.probability, self.scale)
        self.delta = upstream_delta  # alias (ndropoutlayer.pas:130)


class RefConnected:
    trainable = True

    def __init__(self, batch, inputs, outputs, activation, weights, biases):
        self.batch, self.inputs, self.outputs, self.activation = batch, inputs, outputs, activation
        self.weights = np.ascontiguousarray(weights, f32).reshape(outputs, inputs).copy()
        self.biases = np.ascontiguousarray(biases, f32).copy()
        self.weight_updates = np.zeros_like(self.weights)
        self.bias_updates = np.zeros_like(self.biases)
        self.output = np.zeros((batch, outputs), f32)
        self.delta = np.zeros((batch, outputs), f32)

    def forward(self, x, training=False):  # nconnectedlayer.pas:157-242
        self.input = np.ascontiguousarray(x, f32).reshape(self.batch, self.inputs)
        orc.sgemm(False, True, self.batch, self.outputs, self.inputs, 1.0, self.input, self.inputs, self.weights,
                  self.inputs, 0.0, self.output, self.outputs)
        orc.forward_bias(self.output, self.biases, self.batch)
        orc.activate_array(self.output, self.activation)
        return self.output

    def backward(self, upstream_delta=None):  # nconnectedlayer.pas:244-322
        orc.clamp(self.delta, -1.0, 1.0)
        orc.gradient_array(self.output, self.activation, self.delta)
        orc.add_sums(self.bias_updates, self.delta, self.batch)
        orc.sgemm(True, False, self.outputs, self.inputs, self.batch, 1.0, self.delta, self.outputs, self.input,
                  self.inputs, 1.0, self.weight_updates, self.inputs)
        if upstream_delta is not None:
            orc.sgemm(False, False, self.batch, self.inputs, self.outputs, 1.0, self.delta, self.outputs,
                      self.weights, self.inputs, 1.0, upstream_delta.reshape(self.batch, self.inputs), self.inputs)

    def update(self, learning_rate, momentum, decay, batch, replicas=1):  # nconnectedlayer.pas:324-359
        momentum = momentum / replicas  # data-parallel bookkeeping (tensorium_b200/dataparallel.py)
        lr_b = f32(f32(learning_rate) / f32(batch))
        orc.axpy(lr_b, self.bias_updates, self.biases)
        orc.scal(momentum, self.bias_updates)
        orc.axpy(f32(f32(-decay) * f32(batch)), self.weights, self.weight_updates)
        orc.axpy(lr_b, self.weight_updates, self.weights)
        orc.scal(momentum, self.weight_updates)


class RefSoftmax:
    trainable = False

    def __init__(self, batch, inputs, groups=1):
        self.batch, self.inputs, self.groups = batch, inputs, groups
        self.output = np.zeros((batch, inputs), f32)
        self.delta = np.zeros((batch, inputs), f32)
        self.cost = 0.0

    def forward(self, x, training=False, truth=None):  # nsoftmaxlayer.pas:139-167
        self.output = orc.softmax_batch(np.ascontiguousarray(x, f32).reshape(self.batch, self.inputs), self.groups)
        if truth is not None:
            self.delta, self.loss = orc.softmax_cross_entropy(self.output, np.ascontiguousarray(truth, f32))
            self.cost = orc.vsum(self.loss)
        return self.output

    def backward(self, upstream_delta=None):  # nsoftmaxlayer.pas:169-182
        if upstream_delta is not None:
            upstream_delta += self.delta.reshape(upstream_delta.shape)


class RefConvTrain(RefConvLayer):
    trainable = True


class RefNet:
    """TNNet with lrpCONSTANT policy (nnet.pas:210-273 returns learningRate unchanged for it)."""

    def __init__(self, layers, batch, learning_rate=0.001, momentum=0.9, decay=0.0001, replicas=1):
        self.layers, self.batch, self.replicas = layers, batch, replicas
        self.learning_rate, self.momentum, self.decay = learning_rate, momentum, decay

    def forward(self, x, truth=None, training=True):  # nnet.pas:275-317
        for l in self.layers:
            if training and l.delta is not None:
                l.delta[...] = 0
            x = l.forward(x, training, truth) if isinstance(l, RefSoftmax) else l.forward(x, training)
        return x

    def backward(self):  # nnet.pas:319-376 (state.delta = nil for layer 0)
        for i in range(len(self.layers) - 1, -1, -1):
            self.layers[i].backward(self.layers[i - 1].delta if i else None)

    def update(self):  # nnet.pas:378-403: every layer but the last
        for l in self.layers[:-1]:
            if l.trainable:
                l.update(self.learning_rate, self.momentum, self.decay, self.batch * self.replicas, self.replicas)

    def propagate(self, x, truth):  # nnet.pas:404-432
        self.forward(x, truth, True)
        self.backward()
        return self.layers[-1].cost

    def cost(self):
        return self.layers[-1].cost


def deepcifar10_params(seed=7, c_in=3, size=32, classes=10):
    """Random-init parameters of deepCIFAR10 (nmodels.pas:82-111), shapes only depend on the spec."""
    rng = np.random.default_rng(seed)
    P = []
    c = c_in
    for f in (32, 32, 64, 64, 128, 128):
        s = np.sqrt(2.0 / (9 * c))
        P.append(dict(w=rng.uniform(-s, s, (f, c * 9)).astype(f32), b=np.zeros(f, f32), scales=np.ones(f, f32),
                      rm=np.zeros(f, f32), rv=np.zeros(f, f32)))
        c = f
    n_in = 128 * (size // 8) ** 2
    for n_out in (1024, classes):
        s = np.sqrt(2.0 / n_in)
        P.append(dict(w=rng.uniform(-s, s, (n_out, n_in)).astype(f32), b=np.zeros(n_out, f32)))
        n_in = n_out
    return P


def build_deepcifar10(batch, params, size=32, c_in=3, classes=10, dropout=0.2, **hyper):
    A = orc.ACT
    L = []
    c, h = c_in, size
    pi = 0
    for stage in range(3):
        for _ in range(2):
            p = params[pi]
            pi += 1
            f = p["b"].size
            L.append(RefConvTrain(batch, c, h, h, f, 3, 1, 1, A["RELU"], True, p["w"], p["b"], p["scales"], p["rm"],
                                  p["rv"]))
            c = f
        L.append(RefMaxPool(batch, c, h, h, 2))
        h //= 2
    L.append(RefDropout(batch, c * h * h, dropout))
    p = params[pi]
    L.append(RefConnected(batch, c * h * h, 1024, A["RELU"], p["w"], p["b"]))
    L.append(RefDropout(batch, 1024, dropout))
    p = params[pi + 1]
    L.append(RefConnected(batch, 1024, classes, A["LINEAR"], p["w"], p["b"]))
    L.append(RefSoftmax(batch, classes))
    return RefNet(L, batch, **hyper)


def lenet_mnist_params(seed=5):
    """Random-init parameters of leNetMNIST (nmodels.pas:50-64)."""
    rng = np.random.default_rng(seed)
    P = []
    for c, f in ((1, 6), (6, 16), (16, 120)):
        s = np.sqrt(2.0 / (25 * c))
        P.append(dict(w=rng.uniform(-s, s, (f, c * 25)).astype(f32), b=rng.uniform(-0.1, 0.1, f).astype(f32)))
    for n_in, n_out in ((120, 84), (84, 10)):
        s = np.sqrt(2.0 / n_in)
        P.append(dict(w=rng.uniform(-s, s, (n_out, n_in)).astype(f32), b=rng.uniform(-0.1, 0.1, n_out).astype(f32)))
    return P


def build_lenet_mnist(batch, params, **hyper):
    A = orc.ACT
    cv = lambda p, c, h, pad: RefConvTrain(batch, c, h, h, p["b"].size, 5, 1, pad, A["RELU"], False, p["w"], p["b"])
    L = [cv(params[0], 1, 28, 2), RefMaxPool(batch, 6, 28, 28, 2), cv(params[1], 6, 14, 0), RefMaxPool(batch, 16, 10, 10, 2),
         cv(params[2], 16, 5, 0), RefConnected(batch, 120, 84, A["RELU"], params[3]["w"], params[3]["b"]),
         RefConnected(batch, 84, 10, A["LINEAR"], params[4]["w"], params[4]["b"]), RefSoftmax(batch, 10)]
    return RefNet(L, batch, **hyper)


class RefLSTM:
    """TLSTMLayer.forward (source/nlstmlayer.pas:518-611): eight linear connected layers (wf wi wg wo on the hidden
    state, uf ui ug uo on the step's input; constructor :42-82, `batch := aBatch div aSteps`), gates
    f,i,o = logistic, g = tanh, c = c*f + i*g, h = tanh(c)*o; step t reads input rows [t*batch, (t+1)*batch)."""
    GATES = ("f", "i", "g", "o")

    def __init__(self, batch_total, inputs, outputs, steps, params):
        assert batch_total % steps == 0
        self.batch, self.inputs, self.outputs, self.steps = batch_total // steps, inputs, outputs, steps
        self.w = {g: np.ascontiguousarray(params["w" + g]["w"], f32).reshape(outputs, outputs) for g in self.GATES}
        self.u = {g: np.ascontiguousarray(params["u" + g]["w"], f32).reshape(outputs, inputs) for g in self.GATES}
        self.bw = {g: np.ascontiguousarray(params["w" + g]["b"], f32) for g in self.GATES}
        self.bu = {g: np.ascontiguousarray(params["u" + g]["b"], f32) for g in self.GATES}
        self.output = np.zeros((batch_total, outputs), f32)
        self.cell = np.zeros((batch_total, outputs), f32)
        self.c = np.zeros((self.batch, outputs), f32)
        self.h = np.zeros((self.batch, outputs), f32)
        # training state: the sub-layers' linear outputs per step (wf.output ... uo.output), their update buffers,
        # the layer's delta (filled by the caller before backward) and the carried dc
        self.wout = {g: np.zeros((batch_total, outputs), f32) for g in self.GATES}
        self.uout = {g: np.zeros((batch_total, outputs), f32) for g in self.GATES}
        self.w_updates = {g: np.zeros_like(self.w[g]) for g in self.GATES}
        self.u_updates = {g: np.zeros_like(self.u[g]) for g in self.GATES}
        self.bw_updates = {g: np.zeros_like(self.bw[g]) for g in self.GATES}
        self.bu_updates = {g: np.zeros_like(self.bu[g]) for g in self.GATES}
        self.delta = np.zeros((batch_total, outputs), f32)
        self.dc = np.zeros((self.batch, outputs), f32)

    def _connected(self, x, w, b, n_in):
        out = np.zeros((self.batch, self.outputs), f32)
        orc.sgemm(False, True, self.batch, self.outputs, n_in, 1.0, x, n_in, w, n_in, 0.0, out, self.outputs)
        orc.forward_bias(out, b, self.batch)
        return out

    def forward(self, x, training=False):
        A = orc.ACT
        x = np.ascontiguousarray(x, f32).reshape(self.steps * self.batch, self.inputs)
        B = self.batch
        for t in range(self.steps):
            xt = np.ascontiguousarray(x[t * B:(t + 1) * B])
            gate = {}
            for g in self.GATES:
                wo = self._connected(self.h, self.w[g], self.bw[g], self.outputs)
                uo = self._connected(xt, self.u[g], self.bu[g], self.inputs)
                self.wout[g][t * B:(t + 1) * B] = wo
                self.uout[g][t * B:(t + 1) * B] = uo
                gate[g] = (wo + uo).astype(f32)  # addvv
                orc.activate_array(gate[g], A["TANH"] if g == "g" else A["LOGISTIC"])
            temp = (gate["i"] * gate["g"]).astype(f32)           # mulvv
            self.c = (self.c * gate["f"]).astype(f32)            # _c.Multiply(_f)
            self.c = (self.c + temp).astype(f32)                 # _c.add(temp)
            self.h = self.c.copy()
            orc.activate_array(self.h, A["TANH"])
            self.h = (self.h * gate["o"]).astype(f32)
            self.cell[t * B:(t + 1) * B] = self.c
            self.output[t * B:(t + 1) * B] = self.h
        self.input = x
        return self.output

    def _connected_backward(self, d, inp, w, w_updates, b_updates, upstream, n_in):
        """TConnectedLayer.backward (nconnectedlayer.pas:244-322) of one linear sub-layer at one step: clamp, bias
        sums, weight_updates += d^T inp, upstream += d W."""
        orc.clamp(d, -1.0, 1.0)
        orc.add_sums(b_updates, d, self.batch)
        orc.sgemm(True, False, self.outputs, n_in, self.batch, 1.0, d, self.outputs, inp, n_in, 1.0, w_updates, n_in)
        if upstream is not None:
            orc.sgemm(False, False, self.batch, n_in, self.outputs, 1.0, d, self.outputs, w, n_in, 1.0, upstream, n_in)

    def backward(self, upstream_delta=None, stale_first_step=True):
        """TLSTMLayer.backward (source/nlstmlayer.pas:621-841): back-propagation through time, last step first.  self.delta
        holds dL/dh of every step on entry; the hidden-side sub-layers add W^T d into the PREVIOUS step's delta, the
        input-side ones into `upstream_delta` rows of the step.

        stale_first_step=True restates the reference literally: prevCell / prevState are only refreshed `if i <> 0`
        (:664-669), so at step 0 they still hold what the step-1 iteration copied — cell[0] and output[0] — not the
        state before the sequence.  (darknet's backward_lstm_layer has the same behaviour.)  False gives the textbook
        recurrence (zeros before the sequence), which tests check against finite differences."""
        A = orc.ACT
        B, H = self.batch, self.outputs
        x = self.input
        up = None if upstream_delta is None else upstream_delta.reshape(self.steps * B, self.inputs)
        for t in range(self.steps - 1, -1, -1):
            sl = slice(t * B, (t + 1) * B)
            if t:
                prev_cell, prev_state = self.cell[(t - 1) * B:t * B].copy(), self.output[(t - 1) * B:t * B].copy()
            elif stale_first_step and self.steps > 1:
                prev_cell, prev_state = self.cell[0:B].copy(), self.output[0:B].copy()   # left over from the t = 1 pass
            else:
                prev_cell, prev_state = np.zeros((B, H), f32), np.zeros((B, H), f32)      # fresh tensors are zero
            c = self.cell[sl].copy()
            gate = {}
            for g in self.GATES:
                gate[g] = (self.wout[g][sl] + self.uout[g][sl]).astype(f32)
                orc.activate_array(gate[g], A["TANH"] if g == "g" else A["LOGISTIC"])
            temp3 = self.delta[sl].copy()
            temp = c.copy()
            orc.activate_array(temp, A["TANH"])
            temp2 = (temp3 * gate["o"]).astype(f32)
            orc.gradient_array(temp, A["TANH"], temp2)
            temp2 = (temp2 + self.dc).astype(f32)
            temp = c.copy()
            orc.activate_array(temp, A["TANH"])
            temp = (temp * temp3).astype(f32)
            orc.gradient_array(gate["o"], A["LOGISTIC"], temp)
            dh = self.delta[(t - 1) * B:t * B] if t else None
            xt = np.ascontiguousarray(x[sl])
            upt = None if up is None else up[sl]

            def both(g, d):
                dw = d.copy()
                self._connected_backward(dw, prev_state, self.w[g], self.w_updates[g], self.bw_updates[g], dh, H)
                du = d.copy()
                self._connected_backward(du, xt, self.u[g], self.u_updates[g], self.bu_updates[g], upt, self.inputs)

            both("o", temp)
            temp = (temp2 * gate["i"]).astype(f32)
            orc.gradient_array(gate["g"], A["TANH"], temp)
            both("g", temp)
            temp = (temp2 * gate["g"]).astype(f32)
            orc.gradient_array(gate["i"], A["LOGISTIC"], temp)
            both("i", temp)
            temp = (temp2 * prev_cell).astype(f32)
            orc.gradient_array(gate["f"], A["LOGISTIC"], temp)
            both("f", temp)
            self.dc = (temp2 * gate["f"]).astype(f32)

    def update(self, learning_rate, momentum, decay, batch):
        """TLSTMLayer.update (nlstmlayer.pas:843-861): the eight sub-layers' TConnectedLayer.update (nconnectedlayer.pas:324-359)."""
        lr_b = f32(f32(learning_rate) / f32(batch))
        for g in self.GATES:
            for w, wu, b, bu in ((self.w[g], self.w_updates[g], self.bw[g], self.bw_updates[g]),
                                 (self.u[g], self.u_updates[g], self.bu[g], self.bu_updates[g])):
                orc.axpy(lr_b, bu, b)
                orc.scal(momentum, bu)
                orc.axpy(f32(f32(-decay) * f32(batch)), w, wu)
                orc.axpy(lr_b, wu, w)
                orc.scal(momentum, wu)


def lstm_params(inputs, outputs, seed=0):
    """TConnectedLayer's init for the eight sub-layers (nconnectedlayer.pas:80-81: U(-r, r), r = sqrt(2/inputs));
    biases drawn non-zero so that the parity test exercises them."""
    rng = np.random.default_rng(seed)
    P = {}
    for g in RefLSTM.GATES:
        for kind, n_in in (("w", outputs), ("u", inputs)):
            r = np.sqrt(2.0 / n_in)
            P[kind + g] = dict(w=rng.uniform(-r, r, (outputs, n_in)).astype(f32), b=rng.uniform(-0.1, 0.1, outputs).astype(f32))
    return P
