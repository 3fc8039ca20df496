"""Host-side mirror of the layers either side of the convolution on the CIFAR-10 training path
(the sample net is deepCIFAR10, source/nmodels.pas:82-111):

  TMaxPoolLayer    source/nMaxPoolLayer.pas:59-181 (constructor), :656-714 (CUDA path)
  TDropoutLayer    source/ndropoutlayer.pas:37-66, :156-175
  TConnectedLayer  source/nconnectedlayer.pas:62-155, :612-915
  TSoftmaxLayer    source/nsoftmaxlayer.pas:49-65, :279-340

Every body issues the reference's own `cuda.<op>` sequence against the replacement backend; only the
blocking `cuda.finish()` after each pooling call and the per-step host read of the loss tensor are
dropped (the loss is read when `cost` is asked for).
"""
from __future__ import annotations

import numpy as np

from . import nncuda as nn
from .layers import TNNetState, TUpdateArgs
from .tensor import TSingleTensor, cuda

ltMAXPOOL, ltDROPOUT, ltCONNECTED, ltSOFTMAX = "maxpool", "dropout", "connected", "softmax"


class TMaxPoolLayer:
    layerType = ltMAXPOOL
    trainable = False

    def __init__(self, batch, h, w, c, kernelSize, stride_x=0, stride_y=0, padding=0, train=True):
        self.batch, self.h, self.w, self.c = batch, h, w, c
        self.kernelSize = kernelSize
        self.stride_x = stride_x or kernelSize  # nMaxPoolLayer.pas:78-81
        self.stride_y = stride_y or kernelSize
        self.Padding = padding
        self.outC = c
        self.outH = (h + padding - kernelSize) // self.stride_y + 1  # :106-107
        self.outW = (w + padding - kernelSize) // self.stride_x + 1
        self.outputs, self.inputs = self.outH * self.outW * c, h * w * c
        self.train = train
        shape = (batch, c, self.outH, self.outW)
        self.output = TSingleTensor(shape, groups=batch, host=False)
        self.delta = self.indexes = None
        if train:
            self.indexes = TSingleTensor(shape, groups=batch, host=False, dtype=np.int64)  # TSizeIntTensor
            self.delta = TSingleTensor(shape, groups=batch, host=False)

    def forwardGPU(self, state: TNNetState, net=None):
        if not state.input.wasGPU():
            state.input.pushToDevice()
        self.output.setCUDA()
        cuda().forwardMaxPool(self.batch, self.outC, self.outH, self.outW, state.input.devData, self.c, self.h,
                              self.w, self.stride_x, self.stride_y, self.Padding, self.kernelSize,
                              self.indexes.devData if self.indexes is not None else None, self.output.devData)

    def backwardGPU(self, state: TNNetState):
        if state.delta is None or not state.delta.devData:
            return
        cuda().backwardMaxPool(self.batch, self.outC, self.outH, self.outW, state.delta.devData,
                               self.indexes.devData, self.delta.devData)


class TDropoutLayer:
    """Works in place on the previous layer's output and delta (ndropoutlayer.pas:159, :172)."""
    layerType = ltDROPOUT
    trainable = False

    def __init__(self, batch, probability, inputs, rng=None):
        assert probability < 1
        self.batch, self.inputs, self.outputs = batch, inputs, inputs
        self.probability = float(probability)
        self.scale = float(np.float32(1.0 / (1.0 - probability)))
        self.rand = TSingleTensor((batch, inputs), groups=batch, host=False)
        self.output = self.delta = None
        self.train = True
        self._rng = rng or np.random.default_rng(0)
        self.seed = None  # tests pin it; otherwise drawn per call like the wrapper's random($10000000)
        # graph replays cannot take a new host-side seed per step: (device counter, offset) -> seed = *counter + offset
        self.seed_dev = None

    def forwardGPU(self, state: TNNetState, net=None):
        if not state.input.wasGPU():
            state.input.pushToDevice()
        self.output = state.input
        if not state.isTraining:
            return
        if self.seed_dev is not None:
            ctr, off = self.seed_dev
            cuda().forwardDropoutDev(self.output.Size(), self.output.devData, self.probability, self.scale,
                                     self.rand.devData, self.output.devData, ctr, off)
            return
        seed = self.seed if self.seed is not None else int(self._rng.integers(0, 0x10000000))
        self.last_seed = seed
        cuda().forwardDropout(self.output.Size(), self.output.devData, self.probability, self.scale,
                              self.rand.devData, self.output.devData, seed)

    def backwardGPU(self, state: TNNetState):
        if state.delta is None or not state.delta.devData:
            return
        cuda().backwardDropout(state.delta.Size(), state.delta.devData, self.probability, self.scale,
                               self.rand.devData, state.delta.devData)
        self.delta = state.delta


class TConnectedLayer:
    layerType = ltCONNECTED
    trainable = True

    def __init__(self, batch, steps, inputs, outputs, activation=nn.acLINEAR, train=True, rng=None):
        self.batch, self.inputs, self.outputs, self.steps = batch, inputs, outputs, steps
        assert steps == 1 or not train, "backward through time is not built (forward stepping only, rnn_layers.py)"
        self.ActivationType = activation
        self.train = train
        self.learningRateScale = 1.0
        rng = rng or np.random.default_rng(0)
        r = np.sqrt(2.0 / inputs)  # nconnectedlayer.pas:80-81
        self.weights = TSingleTensor((outputs, inputs), groups=1, data=rng.uniform(-r, r, (outputs, inputs)))
        self.biases = TSingleTensor((outputs,), groups=1)
        self.output = TSingleTensor((batch * steps, outputs), groups=batch * steps, host=False)
        if train:
            self.delta = TSingleTensor((batch, outputs), groups=batch, host=False)
            self.weight_updates = TSingleTensor((outputs, inputs), groups=1)
            self.bias_updates = TSingleTensor((outputs,), groups=1)

    def _push(self, *ts):
        for t in ts:
            if not t.wasGPU():
                t.pushToDevice()

    # nconnectedlayer.pas:612-743
    def forwardGPU(self, state: TNNetState, net=None):
        cu = cuda()
        self._push(state.input, self.weights, self.biases)
        self.output.setCUDA()
        outputStep = self.batch * self.outputs
        offset = getattr(state, "step", 0) * outputStep if self.steps > 1 else 0  # nconnectedlayer.pas:620-622
        in_off = getattr(state, "inputStep", 0) * self.batch * self.inputs if self.steps > 1 else 0
        cu.gemm(False, True, self.batch, self.outputs, self.inputs, 1.0, state.input.devData, in_off, self.inputs,
                self.weights.devData, 0, self.inputs, 0.0, self.output.devData, offset, self.outputs)
        cu.forwardBias(outputStep, self.output.devData, offset, self.biases.Size(), self.biases.devData, 1, self.batch)
        cu.ActivateArray(outputStep, self.output.devData, offset, self.ActivationType)

    # nconnectedlayer.pas:745-871
    def backwardGPU(self, state: TNNetState):
        cu = cuda()
        self._push(self.delta, self.weight_updates, self.bias_updates)
        outStepSize = self.batch * self.outputs
        cu.clamp(self.delta.Size(), 1.0, self.delta.devData, self.delta.devData)
        cu.DeriveArray(outStepSize, self.output.devData, 0, self.ActivationType, self.delta.devData)
        cu.backwardBias(self.bias_updates.Size(), self.bias_updates.devData, outStepSize, self.delta.devData, 0, 1,
                        self.delta.Groups)
        cu.gemm(True, False, self.outputs, self.inputs, self.batch, 1.0, self.delta.devData, 0, self.outputs,
                state.input.devData, 0, self.inputs, 1.0, self.weight_updates.devData, 0, self.inputs)
        if state.delta is not None and state.delta.devData:
            cu.gemm(False, False, self.batch, self.inputs, self.outputs, 1.0, self.delta.devData, 0, self.outputs,
                    self.weights.devData, 0, self.inputs, 1.0, state.delta.devData, 0, self.inputs)
        for t in (self.delta, self.weight_updates, self.bias_updates):
            t.setCUDA()

    # nconnectedlayer.pas:873-915
    def updateGPU(self, args: TUpdateArgs):
        cu = cuda()
        self._push(self.biases, self.bias_updates, self.weights, self.weight_updates)
        f32 = np.float32
        rate = float(f32(f32(args.learningRate) * f32(self.learningRateScale) / f32(args.batch)))
        decay_term = float(f32(f32(-args.decay) * f32(args.batch)))
        keep = args.momentum / getattr(args, "replicas", 1)
        cu.sgdUpdate(self.biases.Size(), self.biases.devData, self.bias_updates.devData, rate, keep, 0.0)
        cu.sgdUpdate(self.weights.Size(), self.weights.devData, self.weight_updates.devData, rate, keep, decay_term)
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            t.setCUDA()

    def sgdItems(self, args: TUpdateArgs):
        """See TConvolutionalLayer.sgdItems."""
        self._push(self.biases, self.bias_updates, self.weights, self.weight_updates)
        f32 = np.float32
        rate = float(f32(f32(args.learningRate) * f32(self.learningRateScale) / f32(args.batch)))
        decay_term = float(f32(f32(-args.decay) * f32(args.batch)))
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            t.setCUDA()
        return [(self.biases.devData, self.bias_updates.devData, self.biases.Size(), rate, 0.0),
                (self.weights.devData, self.weight_updates.devData, self.weights.Size(), rate, decay_term)]

    def gradients(self):
        return [self.weight_updates, self.bias_updates]


class TSoftmaxLayer:
    layerType = ltSOFTMAX
    trainable = False

    def __init__(self, batch, inputs, groups=1):
        assert inputs % groups == 0
        self.batch, self.inputs, self.outputs, self.groups = batch, inputs, inputs, groups
        self.temperature = 1.0
        self.train = True
        self.noloss = False
        self.loss = TSingleTensor((batch, inputs), groups=batch)
        self.output = TSingleTensor((batch, inputs), groups=batch)
        self.delta = TSingleTensor((batch, inputs), groups=batch)
        self._cost_stale = False
        self._cost = 0.0

    # nsoftmaxlayer.pas:279-323
    def forwardGPU(self, state: TNNetState, net=None):
        cu = cuda()
        for t in (self.output, self.loss, self.delta):
            t.setCUDA()
        n = self.inputs // self.groups
        cu.softmaxBatch(n, state.input.devData, 0, self.batch, self.inputs, self.groups, n, 1, self.temperature,
                        self.output.devData, 0)
        truth = getattr(state, "truth", None)
        if truth is not None and not self.noloss:
            if not truth.wasGPU():
                truth.pushToDevice()
            cu.crossEntropySoftmax(self.output.Size(), self.output.devData, truth.devData, self.delta.devData,
                                   self.loss.devData)
            self._cost_stale = True

    @property
    def cost(self):
        """cost[0] := loss.Sum() (nsoftmaxlayer.pas:318-319), read back lazily."""
        if self._cost_stale:
            self._cost = float(self.loss.pullFromDevice().sum(dtype=np.float32))
            self.loss.setCUDA()
            self._cost_stale = False
        return self._cost

    # nsoftmaxlayer.pas:325-336
    def backwardGPU(self, state: TNNetState):
        if state.delta is None or not state.delta.devData:
            return
        cuda().addvv(self.delta.Size(), self.delta.devData, 0, 1, state.delta.devData, 0, 1, state.delta.devData, 0,
                     1)
