"""Host-side mirror of the reference layers that sit on the convolution hot path.

TConvolutionalLayer.forwardGPU / backwardGPU / updateGPU follow source/nConvolutionLayer.pas:1022-1330 and
TBaseLayer.batchNormGPU / batchNormBackGPU follow source/nbaselayer.pas:550-714.  Each has two bodies:

* fused=True  (default): the coarse C-ABI entry points (tnn_conv2d_forward, tnn_batchnorm_*, tnn_sgd_update):
  implicit-GEMM tcgen05 kernel with the bias / batch-norm / activation epilogue, no column buffer.
* fused=False: the reference's own call sequence, op for op (im2col per image, gemmStridedBatched,
  forwardBias | normalize+forwardScale+forwardBias, ActivateArray, ...), i.e. what an UNMODIFIED
  nConvolutionLayer.pas would issue against the replacement nnCuda unit.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import nncuda as nn
from .tensor import TSingleTensor, cuda, split_view

ltCONVOLUTIONAL, ltBATCHNORM = "convolutional", "batchnorm"


@dataclass
class TNNetState:  # source/ntypes.pas:46-63 (fields used on this path)
    input: TSingleTensor | None = None
    delta: TSingleTensor | None = None
    workspace: TSingleTensor | None = None
    isTraining: bool = False
    adversarial: bool = False
    step: int = 0
    inputStep: int = 0
    index: int = 0
    truth: TSingleTensor | None = None


@dataclass
class TUpdateArgs:  # source/ntypes.pas:77-87
    batch: int = 1
    learningRate: float = 0.001
    momentum: float = 0.9
    decay: float = 0.0001
    adam: bool = False
    B1: float = 0.9
    B2: float = 0.999
    eps: float = 1e-8
    t: int = 0
    # data-parallel replicas whose summed updates the buffers hold (1 = the reference's single device):
    # the momentum carried in each replica's buffer is kept as momentum/replicas so that the next
    # all-reduce SUM restores momentum * (previous summed update)
    replicas: int = 1


def conv_out(n, pad, k, stride, dil=1):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


class TConvolutionalLayer:
    """source/nConvolutionLayer.pas:140-321 (constructor fields) and :1022-1330 (CUDA path)."""

    layerType = ltCONVOLUTIONAL
    trainable = True

    def __init__(self, batch, h, w, c, filters, kernelSize, stride=1, padding=0, activation=nn.acLEAKY,
                 batch_normalize=False, dilation=1, train=False, fused=True, rng=None, host_output=True):
        self.batch, self.h, self.w, self.c = batch, h, w, c
        self.filters = self.n = filters
        self.kernelSize, self.stride_x, self.stride_y = kernelSize, stride, stride
        self.Padding, self.Dilation = padding, dilation
        self.ActivationType = activation
        self.isBatchNormalized = bool(batch_normalize)
        self.train = train
        self.fused = fused
        self.learningRateScale = 1.0
        self.bnMomentum = 0.1  # nbaselayer.pas:185-188
        self.outH = conv_out(h, padding, kernelSize, stride, dilation)
        self.outW = conv_out(w, padding, kernelSize, stride, dilation)
        self.outC = filters
        self.outputs = self.outH * self.outW * filters
        self.inputs = h * w * c
        self.nweights = filters * c * kernelSize * kernelSize
        rng = rng or np.random.default_rng(0)
        scale = np.sqrt(2.0 / (kernelSize * kernelSize * c))  # nConvolutionLayer.pas:216-220
        self.weights = TSingleTensor((filters, c * kernelSize * kernelSize), groups=1,
                                     data=rng.uniform(-scale, scale, (filters, c * kernelSize * kernelSize)))
        self.biases = TSingleTensor((filters,), groups=1)
        shape = (batch, filters, self.outH, self.outW)
        self.output = TSingleTensor(shape, groups=batch, host=host_output)
        if train:
            self.weight_updates = TSingleTensor(self.weights.FShape, groups=1)
            self.bias_updates = TSingleTensor((filters,), groups=1)
            self.delta = TSingleTensor(shape, groups=batch)
        if self.isBatchNormalized:
            self.scales = TSingleTensor((filters,), groups=1, data=np.ones(filters))  # :249
            self.rolling_mean = TSingleTensor((filters,), groups=1)
            self.rolling_variance = TSingleTensor((filters,), groups=1)
            if train:
                self.scale_updates = TSingleTensor((filters,), groups=1)
                self.mean = TSingleTensor((filters,), groups=1)
                self.variance = TSingleTensor((filters,), groups=1)
                self.mean_delta = TSingleTensor((filters,), groups=1)
                self.variance_delta = TSingleTensor((filters,), groups=1)
                self.x = TSingleTensor(shape, groups=batch)
                self.x_norm = TSingleTensor(shape, groups=batch)
        self._bn_fold = None
        # inference-time fusion of the shortcut layer that follows (set by TNNet.fuseShortcuts): the epilogue adds
        # `_residual` and writes into the shortcut layer's output instead of this layer's own
        self._residual = None
        self._fused_out = None
        # channels-innermost inference graph (TNNet.setChannelsLast): layouts / pixel pitches of the two activations
        self.in_layout = self.out_layout = nn.TNNCuda.LAYOUT_NCHW
        self.out_pitch = 0   # 0 = dense (filters floats per pixel)
        self._in_nhwc = None  # private channels-innermost copy of a planar input the direct kernel cannot take
        self.in_split = self.out_split = False  # pre-split bf16 planes in / out (TNNet.setSplitActivations)
        self.out_f32 = True
        self.res_split = False  # the folded shortcut's other operand is read from its planes (no fp32 twin exists)
        self._yolo = None       # the yolo layer folded into this (head) convolution's epilogue (TNNet.fuseYoloHeads)
        self._desc = nn.TNNCuda.conv_desc(batch, c, h, w, filters, kernelSize, kernelSize, padding, padding, stride,
                                          stride, dilation, dilation)

    # nConvolutionLayer.pas:128-131
    def getWorkspaceSize(self) -> int:
        return self.batch * self.c * self.outH * self.outW * self.kernelSize * self.kernelSize

    # nConvolutionLayer.pas:102-126 (host-side fold; callers push the tensors afterwards)
    def fuseBatchNorm(self):
        if not self.isBatchNormalized:
            return
        for t in (self.weights, self.biases, self.scales, self.rolling_mean, self.rolling_variance):
            if t.wasGPU():
                t.pullFromDevice()
        eps = np.float32(0.000001)
        pre = (self.scales.Data / np.sqrt(np.maximum(self.rolling_variance.Data, eps))).astype(np.float32)
        self.biases.Data[...] = self.biases.Data - self.rolling_mean.Data * pre
        self.weights.Data[...] = self.weights.Data * pre[:, None]
        self.isBatchNormalized = False
        self.weights.setCPU()
        self.biases.setCPU()

    # ---------------------------------------------------------------- forward
    def forwardGPU(self, state: TNNetState, net=None):
        cu = cuda()
        if not state.input.wasGPU():
            state.input.pushToDevice()
        if not self.weights.wasGPU():
            self.weights.pushToDevice()
            # inference: the weights do not change between calls, the backend keeps its prepared copy
            # (declaring again after a push invalidates the copy made from the old contents)
            if self.fused and not self.train:
                cu.setWeightsStatic(self.weights.devData, True)
        if not self.biases.wasGPU():
            self.biases.pushToDevice()
        self.output.setCUDA()
        if self.fused:
            return self._forward_fused(cu, state)
        return self._forward_reference_sequence(cu, state)

    def _forward_fused(self, cu, state):
        self._forward_fused_body(cu, state)
        self._last_kernel = cu.lastConvKernel()  # which kernel the backend dispatched to (see fuseShortcuts)

    def _forward_fused_body(self, cu, state):
        training_bn = self.isBatchNormalized and state.isTraining and not state.adversarial
        if self.isBatchNormalized:
            for t in (self.scales, self.rolling_mean, self.rolling_variance):
                if not t.wasGPU():
                    t.pushToDevice()
        NHWC = nn.TNNCuda.LAYOUT_NHWC
        if self.in_layout == NHWC or self.out_layout == NHWC:
            assert not state.isTraining, "the channels-innermost path is an inference graph"
            sc, bi = self._folded_bn() if self.isBatchNormalized else (None, self.biases)
            src, in_layout, in_pitch = state.input, self.in_layout, getattr(state.input, "pitch", 0) or self.c
            if in_layout != NHWC and self.c * self.kernelSize * self.kernelSize > 64:
                # planar input with a contraction the direct kernel does not take: convert once per call
                if self._in_nhwc is None:
                    pitch = (self.c + 3) // 4 * 4
                    self._in_nhwc = TSingleTensor((self.batch, self.h, self.w, pitch), groups=self.batch, host=False)
                    self._in_nhwc.pitch = pitch
                cu.nchwToNhwc(self.batch, self.c, self.h * self.w, src.devData, self._in_nhwc.devData, self._in_nhwc.pitch)
                src, in_layout, in_pitch = self._in_nhwc, NHWC, self._in_nhwc.pitch
            if self._yolo is not None and in_layout == NHWC:
                # the yolo layer that follows, folded into this head's epilogue (TNNet.fuseYoloHeads): writes ITS output
                yl = self._yolo
                yl.output.setCUDA()
                if self.in_split:
                    in_ptr, pitch, plane = split_view(src)
                    lay = nn.TNNCuda.LAYOUT_NHWC_SPLIT
                else:
                    in_ptr, lay, pitch, plane = src.devData, NHWC, in_pitch, 0
                cu.conv2dForwardYoloHead(self._desc, in_ptr, 0, pitch, lay, plane, self.weights.devData,
                                         sc.devData if sc is not None else None, bi.devData, yl.output.devData, 0,
                                         yl.classes, yl.scaleXY)
                return
            if self.in_split or self.out_split:
                # pre-split (bf16 hi / lo planes) activations in and / or out (TNNet.setSplitActivations; BF16x3 mode)
                SPLIT = nn.TNNCuda.LAYOUT_NHWC_SPLIT
                tgt = self._fused_out if self._residual is not None else self.output
                tgt.setCUDA()
                if self.in_split:
                    in_ptr, in_pitch, in_plane = split_view(src)
                    in_layout = SPLIT
                else:
                    in_ptr, in_plane = src.devData, 0
                res = self._residual
                if res is None:
                    res_ptr, res_pitch, res_layout, res_plane = None, 0, NHWC, 0
                elif self.res_split:   # the shortcut's other operand exists only as planes: residual = hi + lo
                    res_ptr, res_pitch, res_plane = split_view(res)
                    res_layout = SPLIT
                else:
                    res_ptr, res_pitch, res_layout, res_plane = res.devData, getattr(res, "pitch", 0) or self.filters, NHWC, 0
                out_ptr, out_pitch, out_plane = split_view(tgt) if self.out_split else (None, self.filters, tgt.Size())
                cu.conv2dForwardSplitResidual(self._desc, in_ptr, 0, in_pitch, in_layout, in_plane, self.weights.devData,
                                              sc.devData if sc is not None else None, bi.devData, self.ActivationType,
                                              res_ptr, 0, res_pitch, res_layout, res_plane,
                                              tgt.devData if self.out_f32 else None, 0, self.out_pitch or self.filters,
                                              self.out_layout, out_ptr, 0, out_pitch, out_plane)
                return
            if self._residual is not None and in_layout == NHWC and self.out_layout == NHWC:
                # the shortcut that follows, folded into the epilogue (TNNet.fuseShortcuts): writes the shortcut's output
                self._fused_out.setCUDA()
                cu.conv2dForwardNhwcResidual(self._desc, src.devData, 0, in_pitch, in_layout, self.weights.devData,
                                             sc.devData if sc is not None else None, bi.devData, self.ActivationType,
                                             self._residual.devData, 0, getattr(self._residual, "pitch", 0) or self.filters,
                                             self._fused_out.devData, 0, self.filters, self.out_layout)
                return
            cu.conv2dForwardNhwc(self._desc, src.devData, 0, in_pitch, in_layout, self.weights.devData,
                                 sc.devData if sc is not None else None, bi.devData, self.ActivationType,
                                 self.output.devData, 0, self.out_pitch or self.filters, self.out_layout)
            return
        if self._residual is not None and not state.isTraining:
            sc = None
            bi = self.biases
            if self.isBatchNormalized:
                sc, bi = self._folded_bn()
            self._fused_out.setCUDA()
            cu.conv2dForwardResidual(self._desc, state.input.devData, 0, self.weights.devData,
                                     sc.devData if sc is not None else None, bi.devData, self.ActivationType,
                                     self._residual.devData, 0, self._fused_out.devData, 0)
        elif not self.isBatchNormalized:
            cu.conv2dForward(self._desc, state.input.devData, 0, self.weights.devData, None, self.biases.devData,
                             self.ActivationType, self.output.devData, 0)
        elif training_bn and getattr(self, "sync_stats", None) is not None:
            # data-parallel replicas with SyncBN: statistics of the GLOBAL mini-batch (dataparallel.attach(sync_bn=True))
            ss = self.sync_stats
            cu.conv2dForward(self._desc, state.input.devData, 0, self.weights.devData, None, None, nn.acLINEAR,
                             self.output.devData, 0)
            cu.batchNormForwardTrainingSync(ss.comm, self.batch * ss.world, self.batch, self.filters,
                                            self.outH * self.outW, self.output.devData, self.x.devData,
                                            self.x_norm.devData, self.mean.devData, self.variance.devData,
                                            self.rolling_mean.devData, self.rolling_variance.devData, self.bnMomentum,
                                            self.scales.devData, self.biases.devData, self.ActivationType)
            for t in (self.x, self.x_norm, self.mean, self.variance, self.rolling_mean, self.rolling_variance):
                t.setCUDA()
            self._bn_fold = None
        elif training_bn:
            # raw conv output, then the fused training batch-norm (+ activation)
            cu.conv2dForward(self._desc, state.input.devData, 0, self.weights.devData, None, None, nn.acLINEAR,
                             self.output.devData, 0)
            cu.batchNormForwardTraining(self.batch, self.filters, self.outH * self.outW, self.output.devData,
                                        self.x.devData, self.x_norm.devData, self.mean.devData,
                                        self.variance.devData, self.rolling_mean.devData,
                                        self.rolling_variance.devData, self.bnMomentum, self.scales.devData,
                                        self.biases.devData, self.ActivationType)
            for t in (self.x, self.x_norm, self.mean, self.variance, self.rolling_mean, self.rolling_variance):
                t.setCUDA()
            self._bn_fold = None  # rolling statistics changed
        else:
            # inference batch-norm folded into the conv epilogue: out = act(conv*s + b'),
            # s = scale/max(sqrt(var),eps), b' = bias - mean*s
            sc, bi = self._folded_bn()
            cu.conv2dForward(self._desc, state.input.devData, 0, self.weights.devData, sc.devData, bi.devData,
                             self.ActivationType, self.output.devData, 0)

    def _folded_bn(self):
        if self._bn_fold is None:
            eps = np.float32(0.000001)
            for t in (self.scales, self.rolling_mean, self.rolling_variance, self.biases):
                if t.wasGPU():
                    t.pullFromDevice()
                    t.setCUDA()  # the device copy is still the current one: a later forward must not re-push the host copy
            s = (self.scales.Data / np.maximum(np.sqrt(self.rolling_variance.Data), eps)).astype(np.float32)
            b = (self.biases.Data - self.rolling_mean.Data * s).astype(np.float32)
            self._bn_fold = (TSingleTensor(s.shape, 1, s), TSingleTensor(b.shape, 1, b))
            self._bn_fold[0].pushToDevice()
            self._bn_fold[1].pushToDevice()
        return self._bn_fold

    def _forward_reference_sequence(self, cu, state):
        # nConvolutionLayer.pas:1040-1122, call for call
        kSize = self.kernelSize * self.kernelSize
        outImgSize = self.output.area()
        k = self.c * kSize
        imColSize = self.c * kSize * outImgSize
        strideC = outImgSize * self.filters
        if kSize != 1 or self.stride_y * self.stride_x != 1 or self.Dilation * self.Dilation != 1:
            strideA = state.input.volume()
            strideB = imColSize
            _B = state.workspace.devData
            for b in range(self.batch):
                cu.im2col(self.c, self.h, self.w, self.kernelSize, self.kernelSize, self.Padding, self.Padding,
                          self.stride_y, self.stride_x, self.Dilation, self.Dilation, state.input.devData,
                          b * strideA, _B, b * strideB)
        else:
            strideB = state.input.volume()
            _B = state.input.devData
        cu.gemmStridedBatched(False, False, self.filters, outImgSize, k, 1.0, self.weights.devData, 0, k, 0, _B, 0,
                              outImgSize, strideB, 0.0, self.output.devData, 0, outImgSize, strideC, self.batch)
        if self.isBatchNormalized:
            self.batchNormGPU(state)
        else:
            cu.forwardBias(self.output.Size(), self.output.devData, 0, self.biases.Size(), self.biases.devData, 1,
                           self.batch)
        cu.ActivateArray(self.output.Size(), self.output.devData, 0, self.ActivationType)

    # nbaselayer.pas:550-660, call for call
    def batchNormGPU(self, state):
        cu = cuda()
        for t in (self.biases, self.scales, self.rolling_mean, self.rolling_variance):
            if not t.wasGPU():
                t.pushToDevice()
        outputStep = self.batch * self.outputs
        offset = state.step * outputStep
        out = self.output
        if state.isTraining and not state.adversarial:
            cu.means(outputStep, self.mean.Size(), out.Groups, out.devData, offset, self.mean.devData)
            cu.variances(outputStep, self.mean.Size(), out.Groups, out.devData, offset, self.mean.devData,
                         self.variance.devData)
            cu.scale(self.rolling_mean.Size(), 1 - self.bnMomentum, self.rolling_mean.devData, 1)
            cu.axpy(self.rolling_mean.Size(), self.bnMomentum, self.mean.devData, 0, 1, self.rolling_mean.devData, 0, 1)
            cu.scale(self.rolling_variance.Size(), 1 - self.bnMomentum, self.rolling_variance.devData, 1)
            cu.axpy(self.rolling_variance.Size(), self.bnMomentum, self.variance.devData, 0, 1,
                    self.rolling_variance.devData, 0, 1)
            cu.copy(outputStep, out.devData, offset, 1, self.x.devData, offset, 1)
            cu.normalize(self.mean.Size(), outputStep, out.Groups, self.mean.devData, 1, self.variance.devData, 1,
                         out.devData, offset)
            cu.copy(outputStep, out.devData, offset, 1, self.x_norm.devData, offset, 1)
            for t in (self.x, self.x_norm, self.mean, self.variance, self.rolling_mean, self.rolling_variance):
                t.setCUDA()
        else:
            cu.normalize(self.rolling_mean.Size(), outputStep, out.Groups, self.rolling_mean.devData, 1,
                         self.rolling_variance.devData, 1, out.devData, 0)
        cu.forwardScale(outputStep, out.devData, offset, self.scales.Size(), self.scales.devData, 1, out.Groups)
        cu.forwardBias(outputStep, out.devData, offset, self.biases.Size(), self.biases.devData, 1, out.Groups)

    # ---------------------------------------------------------------- backward
    def backwardGPU(self, state: TNNetState):
        cu = cuda()
        for t in (self.delta, self.bias_updates, self.weight_updates):
            if not t.wasGPU():
                t.pushToDevice()
        if not state.input.wasGPU():
            state.input.pushToDevice()
        if not self.weights.wasGPU():
            self.weights.pushToDevice()
        if self.fused:
            self._backward_fused(cu, state)
        else:
            self._backward_reference_sequence(cu, state)
        for t in (self.delta, self.bias_updates, self.weight_updates):
            t.setCUDA()
        if state.delta is not None:
            state.delta.setCUDA()

    def _backward_fused(self, cu, state):
        N = self.output.Size()
        if self.isBatchNormalized:
            if not self.scale_updates.wasGPU():
                self.scale_updates.pushToDevice()
            ss = getattr(self, "sync_stats", None)
            if ss is not None:
                cu.batchNormBackwardSync(ss.comm, self.batch * ss.world, self.batch, self.filters, self.outH * self.outW,
                                         self.output.devData, self.ActivationType, self.delta.devData, self.x.devData,
                                         self.x_norm.devData, self.mean.devData, self.variance.devData,
                                         self.scales.devData, self.scale_updates.devData, self.mean_delta.devData,
                                         self.variance_delta.devData)
            else:
                cu.batchNormBackward(self.batch, self.filters, self.outH * self.outW, self.output.devData,
                                 self.ActivationType, self.delta.devData, self.x.devData, self.x_norm.devData,
                                 self.mean.devData, self.variance.devData, self.scales.devData,
                                 self.scale_updates.devData, self.mean_delta.devData, self.variance_delta.devData)
            self.scale_updates.setCUDA()
        else:
            cu.DeriveArray(N, self.output.devData, 0, self.ActivationType, self.delta.devData)
            cu.backwardBias(self.bias_updates.Size(), self.bias_updates.devData, self.delta.Size(),
                            self.delta.devData, 0, 1, self.batch)
        cu.conv2dBackwardFilter(self._desc, state.input.devData, 0, self.delta.devData, 0,
                                self.weight_updates.devData)
        if state.delta is not None and state.delta.devData:
            cu.conv2dBackwardData(self._desc, self.weights.devData, self.delta.devData, 0, state.delta.devData, 0)

    def _backward_reference_sequence(self, cu, state):
        # nConvolutionLayer.pas:1155-1292, call for call
        i_m = self.filters
        i_n = self.kernelSize * self.kernelSize * self.c
        i_k = self.outH * self.outW
        colSize = self.c * self.outH * self.outW * self.kernelSize * self.kernelSize
        _vol = state.input.volume()
        cu.DeriveArray(self.output.Size(), self.output.devData, 0, self.ActivationType, self.delta.devData)
        if self.isBatchNormalized:
            self.batchNormBackGPU(state)
        else:
            cu.backwardBias(self.bias_updates.Size(), self.bias_updates.devData, self.delta.Size(),
                            self.delta.devData, 0, 1, self.batch)
        ws = state.workspace
        for b in range(self.batch):
            cu.im2col(self.c, self.h, self.w, self.kernelSize, self.kernelSize, self.Padding, self.Padding,
                      self.stride_y, self.stride_x, self.Dilation, self.Dilation, state.input.devData, b * _vol,
                      ws.devData, b * colSize)
        cu.gemmStridedBatched(False, True, i_m, i_n, i_k, 1.0, self.delta.devData, 0, i_k, i_m * i_k, ws.devData, 0,
                              i_k, colSize, 1.0, self.weight_updates.devData, 0, i_n, 0, self.batch)
        if state.delta is not None and state.delta.devData:
            cu.gemmStridedBatched(True, False, i_n, i_k, i_m, 1.0, self.weights.devData, 0, i_n, 0,
                                  self.delta.devData, 0, i_k, i_m * i_k, 0.0, ws.devData, 0, i_k, colSize, self.batch)
            imSize = state.delta.volume()
            for b in range(self.batch):
                cu.col2im(self.c, self.h, self.w, self.kernelSize, self.kernelSize, self.Padding, self.Padding,
                          self.stride_y, self.stride_x, self.Dilation, self.Dilation, ws.devData, b * colSize,
                          state.delta.devData, b * imSize)

    # nbaselayer.pas:662-714, call for call
    def batchNormBackGPU(self, state):
        cu = cuda()
        outputStep = self.batch * self.outputs
        offset = state.step * outputStep
        if not self.scale_updates.wasGPU():
            self.scale_updates.pushToDevice()
        cu.addDots(outputStep, self.scale_updates.Size(), self.delta.Groups, self.x_norm.devData, self.delta.devData,
                   offset, self.scale_updates.devData)
        cu.forwardScale(outputStep, self.delta.devData, offset, self.scales.Size(), self.scales.devData, 1,
                        self.delta.Groups)
        cu.meansAndVarsDelta(outputStep, self.mean.Size(), self.delta.Groups, self.delta.devData, self.x.devData,
                             offset, self.mean.devData, self.variance.devData, self.mean_delta.devData,
                             self.variance_delta.devData)
        cu.normalizeDelta(outputStep, self.mean.Size(), self.delta.Groups, self.delta.devData, self.x.devData, offset,
                          self.mean.devData, self.variance.devData, self.mean_delta.devData,
                          self.variance_delta.devData)
        self.scale_updates.setCUDA()

    # ---------------------------------------------------------------- update (nConvolutionLayer.pas:1294-1330)
    def updateGPU(self, args: TUpdateArgs):
        cu = cuda()
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            if not t.wasGPU():
                t.pushToDevice()
        f32 = np.float32
        learning_rate = f32(args.learningRate) * f32(self.learningRateScale)
        rate = float(f32(learning_rate / f32(args.batch)))
        decay_term = float(f32(f32(-args.decay) * f32(args.batch)))
        if self.fused:
            keep = args.momentum / args.replicas
            cu.sgdUpdate(self.biases.Size(), self.biases.devData, self.bias_updates.devData, rate, keep, 0.0)
            cu.sgdUpdate(self.weights.Size(), self.weights.devData, self.weight_updates.devData, rate, keep,
                         decay_term)
            if self.isBatchNormalized:
                cu.sgdUpdate(self.scales.Size(), self.scales.devData, self.scale_updates.devData, rate, keep, 0.0)
        else:
            assert args.replicas == 1, "the reference call sequence is single-device"
            cu.axpy(self.biases.Size(), rate, self.bias_updates.devData, 0, 1, self.biases.devData, 0, 1)
            cu.scale(self.bias_updates.Size(), args.momentum, self.bias_updates.devData, 1)
            cu.axpy(self.weight_updates.Size(), decay_term, self.weights.devData, 0, 1, self.weight_updates.devData,
                    0, 1)
            cu.axpy(self.weights.Size(), rate, self.weight_updates.devData, 0, 1, self.weights.devData, 0, 1)
            cu.scale(self.weight_updates.Size(), args.momentum, self.weight_updates.devData, 1)
            if self.isBatchNormalized:
                cu.axpy(self.scales.Size(), rate, self.scale_updates.devData, 0, 1, self.scales.devData, 0, 1)
                cu.scale(self.scale_updates.Size(), args.momentum, self.scale_updates.devData, 1)
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            t.setCUDA()
        self._bn_fold = None  # biases / scales changed: the inference fold is stale

    def sgdItems(self, args: TUpdateArgs):
        """The tensors updateGPU touches, for ONE multi-tensor launch over the whole net (TNNet._update_layers): the same
        per-tensor rate / decay term / arithmetic as the fused branch of updateGPU."""
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            if not t.wasGPU():
                t.pushToDevice()
        f32 = np.float32
        rate = float(f32(f32(args.learningRate) * f32(self.learningRateScale) / f32(args.batch)))
        decay_term = float(f32(f32(-args.decay) * f32(args.batch)))
        items = [(self.biases.devData, self.bias_updates.devData, self.biases.Size(), rate, 0.0),
                 (self.weights.devData, self.weight_updates.devData, self.weights.Size(), rate, decay_term)]
        if self.isBatchNormalized:
            items.append((self.scales.devData, self.scale_updates.devData, self.scales.Size(), rate, 0.0))
        for t in (self.biases, self.bias_updates, self.weights, self.weight_updates):
            t.setCUDA()
        self._bn_fold = None
        return items

    def gradients(self):
        """The update buffers a data-parallel run sums across replicas before updateGPU."""
        g = [self.weight_updates, self.bias_updates]
        if self.isBatchNormalized:
            g.append(self.scale_updates)
        return g
