"""CPU sequencing of the reference's conv layer, built on the oracle primitives.

TEST INFRASTRUCTURE ONLY (see oracle/tensorium_oracle.c header; parity unpinned).
Follows TConvolutionalLayer.forward/backward/update (nConvolutionLayer.pas:457-710) and
TBaseLayer.batchNorm/batchNormBack (nbaselayer.pas:336-396) call for call.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

BN_MOMENTUM = 0.1  # TBaseLayer.Create nbaselayer.pas:185-188


class RefConvLayer:
    """State of one TConvolutionalLayer (fields as in nConvolutionLayer.pas:140-321)."""

    def __init__(self, batch, c, h, w, filters, ksize, stride, pad, activation, batch_normalize,
                 weights, biases, scales=None, rolling_mean=None, rolling_variance=None, dilation=1):
        self.batch, self.c, self.h, self.w = batch, c, h, w
        self.filters, self.ksize, self.stride, self.pad = filters, ksize, stride, pad
        self.dilation = dilation
        self.activation = activation
        self.bn = bool(batch_normalize)
        self.outH = orc.out_dim(h, pad, ksize, stride, dilation)
        self.outW = orc.out_dim(w, pad, ksize, stride, dilation)
        f32 = np.float32
        self.weights = np.ascontiguousarray(weights, dtype=f32).reshape(filters, -1).copy()
        self.biases = np.ascontiguousarray(biases, dtype=f32).copy()
        self.weight_updates = np.zeros_like(self.weights)
        self.bias_updates = np.zeros_like(self.biases)
        shape = (batch, filters, self.outH, self.outW)
        self.output = np.zeros(shape, dtype=f32)
        self.delta = np.zeros(shape, dtype=f32)
        if self.bn:
            self.scales = np.ascontiguousarray(scales, dtype=f32).copy()
            self.scale_updates = np.zeros_like(self.scales)
            self.rolling_mean = np.ascontiguousarray(rolling_mean, dtype=f32).copy()
            self.rolling_variance = np.ascontiguousarray(rolling_variance, dtype=f32).copy()
            self.mean = np.zeros(filters, dtype=f32)
            self.variance = np.zeros(filters, dtype=f32)
            self.mean_delta = np.zeros(filters, dtype=f32)
            self.variance_delta = np.zeros(filters, dtype=f32)
            self.x = np.zeros(shape, dtype=f32)
            self.x_norm = np.zeros(shape, dtype=f32)

    # TBaseLayer.batchNorm nbaselayer.pas:336-370
    def _batchnorm(self, training):
        out = self.output
        g = self.batch
        if training:
            self.mean, self.variance = orc.means_and_vars(out, g, self.filters)
            orc.scal(1 - BN_MOMENTUM, self.rolling_mean)
            orc.axpy(BN_MOMENTUM, self.mean, self.rolling_mean)
            orc.scal(1 - BN_MOMENTUM, self.rolling_variance)
            orc.axpy(BN_MOMENTUM, self.variance, self.rolling_variance)
            self.x[...] = out
            orc.block_normalize(out, self.mean, self.variance, g)
            self.x_norm[...] = out
        else:
            orc.block_normalize(out, self.rolling_mean, self.rolling_variance, g)
        orc.forward_scale(out, self.scales, g)
        orc.forward_bias(out, self.biases, g)

    # TConvolutionalLayer.forward nConvolutionLayer.pas:457-569
    def forward(self, x, training=False):
        self.input = np.ascontiguousarray(x, dtype=np.float32)
        self.output = orc.conv2d(self.input, self.weights, self.pad, self.stride, self.stride, self.dilation,
                                 self.ksize, self.ksize)
        if self.bn:
            self._batchnorm(training)
        else:
            orc.forward_bias(self.output, self.biases, self.batch)
        orc.activate_array(self.output, self.activation)
        return self.output

    # TBaseLayer.batchNormBack nbaselayer.pas:372-396
    def _batchnorm_back(self):
        g = self.batch
        orc.add_dots(self.scale_updates, self.x_norm, self.delta, g)
        orc.forward_scale(self.delta, self.scales, g)
        self.mean_delta, self.variance_delta = orc.mean_and_variance_delta(self.delta, self.x, self.mean,
                                                                           self.variance, g)
        orc.normalize_delta(self.delta, self.x, self.mean, self.variance, self.mean_delta,
                            self.variance_delta, g)

    # TConvolutionalLayer.backward nConvolutionLayer.pas:571-671.  `delta` (dL/d output, sign as the
    # reference: truth - pred) must be in self.delta; upstream_delta (state.delta) is accumulated into.
    def backward(self, upstream_delta=None):
        k, s, p, d = self.ksize, self.stride, self.pad, self.dilation
        i_m = self.filters
        i_n = k * k * self.c
        i_k = self.outH * self.outW
        orc.gradient_array(self.output, self.activation, self.delta)
        if self.bn:
            self._batchnorm_back()
        else:
            orc.add_sums(self.bias_updates, self.delta, self.batch)
        col = orc.im2col_batched(self.input, k, k, p * d, p * d, s, s, d, d)  # :628
        for b in range(self.batch):  # :633-637  NT, beta=1 accumulates over the batch
            orc.sgemm(False, True, i_m, i_n, i_k, 1.0, self.delta[b], i_k, col[b], i_k, 1.0,
                      self.weight_updates, i_n)
        if upstream_delta is not None:
            ws = np.zeros_like(col)
            orc.sgemm_strided_batched(True, False, i_n, i_k, i_m, 1.0, self.weights, i_n, 0, self.delta, i_k,
                                      i_m * i_k, 0.0, ws, i_k, ws[0].size, self.batch)  # :645-649
            orc.col2im_batched(ws, self.batch, self.c, self.h, self.w, k, k, p * d, p * d, s, s, d, d,
                               im=upstream_delta)  # :653
        return upstream_delta

    # TConvolutionalLayer.update nConvolutionLayer.pas:673-710 (SGD + momentum + L2)
    def update(self, learning_rate, momentum, decay, batch, replicas=1):
        # replicas > 1: data-parallel bookkeeping of tensorium_b200/dataparallel.py (not in the reference)
        momentum = momentum / replicas
        f32 = np.float32
        lr_b = f32(f32(learning_rate) / f32(batch))
        orc.axpy(f32(f32(-decay) * f32(batch)), self.weights, self.weight_updates)
        orc.axpy(lr_b, self.weight_updates, self.weights)
        orc.scal(momentum, self.weight_updates)
        orc.axpy(lr_b, self.bias_updates, self.biases)
        orc.scal(momentum, self.bias_updates)
        if self.bn:
            orc.axpy(lr_b, self.scale_updates, self.scales)
            orc.scal(momentum, self.scale_updates)

    # TBaseConvolutionalLayer.fuseBatchNorm nConvolutionLayer.pas:102-126
    def fuse_batchnorm(self):
        if self.bn:
            orc.fuse_batchnorm(self.weights, self.biases, self.scales, self.rolling_mean, self.rolling_variance)
            self.bn = False
