"""Model builders of the reference's samples that run on the convolution hot path.

deepCIFAR10 follows source/nmodels.pas:82-111 layer for layer (the net examples/CIFAR10Conv trains):
2x[conv3x3 relu BN] - maxpool2 at 32, 64 and 128 filters, dropout .2, connected 2048->1024 relu,
dropout .2, connected 1024->10 linear, softmax.

leNetMNIST follows source/nmodels.pas:50-64 (the convolutional alternative of examples/MNISTConnected,
MNISTConnected.pas:52): conv5x5 pad 2 (1->6) - maxpool2 - conv5x5 (6->16) - maxpool2 - conv5x5 (16->120) -
connected 120->84 relu - connected 84->10 linear - softmax; ReLU, no batch normalisation, so the biases learn.
"""
from __future__ import annotations

import numpy as np

from . import nncuda as nn
from .layers import TConvolutionalLayer
from .nnet import TNNet
from .train_layers import TConnectedLayer, TDropoutLayer, TMaxPoolLayer, TSoftmaxLayer


def deepCIFAR10(batch=1, size=32, channels=3, classes=10, fused=True, seed=0, params=None, dropout=0.2, **hyper):
    """Returns a TNNet in training mode (CIFAR10Conv.pas:139-170 sets batch, then setTraining(true)).
    `params`: optional list of 8 dicts (6 conv: w,b,scales,rm,rv; 2 connected: w,b) to load instead of the
    seeded random init."""
    rng = np.random.default_rng(seed)
    L = []
    c, h = channels, size
    for f in (32, 64, 128):
        for _ in range(2):
            L.append(TConvolutionalLayer(batch, h, h, c, f, 3, 1, 1, nn.acRELU, True, train=True, fused=fused,
                                         rng=rng, host_output=False))
            c = f
        L.append(TMaxPoolLayer(batch, h, h, c, 2))
        h //= 2
    L.append(TDropoutLayer(batch, dropout, h * h * c, rng=rng))
    L.append(TConnectedLayer(batch, 1, h * h * c, 1024, nn.acRELU, rng=rng))
    L.append(TDropoutLayer(batch, dropout, 1024, rng=rng))
    L.append(TConnectedLayer(batch, 1, 1024, classes, nn.acLINEAR, rng=rng))
    L.append(TSoftmaxLayer(batch, classes))
    _load_params(L, params)
    return TNNet(L, batch, channels, size, size, **hyper)


def _load_params(L, params):
    if params is None:
        return
    trainable = [l for l in L if getattr(l, "trainable", False)]
    assert len(trainable) == len(params)
    for l, p in zip(trainable, params):
        l.weights.Data[...] = p["w"].reshape(l.weights.FShape)
        l.biases.Data[...] = p["b"]
        if "scales" in p:
            l.scales.Data[...] = p["scales"]
            l.rolling_mean.Data[...] = p["rm"]
            l.rolling_variance.Data[...] = p["rv"]


def leNetMNIST(batch=1, fused=True, seed=0, params=None, **hyper):
    """Returns a TNNet in training mode; `params`: optional list of 5 dicts (3 conv, 2 connected: w, b)."""
    rng = np.random.default_rng(seed)
    conv = lambda h, c, f, pad: TConvolutionalLayer(batch, h, h, c, f, 5, 1, pad, nn.acRELU, False, train=True,
                                                    fused=fused, rng=rng, host_output=False)
    L = [conv(28, 1, 6, 2), TMaxPoolLayer(batch, 28, 28, 6, 2), conv(14, 6, 16, 0), TMaxPoolLayer(batch, 10, 10, 16, 2),
         conv(5, 16, 120, 0), TConnectedLayer(batch, 1, 120, 84, nn.acRELU, rng=rng),
         TConnectedLayer(batch, 1, 84, 10, nn.acLINEAR, rng=rng), TSoftmaxLayer(batch, 10)]
    _load_params(L, params)
    return TNNet(L, batch, 1, 28, 28, **hyper)
