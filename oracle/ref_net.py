"""CPU forward of the YOLOv3 graph with the oracle's conv path — TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Follows the reference's CPU layer sequencing: TNNet.forward (nnet.pas:275-317), TConvolutionalLayer.forward
(nConvolutionLayer.pas:457-569, after fuseBatchNorm :102-126), TAddLayer (naddlayer.pas), route as a per-image
channel concat (nconcatlayer.pas; the reference's B>1 behaviour is documented in SURVEY.md §7), upsample
(nupsamplelayer.pas), yolo logistic planes (nyololayer.pas:786-815).  Parity unpinned (see tensorium_oracle.c).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc
from .ref_layers import RefConvLayer


def forward(spec, params, x, fuse_bn=True, collect=None):
    """x: [B,3,H,W] float32.  Returns the list of yolo-layer outputs.  `collect`: optional dict index -> output copy."""
    B = x.shape[0]
    outs = []
    yolo = []
    cur = np.ascontiguousarray(x, dtype=np.float32)
    for i, item in enumerate(spec):
        kind = item[0]
        if kind == "conv":
            _, f, k, s, bn, act = item
            p = params[i]
            c, h, w = cur.shape[1:]
            L = RefConvLayer(B, c, h, w, f, k, s, k // 2, orc.ACT["LEAKY"] if act == "leaky" else orc.ACT["LINEAR"], bn,
                             p["w"], p["b"], p.get("scales"), p.get("rm"), p.get("rv"))
            if bn and fuse_bn:
                L.fuse_batchnorm()
            cur = L.forward(cur, training=False)
        elif kind == "shortcut":
            cur = (cur + outs[i + item[1]]).astype(np.float32)
        elif kind == "route":
            idx = [(i + j) if j < 0 else j for j in item[1]]
            cur = np.ascontiguousarray(np.concatenate([outs[j] for j in idx], axis=1))
        elif kind == "upsample":
            cur = np.ascontiguousarray(np.repeat(np.repeat(cur, item[1], axis=2), item[1], axis=3))
        elif kind == "yolo":
            b, c, h, w = cur.shape
            y = cur.reshape(b, 3, 85, h * w).copy()
            for e in (0, 1) + tuple(range(4, 85)):
                plane = np.ascontiguousarray(y[:, :, e, :])
                orc.activate_array(plane, orc.ACT["LOGISTIC"])
                y[:, :, e, :] = plane
            cur = y.reshape(b, c, h, w)
            yolo.append(cur)
        outs.append(cur)
        if collect is not None and i in collect:
            collect[i] = cur.copy()
    return yolo
