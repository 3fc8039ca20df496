"""Time every distinct YOLOv3-416 convolution shape (batch 64) on the channels-innermost kernel and on the planar
(NCHW) dispatch, CUDA events on the launching stream.  usage: python tools/nhwc_layers.py [3xtf32|tf32] [out.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import tensor  # noqa: E402
from tensorium_b200.nnet import layer_channels, yolov3_spec  # noqa: E402

MATH = {"3xtf32": 1, "tf32": 2, "bf16x3": 3}


def shapes(size=416):
    spec = yolov3_spec()
    outc = layer_channels(spec)
    hw = size
    hws = []
    seen = {}
    h = size
    hs = []
    # replay the graph's spatial sizes
    cur = size
    sizes = []
    for i, it in enumerate(spec):
        if it[0] == "conv":
            cur_in = cur
            if it[3] == 2:
                cur = cur // 2
            sizes.append((cur_in, cur))
        elif it[0] == "upsample":
            cur *= 2
            sizes.append((cur, cur))
        elif it[0] == "route":
            idx = [(i + j) if j < 0 else j for j in it[1]]
            cur = sizes[idx[0]][1]
            sizes.append((cur, cur))
        else:
            sizes.append((cur, cur))
    out = {}
    for i, it in enumerate(spec):
        if it[0] != "conv":
            continue
        c = outc[i - 1] if i else 3
        key = (c, it[1], it[2], it[3], sizes[i][0])
        out[key] = out.get(key, 0) + 1
    return out


def main():
    math = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
    B = int(os.environ.get("B", "64"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    cu = tensor.initCUDART(0)
    cu.set_stream(stream.cuda_stream)
    cu.set_math_mode(MATH[math])
    rows = []
    tot = {"nhwc": 0.0, "nchw": 0.0}
    for (c, f, k, s, h), count in sorted(shapes().items(), key=lambda kv: -kv[0][4]):
        if c < 16:
            continue
        p = k // 2
        oh = (h + 2 * p - k) // s + 1
        x = torch.rand(B * h * h * c, device="cuda") - 0.5
        w = (torch.rand(f * c * k * k, device="cuda") - 0.5) * float(np.sqrt(2.0 / (k * k * c)))
        b = torch.zeros(f, device="cuda")
        y = torch.empty(B * oh * oh * f, device="cuda")
        desc = cu.conv_desc(B, c, h, h, f, k, k, p, p, s, s)
        cu.setWeightsStatic(w.data_ptr(), True)

        def run_nhwc():
            cu.conv2dForwardNhwc(desc, x.data_ptr(), 0, c, 1, w.data_ptr(), None, b.data_ptr(), 9, y.data_ptr(), 0, f,
                                 1 if f % 4 == 0 else 0)

        def run_nchw():
            cu.conv2dForward(desc, x.data_ptr(), 0, w.data_ptr(), None, b.data_ptr(), 9, y.data_ptr(), 0)

        res = {}
        for name, fn in (("nhwc", run_nhwc), ("nchw", run_nchw)):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / reps
        kern = cu.lastConvKernel()
        cu.setWeightsStatic(w.data_ptr(), False)
        fl = 2.0 * B * f * oh * oh * c * k * k
        floor_hbm = 4.0 * (B * c * h * h + B * f * oh * oh) / 6.45e12 * 1e3
        rows.append(dict(c=c, f=f, k=k, s=s, h=h, count=count, ms_nhwc=res["nhwc"], ms_nchw=res["nchw"],
                         tflops_nhwc=fl / res["nhwc"] / 1e9, tflops_nchw=fl / res["nchw"] / 1e9, hbm_floor_ms=floor_hbm,
                         nchw_kernel=kern))
        tot["nhwc"] += res["nhwc"] * count
        tot["nchw"] += res["nchw"] * count
        print(f"{c:5d}->{f:5d} k{k} s{s} @{h:3d} x{count:2d}  nhwc {res['nhwc']:.3f} ms {fl / res['nhwc'] / 1e9:6.1f} TF/s | "
              f"nchw {res['nchw']:.3f} ms {fl / res['nchw'] / 1e9:6.1f} TF/s  (hbm floor {floor_hbm:.3f})  [{kern}]", flush=True)
        del x, w, y
    print(f"sum over the 74 tensor-core convolutions of a step: nhwc {tot['nhwc']:.2f} ms, nchw {tot['nchw']:.2f} ms ({math})")
    if len(sys.argv) > 2:
        json.dump(dict(math=math, batch=B, total_ms=tot, layers=rows), open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
