"""Run single conv configurations through tnn_conv2d_forward (for ncu / timing).  usage: conv_one.py [math] [cfg ...]
cfg = B,C,H,F,k,s      TNN_NHWC=1: the channels-innermost entry point (tnn_conv2d_forward_nhwc; random data, so the planar
input buffer is simply read as rows of channels)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev

MATH = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16x3": 3}
NHWC = os.environ.get("TNN_NHWC", "0") == "1"
SPLIT = os.environ.get("TNN_SPLIT", "0") in ("1", "2")   # pre-split bf16 planes in and out (tnn_conv2d_forward_split, bf16x3 only)
SPLIT_RES = os.environ.get("TNN_SPLIT", "0") == "2"       # ... plus the folded shortcut's operand read from planes
math = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(64, 256, 26, 512, 3, 1)]
cu = tensor.initCUDART(0)
cu.set_math_mode(MATH[math])
rng = np.random.default_rng(0)
for (B, C, H, F, k, s) in cfgs:
    p = k // 2
    oh = (H + 2 * p - k) // s + 1
    x = rng.random((B, C, H, H), dtype=np.float32)
    w = rng.uniform(-0.05, 0.05, (F, C * k * k)).astype(np.float32)
    b = np.zeros(F, np.float32)
    dx, dw, db, do = Dev(cu, x), Dev(cu, w), Dev(cu, b), Dev(cu, n=B * F * oh * oh)
    desc = cu.conv_desc(B, C, H, H, F, k, k, p, p, s, s)
    if SPLIT:
        dxs, dos = Dev(cu, n=x.size), Dev(cu, n=B * F * oh * oh)
        cu.nhwcToSplit(B * H * H, C, dx.ptr, C, dxs.ptr, C, x.size)
        nout = B * F * oh * oh
        if SPLIT_RES:
            drs = Dev(cu, rng.random(nout, dtype=np.float32))   # any bits will do as hi / lo planes for timing
            run = lambda: cu.conv2dForwardSplitResidual(desc, dxs.ptr, 0, C, 2, x.size, dw.ptr, None, db.ptr, 9, drs.ptr, 0, F,
                                                        2, nout, None, 0, F, 1, dos.ptr, 0, F, nout)
        else:
            run = lambda: cu.conv2dForwardSplit(desc, dxs.ptr, 0, C, 2, x.size, dw.ptr, None, db.ptr, 9, None, 0, 0, None, 0, F,
                                                1, dos.ptr, 0, F, nout)
    elif NHWC:
        run = lambda: cu.conv2dForwardNhwc(desc, dx.ptr, 0, C, 1, dw.ptr, None, db.ptr, 9, do.ptr, 0, F, 1)
    else:
        run = lambda: cu.conv2dForward(desc, dx.ptr, 0, dw.ptr, None, db.ptr, 9, do.ptr, 0)
    for _ in range(2):
        run()
    cu.finish()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        run()
    cu.finish()
    dt = (time.perf_counter() - t0) / n
    fl = 2.0 * B * F * oh * oh * C * k * k
    print(f"B{B} C{C} H{H} F{F} k{k} s{s}: {dt*1e3:.3f} ms  {fl/dt/1e12:.1f} TFLOP/s ({math}, {cu.lastConvKernel()})", flush=True)
    if os.environ.get("TNN_BWD"):
        dd, dwu, ddx = Dev(cu, rng.uniform(-1, 1, (B, F, oh, oh)).astype(np.float32)), Dev(cu, n=w.size), Dev(cu, n=x.size)
        for name, fn in (("wgrad", lambda: cu.conv2dBackwardFilter(desc, dx.ptr, 0, dd.ptr, 0, dwu.ptr)),
                         ("dgrad", lambda: cu.conv2dBackwardData(desc, dw.ptr, dd.ptr, 0, ddx.ptr, 0))):
            fn(); cu.finish()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            cu.finish()
            dt = (time.perf_counter() - t0) / n
            print(f"    {name}: {dt*1e3:.3f} ms  {fl/dt/1e12:.1f} TFLOP/s", flush=True)
