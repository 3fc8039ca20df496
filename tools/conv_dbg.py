"""Debug helper: run conv forward for cfg = B,C,H,W,F,k,p (stride 1) and compare with numpy fp64. usage: conv_dbg.py cfg..."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev
cu = tensor.initCUDART(0)
rng = np.random.default_rng(0)
for a in sys.argv[1:]:
    vals = [int(v) for v in a.split(",")]
    B, C, H, W, F, k, p = vals[:7]
    st = vals[7] if len(vals) > 7 else 1
    x = rng.random((B, C, H, W), dtype=np.float32)
    w = rng.uniform(-0.05, 0.05, (F, C, k, k)).astype(np.float32)
    oh, ow = (H + 2 * p - k) // st + 1, (W + 2 * p - k) // st + 1
    PAD = int(os.environ.get('PADF', '0'))
    dx, dw, do = Dev(cu, x, pad_front=PAD), Dev(cu, w), Dev(cu, n=B * F * oh * ow)
    try:
        cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, st, st), dx._base, PAD, dw.ptr, None, None, 4, do.ptr, 0)
        got = do.read((B, F, oh, ow))
    except Exception as e:
        print(a, "CRASH", str(e)[-60:]); sys.exit(1)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (p, p), (p, p)))
    want = np.zeros((B, F, oh, ow))
    for kr in range(k):
        for kc in range(k):
            want += np.einsum("bchw,fc->bfhw", xp[:, :, kr:kr + st * (oh - 1) + 1:st, kc:kc + st * (ow - 1) + 1:st], w[:, :, kr, kc].astype(np.float64))
    print(a, "max rel err", float(np.abs(got - want).max() / np.abs(want).max()), flush=True)
