"""Diagnostic: index-encoded input, one-hot filters -> shows which input element each output picked up."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev
cu = tensor.initCUDART(0)
B, C, H, W, k, p, st = 1, 32, 16, 8, 3, 1, 1
if len(sys.argv) > 1:
    B, C, H, W, k, p, st = (int(v) for v in sys.argv[1].split(","))
c_, h_, w_ = np.meshgrid(np.arange(C), np.arange(H), np.arange(W), indexing="ij")
x = np.broadcast_to((c_ * 4096 + h_ * 64 + w_ + 1).astype(np.float32), (B, C, H, W)).copy()
taps = [(c, kr, kc) for c in (0, 5, C - 1) for kr in range(k) for kc in range(k)]
F = len(taps)
w = np.zeros((F, C, k, k), np.float32)
for f, (c, kr, kc) in enumerate(taps):
    w[f, c, kr, kc] = 1
xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
OH, OW = (H + 2 * p - k) // st + 1, (W + 2 * p - k) // st + 1
want = np.stack([xp[:, c, kr:kr + st * (OH - 1) + 1:st, kc:kc + st * (OW - 1) + 1:st] for (c, kr, kc) in taps], axis=1)
dx, dw, do = Dev(cu, x), Dev(cu, w), Dev(cu, n=want.size)
cu.conv2dForward(cu.conv_desc(B, C, H, W, F, k, k, p, p, st, st), dx.ptr, 0, dw.ptr, None, None, 4, do.ptr, 0)
got = do.read(want.shape)
def dec(v):
    v = int(v) - 1
    return "zero" if v < 0 else f"c{v // 4096}h{(v // 64) % 64}w{v % 64}"
print("mismatch fraction", float((got != want).mean()))
for f, (c, kr, kc) in enumerate(taps[:12]):
    bad = np.argwhere(got[0, f] != want[0, f])
    print(f"filter(c{c},kr{kr},kc{kc}): {len(bad)} bad of {OH*OW}", [(int(a), int(b), dec(got[0, f, a, b]), dec(want[0, f, a, b])) for a, b in bad[:6]])
