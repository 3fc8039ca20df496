"""Times the two launches of one fused LSTM step separately (GEMM on the hidden state, gate kernel)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorium_b200 import tensor
from tensorium_b200.layers import TNNetState
from tensorium_b200.rnn_layers import TLSTMLayer
from tensorium_b200.tensor import TSingleTensor
cu = tensor.initCUDART(0)
S, B, I, H = 4, 256, 256, 1024
l = TLSTMLayer(S * B, I, H, S)
x = np.zeros((S * B, I), np.float32)
l.forwardGPU(TNNetState(input=TSingleTensor(x.shape, S * B, data=x)))
cu.finish()
def t(f, n=200):
    f(); cu.finish()
    t0 = time.perf_counter()
    for _ in range(n): f()
    cu.finish()
    return (time.perf_counter() - t0) / n * 1e6
H4 = 4 * H
print("gemm  [%d x %d x %d] beta=0: %.1f us" % (B, H4, H, t(lambda: cu.gemm(False, True, B, H4, H, 1.0, l._h.devData, 0, H, l.w_all.devData, 0, H, 0.0, l.wz.devData, 0, H4))))
print("gemm  same, beta=1 (no memset): %.1f us" % t(lambda: cu.gemm(False, True, B, H4, H, 1.0, l._h.devData, 0, H, l.w_all.devData, 0, H, 1.0, l.wz.devData, 0, H4)))
print("gemm  [%d x %d x %d]: %.1f us" % (1024, 1024, 1024, t(lambda: cu.gemm(False, True, 1024, 1024, 1024, 1.0, l.w_all.devData, 0, H, l.w_all.devData, 0, H, 0.0, l.uz.devData, 0, 1024))))
print("gates: %.1f us" % t(lambda: cu.lstmGates(B, H, l.wz.devData, l.bw_all.devData, l.uz.devData, 0, l._c.devData, l._h.devData, l.Cell.devData, l.output.devData, 0)))
print("fill 4MB: %.1f us" % t(lambda: cu.fill(B * H4, l.wz.devData, 0, 0.0, 1)))
S = 128
l = TLSTMLayer(S * B, I, H, S)
x = np.zeros((S * B, I), np.float32)
xin = TSingleTensor(x.shape, S * B, data=x)
st = TNNetState(input=xin)
l.forwardGPU(st); cu.finish()
print("whole forward, %d steps: %.1f us" % (S, t(lambda: l.forwardGPU(st), 5)))
print("up-front gemm [%d x %d x %d]: %.1f us" % (B * S, H4, I, t(lambda: cu.gemm(False, True, B * S, H4, I, 1.0, xin.devData, 0, I, l.u_all.devData, 0, I, 0.0, l.uz.devData, 0, H4), 5)))
print("up-front bias: %.1f us" % t(lambda: cu.forwardBias(B * S * H4, l.uz.devData, 0, H4, l.bu_all.devData, 1, B * S), 5))
def loop():
    for i in range(S):
        cu.gemm(False, True, B, H4, H, 1.0, l._h.devData, 0, H, l.w_all.devData, 0, H, 0.0, l.wz.devData, 0, H4)
        cu.lstmGates(B, H, l.wz.devData, l.bw_all.devData, l.uz.devData, i * B * H4, l._c.devData, l._h.devData, l.Cell.devData, l.output.devData, i * B * H)
print("step loop: %.1f us" % t(loop, 5))
