"""Sample SM clock and board power while one convolution runs back to back (is a kernel power-capped?).
usage: clock_probe.py math B,C,H,F,k,s [seconds]"""
import os, subprocess, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorium_b200 import tensor
from util import Dev
MATH = {"fp32": 0, "3xtf32": 1, "tf32": 2}
cu = tensor.initCUDART(0)
cu.set_math_mode(MATH[sys.argv[1]])
B, C, H, F, k, s = (int(v) for v in sys.argv[2].split(","))
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
p = k // 2
oh = (H + 2 * p - k) // s + 1
rng = np.random.default_rng(0)
dx = Dev(cu, rng.random((B, C, H, H), dtype=np.float32))
dw = Dev(cu, rng.uniform(-0.05, 0.05, (F, C * k * k)).astype(np.float32))
db, do = Dev(cu, np.zeros(F, np.float32)), Dev(cu, n=B * F * oh * oh)
desc = cu.conv_desc(B, C, H, H, F, k, k, p, p, s, s)
samples = []
stop = False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        samples.append(r.stdout.strip())
        time.sleep(0.2)
th = threading.Thread(target=sampler); th.start()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < secs:
    for _ in range(50):
        cu.conv2dForward(desc, dx.ptr, 0, dw.ptr, None, db.ptr, 9, do.ptr, 0)
    cu.finish(); n += 50
dt = time.perf_counter() - t0
stop = True; th.join()
fl = 2.0 * B * F * oh * oh * C * k * k
print(f"{sys.argv[1]} {sys.argv[2]}: {dt/n*1e3:.3f} ms/launch sustained, {fl*n/dt/1e12:.1f} TFLOP/s; samples (MHz, W, reasons):")
print("  " + " | ".join(samples[2:12]))
