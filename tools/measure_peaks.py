"""Measure the roofline denominators MEASURED_PEAKS.json does not carry: dense TF32 and fp32-SIMT
matmul throughput (cuBLAS through torch.matmul, 8192^3) burst and sustained, plus an fp32 copy bandwidth check.
Library calls are used here ONLY as the measuring stick (the same way the driver measures the bf16 peak).

  python tools/measure_peaks.py [out.json]
"""
import json
import subprocess
import sys
import time

import torch


def clocks():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "--id=0"], capture_output=True, text=True).stdout.strip()
        return o
    except Exception:
        return ""


def matmul_rate(n, dtype, allow_tf32, seconds=4.0):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    fl = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    # sustained: back to back for `seconds`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, int(seconds * best * 1e12 / fl))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    mid = None
    time.sleep(min(1.0, seconds / 2))
    mid = clocks()
    torch.cuda.synchronize()
    sus = reps * fl / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sus, mid


def copy_rate(nbytes=1 << 31):
    a = torch.empty(nbytes // 4, device="cuda", dtype=torch.float32).normal_()
    b = torch.empty_like(a)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "n": 8192,
           "how": "torch.matmul fp32 8192^3 (2*N^3 flops), best of 10 (burst) and back to back for 4 s (sustained); "
                  "allow_tf32=True -> cuBLAS TF32 tensor-core GEMM, False -> cuBLAS fp32 SIMT SGEMM (what the "
                  "reference's useBLAS=1 path calls, nncuda.pas:674-688)"}
    b, s, c = matmul_rate(8192, torch.float32, True)
    out.update(tf32_tflops=b, tf32_tflops_sustained=s, tf32_clocks_mid_run=c)
    b, s, c = matmul_rate(8192, torch.float32, False, seconds=3.0)
    out.update(fp32_simt_tflops=b, fp32_simt_tflops_sustained=s, fp32_clocks_mid_run=c)
    b, s, c = matmul_rate(8192, torch.bfloat16, False)
    out.update(bf16_tflops=b, bf16_tflops_sustained=s, bf16_clocks_mid_run=c)
    out["hbm_copy_gbs_fp32_2GiB"] = copy_rate()
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
