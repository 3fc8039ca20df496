#!/usr/bin/env python
"""bench.py — YOLOv3-416 inference images/s on B200 through the Tensorium drop-in backend.

Workload = BASELINE.json configs[3]: YOLOv3 416x416 inference, batch 64 per GPU, random-init weights
(BN folded like MSCOCOYolo.pas:327-340), synthetic U[0,1) images; batch-sharded across GPUs with no collective
(weak scaling: every rank runs its own batch of 64).  One "step" = one forward pass of the 107-layer net.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math bf16x3|3xtf32|tf32|fp32]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` includes the pinned-host H2D copy of
the batch and the D2H read of the three yolo outputs every step.  `roofline` is for the dominant kernel (the
tcgen05 implicit-GEMM conv): algorithmic FLOPs (2*B*F*OH*OW*C*k*k summed over the 75 convs) / summed CUDA-event
durations of those launches.  `cpu_baseline` / `--impl reference` time the reference's CPU algorithm (oracle/, the
C restatement — FreePascal is not available to build the reference itself) on the host cores.

Default arithmetic: bf16x3 — every fp32 operand split into two bf16 numbers, three tensor-core MMAs per product,
fp32 accumulation; the same 1e-4 per-layer gate as 3xTF32 (tests/test_gpu_nhwc.py, test_gpu_yolo.py; measured error
vs fp64 is LOWER than 3xTF32's at K >= 1024, profiles/r02_accuracy_vs_k.txt).  The line also carries, as secondary
objects measured in the same run: `modes` (the step in 3xtf32 and tf32 arithmetic), `cifar10_train` (BASELINE
configs[2]: deepCIFAR10 training with the NCCL all-reduce of the update bucket captured in the step graph, global
batch 128 split over the GPUs and 128 per GPU) and `roofline_membound` (the HBM-bound kernels against the measured
copy bandwidth).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
SIZE = 416
METRIC = "YOLOv3-416 inference images/sec"
MATH = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16x3": 3}
MMA_PER_PRODUCT = {"fp32": 1, "3xtf32": 3, "tf32": 1, "bf16x3": 3}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


def rank_seed(rank: int) -> int:
    """Every rank draws its own batch (batch sharding: independent images per GPU, no data-path collective)."""
    return 1234 + rank


def reduce_max_ms(ms: float, world: int, device="cuda") -> float:
    """Timing contract: the step time of the job is the MAX over ranks."""
    if world <= 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_images_per_s(world: int, steps: int, ms_total: float, per_gpu_batch: int = BATCH) -> float:
    """Whole-job throughput: all ranks' images divided by the max-over-ranks time."""
    return per_gpu_batch * world * steps / (ms_total / 1000.0)


def synthetic_images(batch, seed=1234):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random((batch, 3, SIZE, SIZE), dtype=np.float32)


# ------------------------------------------------------------------------------------------ CPU (reference) arm
def cpu_forward_rate(n_images, reps=1, seed=1234):
    """images/s of the reference's CPU algorithm (oracle restatement, AVX2/FMA build flavour of the YOLO sample,
    MSCOCO_Yolo/Makefile:8) on all host threads.  Returns (images_per_s, seconds, threads)."""
    from oracle import oracle as orc
    from oracle import ref_net
    from tensorium_b200.nnet import init_yolov3_params, yolov3_spec
    orc.build()
    orc.set_mode(orc.MODE_FMA)
    orc.set_threads(os.cpu_count() or 1)
    spec = yolov3_spec()
    params = init_yolov3_params(spec, seed=seed)
    x = synthetic_images(n_images, seed)
    t0 = time.perf_counter()
    for _ in range(reps):
        ref_net.forward(spec, params, x)
    dt = time.perf_counter() - t0
    orc.set_mode(orc.MODE_PAS)
    return n_images * reps / dt, dt, orc.max_threads()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank, world):
    if rank != 0:
        return
    rates = []
    for _ in range(args.warmup):      # every warm-up pass is a full single-image forward (about 1 s each)
        cpu_forward_rate(1)
    t_all = 0.0
    for _ in range(args.steps):
        r, dt, threads = cpu_forward_rate(1)
        rates.append(r); t_all += dt
    value = args.steps / t_all
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"YOLOv3-{SIZE} inference (107 layers, BN fused), reference CPU algorithm "
                               f"(C restatement of ntensors.pas cblas_sgemm+sim2Col), 1 image per step",
                   "global_batch": 1, "image": SIZE,
                   "same_config": "same net, weights and image size as the GPU arm; the step is ONE image instead of 64 "
                                  "(the reference's CPU path loops per image - Conv2D ntensors.pas:8326-8333 - so images/s "
                                  "does not depend on the batch); a bounded sample so that the run ends within minutes"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x 1 image full forward ({cpu_model()})"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def membound_roofline(cu, stream, hbm_gbs):
    """The HBM-bound kernels of the path against the measured copy bandwidth: algorithmic bytes per element (DESIGN.md
    §4) x elements / CUDA-event time on tensors far larger than the 126 MB L2."""
    import torch
    out = {}

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def entry(name, kernel, bytes_per_elem, n, ms, what):
        gbs = bytes_per_elem * n / (ms * 1e-3) / 1e9
        out[name] = {"kernel": kernel, "elements": n, "bytes_per_element": bytes_per_elem, "ms": round(ms, 4),
                     "achieved": round(gbs, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(gbs / hbm_gbs, 3), "what": what}

    B, C, HW = 64, 64, 208 * 208                      # the YOLOv3 64-channel 208x208 tensor at batch 64: 709 MB
    n = B * C * HW
    a, b_, c_ = (torch.rand(n, device="cuda") for _ in range(3))
    rows = B * HW
    entry("shortcut_add", "nhwc::k_rows<true> (tnn_rows_add)", 12, n,
          timed(lambda: cu.rowsAdd(rows, C, a.data_ptr(), C, b_.data_ptr(), C, c_.data_ptr(), C)),
          "TAddLayer.forwardGPU (naddlayer.pas:891-910) on channels-innermost rows: 2 reads + 1 write")
    entry("leaky_activate", "k_map_vec (tnn_activate_array)", 8, n, timed(lambda: cu.ActivateArray(n, a.data_ptr(), 0, 9)),
          "ActivateArray(leaky) in place (nncuda.pas:524-542): 1 read + 1 write")
    entry("axpy", "k_map_vec (tnn_axpy)", 12, n, timed(lambda: cu.axpy(n, 0.5, a.data_ptr(), 0, 1, b_.data_ptr(), 0, 1)),
          "y += a*x (nncuda.pas:945-964): 2 reads + 1 write")
    del c_
    # batch-norm, training: deepCIFAR10's 32-channel 32x32 layer scaled to batch 2048 (268 MB per tensor)
    Bb, Cb, HWb = 2048, 32, 32 * 32
    nb = Bb * Cb * HWb
    x, xn = torch.empty(nb, device="cuda"), torch.empty(nb, device="cuda")
    o = a[:nb]
    dl = b_[:nb]
    vec = [torch.ones(Cb, device="cuda") for _ in range(9)]
    p = [v.data_ptr() for v in vec]
    entry("batchnorm_forward_training", "k_channel_reduce<2> + planemap::k_plane<0,2> (tnn_batchnorm_forward_training)", 20, nb,
          timed(lambda: cu.batchNormForwardTraining(Bb, Cb, HWb, o.data_ptr(), x.data_ptr(), xn.data_ptr(), p[0], p[1], p[2],
                                                    p[3], 0.1, p[4], p[5], 1)),
          "batchNormGPU training (nbaselayer.pas:583-643) fused: 1 statistics read + 1 read / 3 writes (out, x, x_norm)")
    entry("batchnorm_backward", "k_channel_reduce<3> + planemap::k_plane<2,0> (tnn_batchnorm_backward)", 32, nb,
          timed(lambda: cu.batchNormBackward(Bb, Cb, HWb, o.data_ptr(), 1, dl.data_ptr(), x.data_ptr(), xn.data_ptr(), p[0],
                                             p[1], p[4], p[6], p[7], p[8])),
          "DeriveArray + batchNormBackGPU (nbaselayer.pas:662-714) in two passes: reads delta, out, x, x_norm, then "
          "reads delta, out, x and writes delta")
    np_ = 1 << 27                                       # 134 M parameters (512 MB per tensor)
    entry("sgd_update", "k_sgd (tnn_sgd_update)", 16, np_,
          timed(lambda: cu.sgdUpdate(np_, a.data_ptr(), b_.data_ptr(), 1e-6, 0.9, -1e-6)),
          "updateGPU (nConvolutionLayer.pas:1309-1322) fused: read+write params and updates")
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from tensorium_b200 import tensor
    from tensorium_b200.layers import TConvolutionalLayer, TNNetState
    from tensorium_b200.nnet import build_yolov3

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    cu = tensor.initCUDART(local_rank)  # raises if libtensorium_b200.so / a B200 is missing: no fallback
    cu.set_stream(stream.cuda_stream)
    cu.set_math_mode(MATH[args.math])
    # channels-innermost graph: the 23 shortcut adds are folded into the preceding convolution's epilogue
    # (tnn_conv2d_forward_nhwc_residual; bit-identical to the separate layers, tests/test_gpu_yolo.py)
    net = build_yolov3(BATCH, SIZE, seed=1234, channels_last=args.layout == "nhwc",
                       fuse_shortcuts=args.layout == "nhwc" and not args.no_fuse,
                       fuse_yolo=args.layout == "nhwc" and not args.no_fuse)   # yolo activation in the heads' epilogues
    # BF16x3: activations between convolutions are pre-split bf16 hi/lo planes (tnn_conv2d_forward_split_residual) — the
    # planes the kernel would otherwise derive in shared memory for every tap; the folded shortcuts read their operand
    # from its planes (hi + lo = the fp32 value to 2^-18 relative), so every tensor is written once
    can_split = args.layout == "nhwc" and not args.no_fuse and not args.no_split
    net.setSplitActivations(can_split and args.math == "bf16x3", split_into_1x1=args.split_1x1,
                            planes_only=not args.fp32_shortcuts)
    def elide(m):   # planes-only graph: sources written straight into the route buffers, single-source routes aliased
        return net.elideRoutes() if can_split and m == "bf16x3" and not args.fp32_shortcuts and not args.no_elide else 0
    routes_elided = elide(args.math)
    # shortcut fusion into the conv epilogue: "" (off, default), "gather" or "all".  Measured (B200, batch 64): 27.7 ms
    # un-fused, 28.0 ms with the MMA-bound gathering layers fused, 28.9 ms with all 23 fused — the residual read in
    # the four epilogue warps costs what the separate add kernels cost, so it stays off.
    FUSE = os.environ.get("TNN_FUSE_SHORTCUTS", "")
    x_host = torch.from_numpy(synthetic_images(BATCH, rank_seed(rank))).pin_memory()
    nbytes_in = x_host.numel() * 4
    cu.writeBuffer(net.input.devData, nbytes_in, x_host.numpy())
    yolos = net.yolo_layers()
    out_host = [torch.empty(l.output.FShape, dtype=torch.float32).pin_memory() for l in yolos]
    nbytes_out = sum(t.numel() * 4 for t in out_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()  # the last result has reached the host before the end event is recorded
        e1.record(stream)
        barrier()
        return reduce_max_ms(e0.elapsed_time(e1), world)

    def step_resident():
        net.predict()

    # ---- end-to-end pipeline through the public API: every step copies ITS batch from pinned host memory and reads
    # ITS three yolo outputs back; a second TNNCuda instance (own stream) carries the copies so that H2D of batch i+1
    # and D2H of result i-1 overlap the compute of batch i (double-buffered input, staged outputs).
    from tensorium_b200.nncuda import TNNCuda
    from tensorium_b200.tensor import TSingleTensor
    cu_copy = TNNCuda(local_rank)
    in_dev = [net.input, TSingleTensor(net.input.FShape, groups=BATCH, host=False)]
    out_stage = [[TSingleTensor(l.output.FShape, groups=BATCH, host=False) for l in yolos] for _ in range(2)]
    # the step's RESULT as the host needs it: the cells that pass getDetections' `objectness > thresh` test, found on the
    # device (tnn_yolo_candidates) — a list of at most MAX_CAND records instead of the three output tensors
    MAX_CAND, THRESH = 8192, 0.45   # MSCOCOYolo.pas:252
    cand_floats = 4 + MAX_CAND * (yolos[0].classes + 9)
    cand_stage = [TSingleTensor((cand_floats,), groups=1, host=False) for _ in range(2)]
    cand_host = torch.empty(cand_floats, dtype=torch.float32).pin_memory()
    nbytes_cand = cand_floats * 4
    e2e_mode = {"full": False}
    ev_in = [cu.createEvent(), cu.createEvent()]        # H2D of the batch in in_dev[k] has landed
    ev_done = [cu.createEvent(), cu.createEvent()]      # compute that used in_dev[k] / out_stage[k] is finished
    ev_drained = [cu.createEvent(), cu.createEvent()]   # D2H of out_stage[k] is finished
    e2e_state = {"i": 0}
    cu_copy.writeBufferAsync(in_dev[0].devData, nbytes_in, x_host.data_ptr())
    cu_copy.recordEvent(ev_in[0])
    for k in range(2):
        cu.recordEvent(ev_done[k])
        cu_copy.recordEvent(ev_drained[k])

    def step_e2e():
        # The host never blocks; ordering is by events.  Copy stream: H2D(i+1) then D2H(i); compute stream: step i.
        i = e2e_state["i"]
        k, kn = i & 1, (i + 1) & 1
        cu_copy.waitEvent(ev_done[kn])                     # compute i-1 no longer reads in_dev[kn]
        cu_copy.writeBufferAsync(in_dev[kn].devData, nbytes_in, x_host.data_ptr())   # H2D of batch i+1
        cu_copy.recordEvent(ev_in[kn])
        cu.waitEvent(ev_in[k])                             # batch i is on the device
        cu.waitEvent(ev_drained[k])                        # result i-2 has left out_stage[k]
        net.predict(in_dev[k])                             # compute batch i
        if e2e_mode["full"]:
            for l, st in zip(yolos, out_stage[k]):         # results -> staging (device copy, same stream)
                cu.copy(st.Size(), l.output.devData, 0, 1, st.devData, 0, 1)
        else:
            net.collectCandidates(THRESH, MAX_CAND, cand_stage[k])
        cu.recordEvent(ev_done[k])
        cu_copy.waitEvent(ev_done[k])
        if e2e_mode["full"]:
            for st, t in zip(out_stage[k], out_host):      # D2H of result i, overlaps compute of batch i+1
                cu_copy.readBufferAsync(st.devData, t.numel() * 4, t.data_ptr())
        else:
            cu_copy.readBufferAsync(cand_stage[k].devData, nbytes_cand, cand_host.data_ptr())
        cu_copy.recordEvent(ev_drained[k])
        e2e_state["i"] = i + 1

    def e2e_drain():
        cu.finish()
        cu_copy.finish()

    step_resident()  # first pass: every conv layer learns which kernel it dispatches to
    if FUSE and args.layout != "nhwc":
        nf = net.fuseShortcuts(None if FUSE == "all" else (FUSE,))
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = cu.launch_count()
    ms = timed(step_resident, args.steps)
    launches = cu.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e = timed(step_e2e, args.steps, drain=e2e_drain)
    n_cand = int(cand_host[:1].view(torch.int32).item())
    assert n_cand <= MAX_CAND, "candidate list overflow: the e2e result would need the full outputs"
    e2e_mode["full"] = True      # the reference's own hand-over for comparison: all three output tensors to the host
    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e_full = timed(step_e2e, args.steps, drain=e2e_drain)
    e2e_mode["full"] = False

    value = aggregate_images_per_s(world, args.steps, ms)
    e2e_value = aggregate_images_per_s(world, args.steps, ms_e2e)

    # ---- roofline of the dominant kernel: the tcgen05 implicit-GEMM conv, all 75 launches of one step
    convs = [(i, l) for i, l in enumerate(net.layers) if isinstance(l, TConvolutionalLayer)]
    per_layer = []
    reps = 3
    for i, l in convs:
        src = net.input if i == 0 else net.layers[i - 1].output
        st = TNNetState(input=src)
        l.forwardGPU(st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            l.forwardGPU(st)
        e1.record(stream)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        fl = 2.0 * l.batch * l.filters * l.outH * l.outW * l.c * l.kernelSize ** 2
        comp = 4.0 * (l.batch * l.c * l.h * l.w + l.filters * l.c * l.kernelSize ** 2 +
                      l.batch * l.filters * l.outH * l.outW)
        per_layer.append(dict(layer=i, c=l.c, f=l.filters, k=l.kernelSize, s=l.stride_x, hw=l.outH, ms=t,
                              tflops=fl / t / 1e9, flops=fl, compulsory_bytes=comp, kernel=cu.lastConvKernel()))
    conv_ms = sum(p["ms"] for p in per_layer)
    conv_flops = float(net.conv_flops())
    peaks, which = load_peaks()
    # TF32 dense peak: measured on this pool's B200s with tools/measure_peaks.py (cuBLAS TF32 8192^3, sustained — the
    # kernels are timed inside a long step); MEASURED_PEAKS.json carries bf16 only
    tf32_src = f"TF32 dense = bf16_tflops_sustained/2 of {which} MEASURED_PEAKS.json"
    tf32_peak = float(peaks.get("tf32_tflops_sustained", peaks["bf16_tflops_sustained"] / 2.0))
    ppath = os.path.join(ROOT, "profiles", "r02_peaks.json")
    if "tf32_tflops_sustained" not in peaks and os.path.exists(ppath):
        pj = json.load(open(ppath))
        tf32_peak = float(pj["tf32_tflops_sustained"])
        tf32_src = ("profiles/r02_peaks.json: cuBLAS TF32 8192^3 sustained over 4 s on a B200 of this pool "
                    f"(burst {pj['tf32_tflops']:.0f}; tools/measure_peaks.py)")
    # the convolutions dispatch to several kernels; the roofline object describes the one with the largest share of
    # the step, `by_kernel` lists them all (algorithmic FLOPs of their launches / their summed CUDA-event time)
    by_kernel = {}
    for p_ in per_layer:
        b = by_kernel.setdefault(p_["kernel"], dict(launches=0, ms=0.0, flops=0.0, compulsory_bytes=0.0))
        b["launches"] += 1
        b["ms"] += p_["ms"]
        b["flops"] += p_["flops"]
        b["compulsory_bytes"] += p_["compulsory_bytes"]
    for b in by_kernel.values():
        b["tflops"] = b["flops"] / b["ms"] / 1e9
        b["share_of_step"] = b["ms"] / (ms / args.steps)
    # bf16x3 issues kind::f16 MMAs: its bound is the bf16 dense peak (driver-measured, sustained); the tf32 modes are
    # bounded by the TF32 peak measured with tools/measure_peaks.py
    if args.math == "bf16x3":
        pipe_peak, pipe_src = float(peaks["bf16_tflops_sustained"]), f"bf16_tflops_sustained of {which} MEASURED_PEAKS.json"
    else:
        pipe_peak, pipe_src = tf32_peak, tf32_src
    issue = MMA_PER_PRODUCT[args.math]
    dom_name = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
    dom = by_kernel[dom_name]
    achieved = dom["tflops"]
    # DRAM traffic of the dominant kernel's launches, from a separate `ncu` capture (never measured under the timed run)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        ent = tj.get(args.math, {}).get("by_kernel", {}).get(dom_name)
        if ent:
            traffic = ent["dram_bytes_total"] / max(ent["launches"], 1)
            traffic_src = (f"profiles/r02_traffic.json (ncu dram__bytes_read+write, mean of {ent['launches']} launches, "
                           f"{args.math})")

    # ---- secondary objects, measured after the headline regions (all ranks take part: the training run has a collective)
    modes, cifar, membound = None, None, None
    if not args.no_secondary:
        modes = {}
        for m in ("bf16x3", "3xtf32", "tf32"):
            if m == args.math:
                continue
            cu.set_math_mode(MATH[m])
            net.setSplitActivations(can_split and m == "bf16x3", split_into_1x1=args.split_1x1,
                                        planes_only=not args.fp32_shortcuts)
            elide(m)
            for _ in range(3):
                step_resident()
            ms_m = timed(step_resident, 5)
            modes[m] = {"images_per_s": aggregate_images_per_s(world, 5, ms_m), "ms_per_step": ms_m / 5, "steps": 5,
                        "gate": "5e-3" if m == "tf32" else "1e-4"}
        cu.set_math_mode(MATH[args.math])
        net.setSplitActivations(can_split and args.math == "bf16x3", split_into_1x1=args.split_1x1,
                            planes_only=not args.fp32_shortcuts)
        elide(args.math)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cifar_train
        cu.set_math_mode(MATH["3xtf32"])   # the training kernels (wgrad, dgrad, connected layers) are the planar tf32 ones
        cifar = {"workload": "deepCIFAR10 (nmodels.pas:82-111) training step: host batch in, forward, backward, NCCL "
                             "all-reduce of the update bucket (tnn_allreduce_sum, captured in the step's CUDA graph), SGD; "
                             "synthetic 32x32 images, 3xtf32", "math": "3xtf32"}
        # (a failure in a secondary object must not cost the headline line: it is recorded in the object instead; every
        # rank runs the same deterministic code, so all ranks fail or none does and nobody waits in a collective)
        def secondary(name, fn):
            try:
                cifar[name] = fn()
            except Exception as e:  # noqa: BLE001
                cifar[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1:
            secondary("batch128", lambda: cifar_train.train_throughput(cu, stream, 128, 100, 5, rank, world))
        else:
            per = max(128 // world, 1)
            secondary("global_batch_128_split", lambda: cifar_train.train_throughput(cu, stream, per, 100, 5, rank, world))
            secondary("batch128_per_gpu", lambda: cifar_train.train_throughput(cu, stream, 128, 100, 5, rank, world))
            secondary("global_batch_128_split_syncbn",
                      lambda: cifar_train.train_throughput(cu, stream, per, 100, 5, rank, world, sync_bn=True))
        cu.set_math_mode(MATH[args.math])
        if rank == 0:
            try:
                membound = membound_roofline(cu, stream, float(peaks["hbm_gbs"]))
            except Exception as e:  # noqa: BLE001
                membound = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        if args.layers_out:
            os.makedirs(os.path.dirname(args.layers_out) or ".", exist_ok=True)
            json.dump(dict(math=args.math, conv_ms=conv_ms, step_ms=ms / args.steps, per_layer=per_layer),
                      open(args.layers_out, "w"), indent=1)
        cpu = None
        if world == 1 and not args.no_cpu:
            r, dt, threads = cpu_forward_rate(args.cpu_images)
            cpu = {"value": r, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_images} image(s) full YOLOv3-416 forward in {dt:.1f}s, oracle (C restatement of "
                             f"the reference CPU path, FMA flavour), {cpu_model()}"}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.math != "tf32" else "tf32",
            "dtype_note": {"bf16x3": "fp32 tensors and accumulation; every product as 3 bf16 MMAs on hi/lo halves (1e-4 gate)",
                           "3xtf32": "fp32 tensors and accumulation; every product as 3 TF32 MMAs on hi/lo halves (1e-4 gate)",
                           "tf32": "fp32 tensors, TF32 products (5e-3 gate)", "fp32": "fp32 CUDA-core FMA"}[args.math],
            "data": "synthetic",
            "config": {"workload": f"YOLOv3-{SIZE} inference, 107 layers (75 conv), BN folded, random-init weights",
                       "per_gpu_batch": BATCH, "global_batch": BATCH * world, "image": SIZE,
                       "math": args.math, "layout": args.layout,
                       "activations": (("pre-split bf16 hi/lo planes between convolutions (4 B/element, the operands BF16x3 "
                                        "multiplies); " + ("fp32 rows where a shortcut / 1x1 convolution / yolo layer reads"
                                                           if args.fp32_shortcuts else
                                                           "every tensor written once, the folded shortcuts add hi + lo of "
                                                           "their operand's planes; fp32 only for the yolo layers"))
                                       if getattr(net, "split_activations", False) else "fp32"),
                       "shortcuts_fused_into_conv": sum(bool(getattr(l, "_fused_into_conv", False)) and l.layerType != "yolo"
                                                        for l in net.layers),
                       "routes_elided": routes_elided,
                       "yolo_layers_fused_into_heads": sum(getattr(l, "_yolo", None) is not None for l in net.layers),
                       "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": "activations streamed per step = 14.6 GB per GPU, far larger than the 126 MB L2 (no flush needed)"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": nbytes_in,
                    "d2h_bytes_per_step": nbytes_cand, "ms_per_step": ms_e2e / args.steps,
                    "note": "H2D of batch i+1 and D2H of result i run on a second stream and overlap the compute of batch i, "
                            "so the region costs what the compute costs; it is timed after the resident region and may come "
                            "out a percent faster or slower with the clock the power cap allows at that moment",
                    "result": f"detection candidates (objectness > {THRESH}, TYoloLayer.getDetections' test run on the "
                              f"device by tnn_yolo_candidates): count + up to {MAX_CAND} records of 89 floats, read back "
                              f"every step; {n_cand} candidates in the last step (random-init heads, objectness bias -4)",
                    "full_outputs": {"value": aggregate_images_per_s(world, args.steps, ms_e2e_full), "unit": "images/s",
                                     "d2h_bytes_per_step": nbytes_out, "ms_per_step": ms_e2e_full / args.steps,
                                     "result": "all three yolo output tensors copied to the host every step (what the "
                                               "reference's host-side getDetections reads)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pipe_peak, "unit": "TFLOP/s",
                         "frac": achieved / pipe_peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src,
                         "compulsory_bytes": dom["compulsory_bytes"] / dom["launches"],
                         "kernel": f"{dom_name}: tcgen05 kind::{'f16 (bf16 operands)' if args.math == 'bf16x3' else 'tf32'} "
                                   f"implicit-GEMM conv, {dom['launches']} launches/step",
                         "avg_launch_ms": dom["ms"] / dom["launches"],
                         "peak_source": pipe_src,
                         "achieved_note": "ALGORITHMIC FLOPs (2*B*F*OH*OW*C*k*k, counted once) / CUDA-event time; the "
                                          f"kernel issues {issue} MMA(s) per product, so the tensor pipe itself runs at "
                                          "issued_frac of its peak",
                         "frac_of_tf32_peak": achieved / tf32_peak, "tf32_peak": tf32_peak, "tf32_peak_source": tf32_src,
                         "kernel_ms_per_step": dom["ms"], "share_of_step": dom["share_of_step"],
                         "mma_issue_factor": issue,
                         "issued_frac": achieved * issue / pipe_peak,
                         "all_conv": {"launches": len(per_layer), "ms": conv_ms, "tflops": conv_flops / conv_ms / 1e9,
                                      "share_of_step": conv_ms / (ms / args.steps)},
                         "by_kernel": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()
                                           if kk in ("launches", "ms", "tflops", "share_of_step")}
                                       for k, v in by_kernel.items()}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "modes": modes,
            "cifar10_train": cifar,
            "roofline_membound": membound,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="bf16x3", choices=list(MATH))
    ap.add_argument("--no-secondary", action="store_true", help="skip the modes / cifar10_train / roofline_membound objects")
    ap.add_argument("--layout", default="nhwc", choices=["nhwc", "nchw"],
                    help="internal activation layout of the graph: channels-innermost (TMA im2col kernels) or planar")
    ap.add_argument("--no-fuse", action="store_true", help="keep the shortcut adds as separate kernels")
    ap.add_argument("--split-1x1", action="store_true", help="bf16x3: 1x1 convolutions read pre-split planes too (their "
                    "producers then write fp32 rows AND planes)")
    ap.add_argument("--fp32-shortcuts", action="store_true", help="bf16x3: keep fp32 rows for the shortcut operands and the "
                    "1x1 convolutions' inputs (the graph of the first half of round 2) instead of planes everywhere")
    ap.add_argument("--no-elide", action="store_true", help="bf16x3 planes-only graph: keep the copying route layers")
    ap.add_argument("--no-split", action="store_true", help="bf16x3: fp32 activations between all layers (no pre-split planes)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-images", type=int, default=2)
    ap.add_argument("--layers-out", default="", help="write the per-conv-layer timing table (JSON) here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
