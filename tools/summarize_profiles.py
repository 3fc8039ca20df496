"""Turn the raw captures of tools/profile_round.sh (gpurun_out/prof) into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "prof")
DST = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def ncu_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def kernel_key(name):
    if "k_umma_halo" in name:
        return "umma::k_umma_halo"
    if "TmaPatchLoader" in name:
        return "umma::k_umma<TmaPatchLoader> (TMA activations)"
    if "Im2colPixelLoader" in name:
        return "umma::k_umma<Im2colPixelLoader> (gather)"
    if "k_conv_direct" in name:
        return "direct::k_conv_direct"
    name = re.sub(r"\[lambda.*", "[lambda]", name)
    name = name.replace("umma::", "").replace("void ", "")
    return re.sub(r"\(.*", "", name)[:100]


def value_ns(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v * {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(u, 1)


def value_bytes(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def launch_list(src, dst, title):
    rows = [r for r in ncu_csv(src) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        k = kernel_key(r["Kernel Name"])
        ns = value_ns(r)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    with open(dst, "w") as f:
        f.write(f"# {title}\n# gpu__time_duration.sum per launch, --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# unit: ns; total over {len(rows)} launches = {tot:.0f}\nshare%  launches      total_ns  kernel\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * ns / tot:6.2f}  {n:8d}  {ns:12.0f}  {k}\n")
    return len(rows), tot


def traffic(src, dst):
    rows = ncu_csv(src)
    per = collections.OrderedDict()
    for r in rows:
        key = kernel_key(r["Kernel Name"])
        e = per.setdefault(key, dict(launches=set(), dram_bytes_read=0.0, dram_bytes_write=0.0, gpu_time_ns=0.0))
        e["launches"].add(r["ID"])
        if r["Metric Name"] == "dram__bytes_read.sum":
            e["dram_bytes_read"] += value_bytes(r)
        elif r["Metric Name"] == "dram__bytes_write.sum":
            e["dram_bytes_write"] += value_bytes(r)
        elif r["Metric Name"] == "gpu__time_duration.sum":
            e["gpu_time_ns"] += value_ns(r)
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                     "-k regex:k_umma|k_conv_direct -s 225 -c 75 python bench.py --steps 1 --warmup 3 --no-cpu (B200)",
           "note": "one step = the 75 convolution launches of YOLOv3-416 B=64", "by_kernel": {}}
    tot = 0.0
    for k, e in per.items():
        n = len(e["launches"])
        t = e["dram_bytes_read"] + e["dram_bytes_write"]
        tot += t
        out["by_kernel"][k] = dict(launches=n, dram_bytes_read=e["dram_bytes_read"], dram_bytes_write=e["dram_bytes_write"],
                                   dram_bytes_total=t, gpu_time_ns_under_ncu=e["gpu_time_ns"])
    out["dram_bytes_total_per_step"] = tot
    json.dump(out, open(dst, "w"), indent=1)
    return out


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def full_capture(rep, dst, title):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        print("no data in", rep)
        return
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, zip(units, vals)))
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none --import-source on; selected metrics of ONE launch\n")
        f.write(f"# kernel: {d.get('Kernel Name', ('', '?'))[1][:160]}\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k:90s} {d[k][1]:>16s} {d[k][0]}\n")
        s = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        for mn in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS.ARRIVE.TRANS64"):
            f.write(f"SASS {mn:24s} {sum(1 for l in s.splitlines() if mn in l)} static instructions\n")


if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    n, tot = launch_list(os.path.join(SRC, "launches.csv"), os.path.join(DST, f"{TAG}_launch_list_summary.txt"),
                         f"ncu launch list ({TAG}): python bench.py --steps 2 --warmup 3 --no-cpu ; -s 600 -c 400")
    print("launch list:", n, tot)
    launch_list(os.path.join(SRC, "cifar_launches.csv"), os.path.join(DST, f"{TAG}_cifar_launch_list_summary.txt"),
                f"ncu launch list ({TAG}): python tools/cifar_train.py --batch 128 --steps 4 --warmup 3 (all launches)")
    t = traffic(os.path.join(SRC, "conv_dram.csv"), os.path.join(DST, f"{TAG}_traffic.json"))
    print("traffic per step:", t["dram_bytes_total_per_step"])
    for rep, title in (("r01_gather_c256f512", "gathering tcgen05 conv, 256->512 3x3 @26x26, batch 64 (conv_one.py 64,256,26,512,3,1)"),
                       ("r01_halo_c64f128", "halo-tile tcgen05 conv, 64->128 3x3 @104x104, batch 64 (conv_one.py 64,64,104,128,3,1)"),
                       ("r01_tma1x1_c256f128", "TMA-activation tcgen05 conv, 256->128 1x1 @52x52, batch 64 (conv_one.py 64,256,52,128,1,1)"),
                       ("r01_wgrad_c32f32", "tile wgrad, 32->32 3x3 @32x32, batch 128 (TNN_BWD=1 conv_one.py 128,32,32,32,3,1)"),
                       ("r01_sgemm_nn4096", "TMA-fed SGEMM, NN 4096^3 (sgemm_sweep.py 3xtf32 4096)"),
                       ("r01_direct_c3f32", "direct fp32 first layer, 3->32 3x3 @416x416, batch 64 (conv_one.py 64,3,416,32,3,1)")):
        if not os.path.exists(os.path.join(SRC, rep + ".ncu-rep")):
            continue
        full_capture(os.path.join(SRC, rep + ".ncu-rep"), os.path.join(DST, f"{TAG}_ncu_{rep[4:]}.txt"), title)
    for a, b in (("bench_line.json", f"{TAG}_bench_line.json"), ("conv_layers.json", f"{TAG}_conv_layers_3xtf32.json"),
                 ("launches.csv", f"{TAG}_launches.csv")):
        open(os.path.join(DST, b), "w").write(open(os.path.join(SRC, a)).read())
    with open(os.path.join(DST, f"{TAG}_cifar_train.jsonl"), "w") as f:
        for n_ in ("cifar_b128.json", "cifar_b128_nograph.json", "cifar_b512.json", "cifar_b128_unfused.json"):
            f.write(open(os.path.join(SRC, n_)).read().strip() + "\n")
    if os.path.exists(os.path.join(SRC, "sgemm_3xtf32.txt")):
        with open(os.path.join(DST, f"{TAG}_sgemm_sweep.txt"), "w") as f:
            f.write("# tools/sgemm_sweep.py on B200: C = A*B through tnn_gemm, M = N = K = n, fp64-checked on 64 sampled entries\n")
            for n_, title in (("sgemm_3xtf32.txt", "TMA-fed operands (default)"), ("sgemm_tf32.txt", "TMA-fed operands, TF32 mode"),
                              ("sgemm_3xtf32_regstaged.txt", "TNN_TMA_GEMM=0: register-staged loaders (the path unaligned operands take)")):
                if os.path.exists(os.path.join(SRC, n_)):
                    f.write(f"## {title}\n" + open(os.path.join(SRC, n_)).read())
    if os.path.exists(os.path.join(SRC, "bench_tf32.json")):
        open(os.path.join(DST, f"{TAG}_bench_line_tf32.json"), "w").write(open(os.path.join(SRC, "bench_tf32.json")).read())
