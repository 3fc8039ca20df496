"""Turn the raw captures of tools/profile_round2.sh (gpurun_out/prof2) into the committed summaries under profiles/:
  r02_launch_list_summary_<mode>.txt   per-kernel shares of one YOLOv3 step (ncu gpu__time_duration, cold-cache, serialised)
  r02_traffic.json                     DRAM bytes per launch of the convolution kernels (dram__bytes_read + write)
  r02_ncu_<capture>.txt                the headline metrics of each `--set full` capture
usage: python tools/summarize_profiles2.py"""
import collections
import csv
import glob
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "prof2")
DST = os.path.join(ROOT, "profiles")


def ncu_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def kernel_key(name):
    if "k_conv_nhwc_pair" in name:
        return "nhwc::k_conv_nhwc_pair (CTA pairs, TMA im2col activations, TMA store)"
    if "k_conv_nhwc" in name:
        return "nhwc::k_conv_nhwc (TMA im2col activations, TMA store)"
    if "k_conv_image" in name:
        return "nhwc::k_conv_image (planar image in, tensor cores, TMA store)"
    if "k_conv_direct" in name:
        return "direct::k_conv_direct (NHWC out)"
    name = re.sub(r"\[lambda.*", "[lambda]", name)
    name = name.replace("void ", "")
    return re.sub(r"\(.*", "", name)[:90]


UNIT = {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9,
        "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(row):
    return float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1)


def launch_list(mode):
    path = os.path.join(SRC, f"launches_{mode}.csv")
    if not os.path.exists(path):
        return None
    per = collections.OrderedDict()
    for r in ncu_csv(path):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = val(r)
    agg = collections.OrderedDict()
    tot = 0.0
    for d in per.values():
        k = kernel_key(d["name"])
        a = agg.setdefault(k, dict(launches=0, ns=0.0, dram=0.0))
        a["launches"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        tot += d.get("gpu__time_duration.sum", 0.0)
    with open(os.path.join(DST, f"r02_launch_list_summary_{mode}.txt"), "w") as f:
        f.write(f"# YOLOv3-416 batch 64, channels-innermost graph, {mode}: {len(per)} consecutive launches (about four steps)\n")
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write("# (cold-cache, serialised per-launch times: compare SHARES with bench.py's CUDA-event shares, not absolutes)\n")
        f.write(f"# total {tot / 1e6:.3f} ms\nshare%  launches   total_ms   dram_MB/launch  kernel\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            f.write(f"{100 * a['ns'] / tot:6.2f}  {a['launches']:8d}  {a['ns'] / 1e6:9.3f}  {a['dram'] / a['launches'] / 1e6:14.1f}  {k}\n")
    return {k: dict(launches=a["launches"], dram_bytes_total=a["dram"], ms_total=a["ns"] / 1e6) for k, a in agg.items()}


RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def full_capture(rep):
    name = os.path.basename(rep).replace(".ncu-rep", "").replace(".raw.csv", "")
    if rep.endswith(".raw.csv"):
        out = open(rep).read()
    else:
        try:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=600).stdout
        except Exception as e:
            print("ncu failed for", rep, e)
            return
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(DST, f"{name}.txt"), "w") as f:
        f.write(f"# {name}: ncu --set full --clock-control none (one replay set per launch); selected raw metrics\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {d.get('Kernel Name', '')[:150]}\n   grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n")
            for k in hdr:
                if any(k == want or (want in k and len(k) < len(want) + 30) for want in RAW_KEYS):
                    u = units[hdr.index(k)]
                    f.write(f"   {k:95s} {d[k]:>18s} {u}\n")


def main():
    os.makedirs(DST, exist_ok=True)
    traffic = {}
    for mode in ("bf16x3", "tf32", "3xtf32"):
        t = launch_list(mode)
        if t:
            traffic[mode] = {"by_kernel": t, "source": f"ncu launch list of bench.py --math {mode} (tools/profile_round2.sh yolo)"}
    if traffic:
        json.dump(traffic, open(os.path.join(DST, "r02_traffic.json"), "w"), indent=1)
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep")) + glob.glob(os.path.join(SRC, "*.raw.csv"))):
        full_capture(rep)
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
