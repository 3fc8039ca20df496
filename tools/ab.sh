#!/bin/bash
# A/B on ONE box: the YOLOv3 step with each library ab/<name>.so (TNN_LIB; "tree" = the in-tree build), alternating.
# usage (under gpurun): bash tools/ab.sh "base tree" [rounds] [extra bench.py flags]
names=${1:-"base tree"}; rounds=${2:-2}; shift 2
mkdir -p gpurun_out/ab
for r in $(seq $rounds); do
  for v in $names; do
    if [ $v = tree ]; then unset TNN_LIB; else export TNN_LIB=$PWD/ab/$v.so; fi
    python bench.py --no-secondary --no-cpu --steps ${STEPS:-10} "$@" --layers-out gpurun_out/ab/layers_$v.json > gpurun_out/ab/bench_$v.json 2> gpurun_out/ab/bench_$v.err
    python - <<P
import json
d = json.loads(open("gpurun_out/ab/bench_$v.json").read())
print("$v", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
P
  done
done
