#!/bin/bash
# A/B on ONE box with one build: the YOLOv3 step with an environment switch off / on, alternating.
# usage (under gpurun): bash tools/ab_env.sh TNN_PDL [rounds] [extra bench.py flags]
var=$1; rounds=${2:-2}; shift 2
mkdir -p gpurun_out/ab
for r in $(seq $rounds); do
  for v in ${VALS:-0 1}; do
    env $var=$v timeout 300 python bench.py --no-secondary --no-cpu --steps 10 "$@" --layers-out gpurun_out/ab/layers_$var$v.json > gpurun_out/ab/bench_$var$v.json 2> gpurun_out/ab/bench_$var$v.err
    python - <<P
import json
d = json.loads(open("gpurun_out/ab/bench_$var$v.json").read())
print("$var=$v", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
P
  done
done
