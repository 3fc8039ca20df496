#!/bin/bash
# bench.py in the three tensor modes; one summary line each (full JSON lines under gpurun_out/)
for m in bf16x3 3xtf32 tf32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --math $m --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - "$m" <<'EOF'
import json, sys
m = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{m}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(m, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms | e2e", round(d["e2e"]["value"], 1), "| conv",
          round(r["all_conv"]["ms"], 2), "ms", {k: (v["ms"], v["tflops"]) for k, v in r["by_kernel"].items()}, d["clocks"])
except Exception as e:
    print(m, "FAILED", e, open(f"gpurun_out/bench_{m}.err").read()[-600:])
EOF
done
