#!/bin/bash
# One bench line per block type and per configs[4] length (same JSON schema as the headline), into gpurun_out/
for w in transformer fastformer conformer liu2021_m64 liu2021_m256 liu2021_m1024; do
  python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || tail -3 gpurun_out/r02_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_$w.json"))
    print("$w", "%.3f ms/step" % d["ms_per_step"], "%.0f frames/s" % d["value"], "e2e %.0f" % d["e2e"]["value"], "cpu %.0f" % d.get("cpu_baseline", {}).get("value", 0), "frac", d["roofline"]["frac"])
except Exception as e:
    print("$w failed", e)
PY
done
