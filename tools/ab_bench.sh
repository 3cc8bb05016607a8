#!/bin/bash
# A/B the dispatch knobs with the real benchmark (CUDA-graph replay, L2-resident activations): ncu launch lists flush the
# caches before every kernel and overstate the short memory-bound kernels, so decisions are taken on ms_per_step.
# usage: tools/ab_bench.sh "NAME=VAL ..." "NAME=VAL ..."      (one bench run per argument; "" = defaults)
for cfg in "$@"; do
  out=$(env $cfg timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null)
  echo "$cfg => $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step %.4f  e2e_ms %.4f  ffn1_ms %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["launch_ms"]))')"
done
