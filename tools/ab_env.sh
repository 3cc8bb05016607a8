#!/bin/bash
# A/B of environment knobs on ONE box: ms per step of the headline benchmark.  usage: tools/ab_env.sh "A=1" "B=2 C=3" ...
run() { env $1 python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '%.3f ms  e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run "CTTS_X=0"
for k in "$@"; do run "$k"; done
run "CTTS_X=0"
