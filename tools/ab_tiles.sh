#!/bin/bash
# A/B of the tile-shape knobs on ONE box: ms per step of the headline benchmark
run() { env $1 python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '%.3f ms  e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run "CTTS_X=0"
run "CTTS_WIDE_MIN_N=256"
run "CTTS_TILE_192=1"
run "CTTS_WIDE_MIN_N=256 CTTS_TILE_192=1"
run "CTTS_X=0"
