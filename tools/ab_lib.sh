#!/bin/bash
# A/B of two builds of the library on ONE box (lib/libctts_b200_ab.so = the other build): ms per step of the headline benchmark
AB=comprehensive-transformer-tts_b200/lib/libctts_b200_ab.so
run() { env $1 python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '%.3f ms  e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run "CTTS_X=0"
run "CTTS_B200_LIB=$AB"
run "CTTS_X=0"
run "CTTS_B200_LIB=$AB"
