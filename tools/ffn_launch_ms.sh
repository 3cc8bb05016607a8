#!/bin/bash
# ms per launch of the dominant kernel (decoder FFN conv) and ms per step under environment knobs.  usage: tools/ffn_launch_ms.sh "A=1" ...
run() { env $1 python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', 'step %.3f ms  ffn launch %.1f us  frac %.3f'%(d['ms_per_step'], 1000*d['roofline']['launch_ms'], d['roofline']['frac']))"; }
run "CTTS_X=0"
for k in "$@"; do run "$k"; done
