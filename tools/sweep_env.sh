#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ... : one short bench.py line per value of an environment tuning knob
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 --e2e-log2-items 16 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); k=j['kernel_ms_per_step']; print(round(j['value']/1e6,3),'M pairs/s', round(j['ms_per_step'],2),'ms', {a[1:25]:round(b,2) for a,b in k.items() if 'matvec' in a})
"
done
