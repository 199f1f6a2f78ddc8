#!/bin/bash
# A/B of bench.py variants on ONE box: usage scripts/ab_bench.sh "<env A>" "<env B>" [extra bench args]
# prints value / ms_per_step / e2e / clocks for each run (alternating A, B, A, B to average out clock drift)
A="$1"; B="$2"; shift 2
for i in 1 2; do
  for V in "$A" "$B"; do
    env $V python bench.py --no-cpu-baseline --steps 30 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$V', '$*', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
