#!/bin/bash
# same-box A/B of two checkouts of the repo: tools/ab_bench.sh <dirA> <dirB> [bench args]
A=$1; B=$2; shift 2
for r in 1 2 3; do for d in $A $B; do
  python $d/bench.py --steps 18 --warmup 4 --no-cpu-baseline --no-other-configs "$@" > /tmp/ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
print("$d", round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1))
PY
done; done
