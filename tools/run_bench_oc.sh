#!/bin/bash
# bench.py with the other_configs block; prints a digest (used during development under gpurun)
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_oc.json 2> gpurun_out/bench_oc.err
tail -3 gpurun_out/bench_oc.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_oc.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_attention"])
for k,v in d.get("other_configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload",)})
PY
