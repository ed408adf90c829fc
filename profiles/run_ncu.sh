#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash profiles/run_ncu.sh <tag>
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --batch 32"
mkdir -p gpurun_out
if [ -z "${FULL_ONLY:-}" ]; then
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 240 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
fi
[ -n "${LIST_ONLY:-}" ] || ncu --set full --clock-control none --import-source on -k regex:'gemm2?_tcgen05_kernel|attention_fwd_kernel' -s 150 -c 8 \
    -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/plain_${TAG}.log | cut -c1-300
tail -3 gpurun_out/ncu_full_${TAG}.log
