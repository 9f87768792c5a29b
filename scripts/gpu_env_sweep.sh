#!/bin/bash
# Usage: gpu_env_sweep.sh <workload> "ENV1=a ENV2=b" "ENV1=c" ...
mkdir -p gpurun_out
WL=$1; shift
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/sweep_$i.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_$i.log').read().strip().splitlines()[-1])
    print("$WL [$E]", {k: round(v,1) for k,v in d['stage_ms'].items()}, round(d['ms_per_step'],1))
except Exception as e: print("$WL [$E] failed", e)
PY
done
