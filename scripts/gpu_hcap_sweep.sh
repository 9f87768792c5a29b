#!/bin/bash
# Usage: gpu_hcap_sweep.sh <workload> <hcap...>
mkdir -p gpurun_out
WL=$1; shift
for H in "$@"; do
  if [ "$H" = "auto" ]; then unset DAZIM_HCAP; else export DAZIM_HCAP=$H; fi
  timeout 600 python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_${WL}_hcap$H.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${WL}_hcap$H.log').read().strip().splitlines()[-1])
    print("$WL hcap $H", {k: round(v,1) for k,v in d['stage_ms'].items()}, round(d['ms_per_step'],1))
except Exception as e: print("$WL hcap $H failed", e)
PY
done
