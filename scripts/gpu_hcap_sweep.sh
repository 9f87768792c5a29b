#!/bin/bash
mkdir -p gpurun_out
for H in 1024 512 256; do
  DAZIM_HCAP=$H timeout 600 python bench.py --workload S200 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_full_hcap$H.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_full_hcap$H.log').read().strip().splitlines()[-1])
    print("hcap $H", d['stage_ms'], d['ms_per_step'])
except Exception as e: print("hcap $H failed", e)
PY
done
