#!/bin/bash
# round 2, call 30: toggles of the cohort kernel (bit 0: two-ahead prefetch, bit 1: no last-entry cache, bit 2: no fence before the X arrive)
mkdir -p gpurun_out
for pf in 1 3 2 5; do
  DAZIM_COH_PF2=$pf timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2ze_pf$pf.log 2>&1; echo "== opt $pf"; python scripts/show_bench.py gpurun_out/r2ze_pf$pf.log | cut -c1-200
done
