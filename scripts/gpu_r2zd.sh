#!/bin/bash
# round 2, call 29: pop overlapping the record wait + last-entry cache; two-ahead prefetch on / off; cycle counters of both roles
mkdir -p gpurun_out
for pf in 0 1; do
  DAZIM_COH_PF2=$pf timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zd_pf$pf.log 2>&1; echo "== pf2 $pf"; python scripts/show_bench.py gpurun_out/r2zd_pf$pf.log | cut -c1-200
done
DAZIM_COH_PROF=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2zd_prof.log 2>&1
grep "coh prof" gpurun_out/r2zd_prof.log | tail -5 | cut -c1-260
