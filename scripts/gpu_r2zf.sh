#!/bin/bash
# round 2, call 31: records computed ahead (5 075 ms variant) with / without the two-ahead prefetch
mkdir -p gpurun_out
for pf in 0 1; do
  DAZIM_COH_PF2=$pf timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zf_pf$pf.log 2>&1; echo "== pf2 $pf"; python scripts/show_bench.py gpurun_out/r2zf_pf$pf.log | cut -c1-200
done
