#!/bin/bash
# round 2, call 10: cohort kernel with prefetched ancestors -- parity + timing at 3000 / 4000 / 8000 solves
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eikonal or s200" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -n 3 gpurun_out/r2j_pytest.log
for n in 375 500; do
  DAZIM_TPS=1 timeout 300 python bench.py --workload S200-$n --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2j_$n.log 2>&1; python scripts/show_bench.py gpurun_out/r2j_$n.log | cut -c1-200
done
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2j_full.log 2>&1
grep "coh prof" gpurun_out/r2j_full.log | tail -2 | cut -c1-300; python scripts/show_bench.py gpurun_out/r2j_full.log | cut -c1-260
