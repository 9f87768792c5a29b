#!/bin/bash
# round 2, call 23: heap warp with explicit reconvergence points -- parity, cycle split, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eikonal or s200 or fmm or forward_subset or full_surveys or falls_back" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -n 3 gpurun_out/r2w_pytest.log
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2w_full_prof.log 2>&1; grep "coh prof" gpurun_out/r2w_full_prof.log | tail -2 | cut -c1-300
for L in 8 16 32; do
  DAZIM_COH_LANES=$L timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2w_full_L$L.log 2>&1; echo "lanes $L"; python scripts/show_bench.py gpurun_out/r2w_full_L$L.log | cut -c1-230
done
DAZIM_TPS=1 timeout 300 python bench.py --workload S200-500 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2w_500.log 2>&1; python scripts/show_bench.py gpurun_out/r2w_500.log | cut -c1-200
