#!/bin/bash
# round 2, call 5: where do the heap warp's cycles go?  fine split + lanes-per-heap-warp experiment (divergence)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "eikonal or s200_eik" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -n 3 gpurun_out/r2e_pytest.log
for L in 32 16 8 2; do
  DAZIM_COH_PROF=1 DAZIM_COH_LANES=$L timeout 600 python bench.py --workload S200-lite --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2e_lanes_$L.log 2>&1
  echo "== lanes $L"; grep "coh prof" gpurun_out/r2e_lanes_$L.log | tail -4; python scripts/show_bench.py gpurun_out/r2e_lanes_$L.log 2>/dev/null | head -3
done
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2e_full.log 2>&1; grep "coh prof" gpurun_out/r2e_full.log | tail -4; python scripts/show_bench.py gpurun_out/r2e_full.log | head -3
