#!/bin/bash
# round 2, call 6: cohort kernel with back pointers in E (no id table)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eikonal or s200 or fmm or forward_subset or gmatrix or batched or spill" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -n 3 gpurun_out/r2f_pytest.log
DAZIM_COH_PROF=1 timeout 600 python bench.py --workload S200-lite --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_lite.log 2>&1
grep "coh prof" gpurun_out/r2f_lite.log | tail -2; python scripts/show_bench.py gpurun_out/r2f_lite.log | cut -c1-250
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_full_prof.log 2>&1; grep "coh prof" gpurun_out/r2f_full_prof.log | tail -2
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_full.log 2>&1; python scripts/show_bench.py gpurun_out/r2f_full.log | cut -c1-300
timeout 600 python bench.py --workload T1 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_T1.log 2>&1; python scripts/show_bench.py gpurun_out/r2f_T1.log | cut -c1-250
timeout 600 python bench.py --workload YN --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_YN.log 2>&1; python scripts/show_bench.py gpurun_out/r2f_YN.log | cut -c1-250
