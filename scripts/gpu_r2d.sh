#!/bin/bash
# round 2, fourth call (grouped deep fetch + uniform apply): cohort eikonal kernel -- eikonal parity tests, S200 bench with the cycle split, T1 / YN
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "eikonal or s200 or fmm or forward_subset or gmatrix or batched or spill or partition or empty" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -n 5 gpurun_out/r2d_pytest.log
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2d_bench_coh.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_bench_coh.log
grep "coh prof" gpurun_out/r2d_bench_coh.log | tail -4; tail -n 2 gpurun_out/r2d_bench_coh.log | cut -c1-1200
timeout 600 python bench.py --workload T1 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2d_bench_T1.log 2>&1; tail -n 1 gpurun_out/r2d_bench_T1.log | cut -c1-1000
timeout 600 python bench.py --workload YN --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2d_bench_YN.log 2>&1; tail -n 1 gpurun_out/r2d_bench_YN.log | cut -c1-1000
