#!/bin/bash
# first GPU contact: parity tests, smoke, small + full bench with hard timeouts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; command -v gfortran >> gpurun_out/gpu.txt 2>&1 || echo "no gfortran" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --workload S40 --steps 2 --warmup 1 > gpurun_out/bench_s40.log 2>&1; echo "rc=$?" >> gpurun_out/bench_s40.log
timeout 900 python bench.py --workload S200-lite --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_s200lite.log 2>&1; echo "rc=$?" >> gpurun_out/bench_s200lite.log
tail -5 gpurun_out/*.log
