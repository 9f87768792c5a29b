#!/bin/bash
# parity tests + lite and full bench. Usage: gpu_check.sh <tag> [full]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -n 5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --workload S200-lite --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_lite_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/bench_lite_$TAG.log
tail -n 3 gpurun_out/bench_lite_$TAG.log | cut -c1-1500
if [ "$2" = "full" ]; then
timeout 1200 python bench.py --workload S200 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_full_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full_$TAG.log
tail -n 3 gpurun_out/bench_full_$TAG.log | cut -c1-1500
fi
