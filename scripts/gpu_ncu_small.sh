#!/bin/bash
# full ncu capture of one kernel on a small workload. Usage: gpu_ncu_small.sh <tag> <kernel-regex> [workload] [skip]
TAG=${1:-x}; K=${2:-k_fmm}; WL=${3:-S40}; SKIP=${4:-0}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu > gpurun_out/prof_${TAG}.log 2>&1
tail -n 3 gpurun_out/prof_${TAG}.log
