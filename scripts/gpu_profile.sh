#!/bin/bash
# ncu launch list + full capture of the dominant kernels (1 GPU). Usage: gpu_profile.sh <tag> [workload]
TAG=${1:-r1}; WL=${2:-S200-lite}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_fmm -c 1 -o gpurun_out/prof_fmm_$TAG -f $CMD > gpurun_out/prof_fmm_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_disp|k_eigen|k_trace|k_assemble' -c 6 -o gpurun_out/prof_other_$TAG -f $CMD > gpurun_out/prof_other_$TAG.log 2>&1
ls -la gpurun_out
