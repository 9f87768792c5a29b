#!/bin/bash
# Round-1 evidence: (1) launch list of one full S200 step, (2) DRAM bytes per launch on S200-lite,
# (3) --set full captures of every kernel family on the small S40 workload.
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_S200.csv python bench.py --workload S200 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/launches_${TAG}_S200.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/dram_${TAG}_S200lite.csv python bench.py --workload S200-lite --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/dram_${TAG}_S200lite.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fmm|k_disp|k_eigen|k_trace|k_assemble|k_dice' -c 8 -o gpurun_out/prof_${TAG}_S40 -f python bench.py --workload S40 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/prof_${TAG}_S40.log 2>&1
ls -la gpurun_out | tail -8
