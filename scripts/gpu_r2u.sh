#!/bin/bash
# round 2, call 21: source-level counters of the cohort kernel on the REAL S200 launch (8 000 solves)
mkdir -p gpurun_out
timeout 1500 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:k_fmm_coh -c 1 -o gpurun_out/r2_prof_k_fmm_coh8_S200_source -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_prof_k_fmm_coh8_S200_source.log 2>&1
ls -la gpurun_out/r2_prof_k_fmm_coh8_S200_source.ncu-rep; tail -3 gpurun_out/r2_prof_k_fmm_coh8_S200_source.log
