#!/bin/bash
# round 2, call 43: the default bench line on the final code (what the driver runs at round end)
mkdir -p gpurun_out
timeout 112 python bench.py > gpurun_out/r2_bench_final2_n1.json 2> gpurun_out/r2_bench_final2_n1.err; echo rc=$?; tail -n 1 gpurun_out/r2_bench_final2_n1.json | cut -c1-3000
