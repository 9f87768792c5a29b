#!/bin/bash
mkdir -p gpurun_out
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2x_full_prof.log 2>&1; grep "coh prof" gpurun_out/r2x_full_prof.log | tail -2 | cut -c1-300
