#!/bin/bash
# round 2, call 39: feasibility of hiding K1 behind K3 (two streams)
mkdir -p gpurun_out
DAZIM_SM_SLACK=4096 timeout 600 python scripts/overlap_k1_k3.py > gpurun_out/r2zm_overlap.json 2> gpurun_out/r2zm_overlap.err; echo rc=$?; tail -n 1 gpurun_out/r2zm_overlap.json; tail -n 3 gpurun_out/r2zm_overlap.err
