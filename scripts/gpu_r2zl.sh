#!/bin/bash
# round 2, call 38 (2 GPUs): row-distributed tail after moving NCCL's lazy set-up into dazim_comm_create: per-iteration times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist_lsmr.py -q -x > gpurun_out/r2zl_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2zl_pytest.log
grep "row-distributed inversion" gpurun_out/parity_notes.txt | tail -1 | cut -c1-1800
