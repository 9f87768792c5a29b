#!/bin/bash
# round 2, first call: parity suite (with the new S200 tests), reference arm, full S200 bench
mkdir -p gpurun_out
nproc > gpurun_out/r2a_nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -n 5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_ref.log
tail -n 3 gpurun_out/r2a_ref.log | cut -c1-2500
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_bench.log
tail -n 3 gpurun_out/r2a_bench.log | cut -c1-3000
