#!/bin/bash
# GPU check of the outer-iteration stage: parity tests, then the file-based driver bench (iso + joint).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inversion.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/inv_tests.log
cat gpurun_out/inv_tests.log | tail -15
timeout 300 python scripts/bench_invert.py iso 4 1 > gpurun_out/bench_invert_iso.log 2> gpurun_out/bench_invert_iso.err
tail -c 1500 gpurun_out/bench_invert_iso.log; tail -5 gpurun_out/bench_invert_iso.err
timeout 300 python scripts/bench_invert.py joint 3 1 > gpurun_out/bench_invert_joint.log 2> gpurun_out/bench_invert_joint.err
tail -c 1500 gpurun_out/bench_invert_joint.log; tail -5 gpurun_out/bench_invert_joint.err
