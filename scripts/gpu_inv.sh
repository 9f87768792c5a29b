#!/bin/bash
# GPU check of the outer-iteration stage: parity tests (incl. the reference's real test2/test3 examples), the file-based
# driver bench on the real examples and on the synthetic T1-shaped survey, then the whole GPU suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inversion.py tests/test_fortran_abi.py tests/test_lsmr.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/inv_tests.log
tail -5 gpurun_out/inv_tests.log
for c in test2 test3; do
  timeout 400 python scripts/bench_invert.py $c 0 1 > gpurun_out/bench_invert_$c.log 2> gpurun_out/bench_invert_$c.err
  tail -c 2000 gpurun_out/bench_invert_$c.log; tail -3 gpurun_out/bench_invert_$c.err
done
timeout 300 python scripts/bench_invert.py iso 4 1 > gpurun_out/bench_invert_iso.log 2> gpurun_out/bench_invert_iso.err
tail -c 1800 gpurun_out/bench_invert_iso.log; tail -3 gpurun_out/bench_invert_iso.err
timeout 300 python scripts/bench_invert.py joint 3 1 > gpurun_out/bench_invert_joint.log 2> gpurun_out/bench_invert_joint.err
tail -c 1800 gpurun_out/bench_invert_joint.log; tail -3 gpurun_out/bench_invert_joint.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/full_gpu_suite.log
cat gpurun_out/full_gpu_suite.log
