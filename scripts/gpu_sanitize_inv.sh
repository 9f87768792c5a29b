#!/bin/bash
# compute-sanitizer over the outer-iteration kernels (dazim_invert.cu, k_reorth of dazim_lsmr.cu), then the whole GPU suite.
mkdir -p gpurun_out
SEL="cal_ddat_sigma or tikhonov_rows or plan_iterate or iterate_device or calddatsigma_and_tikhonov"
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_inversion.py tests/test_fortran_abi.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_inv_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_inv_memcheck.log
tail -4 gpurun_out/sanitize_inv_memcheck.log
timeout 55 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_inversion.py tests/test_lsmr.py -m gpu -q -x -k "cal_ddat_sigma or lsmr" > gpurun_out/sanitize_inv_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_inv_racecheck.log
tail -4 gpurun_out/sanitize_inv_racecheck.log
timeout 90 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/full_gpu_suite.log
cat gpurun_out/full_gpu_suite.log
