#!/bin/bash
# round 2, call 32: full GPU suite on the final eikonal kernel + racecheck / memcheck of the cohort kernel with records computed ahead
mkdir -p gpurun_out; rm -f gpurun_out/parity_notes.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2zg_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/r2zg_pytest_gpu.log
DAZIM_TPS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact" > gpurun_out/r2zg_racecheck_coh8.log 2>&1; echo "racecheck coh8 rc=$?" | tee -a gpurun_out/r2zg_racecheck_coh8.log
tail -n 4 gpurun_out/r2zg_racecheck_coh8.log
DAZIM_TPS=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact or forward_subset" > gpurun_out/r2zg_memcheck_coh8.log 2>&1; echo "memcheck coh8 rc=$?" | tee -a gpurun_out/r2zg_memcheck_coh8.log
tail -n 4 gpurun_out/r2zg_memcheck_coh8.log
