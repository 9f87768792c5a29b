#!/bin/bash
# round 2, call 20: the new full-survey parity tests + the whole GPU suite timing
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
( time timeout 1800 python -m pytest tests/test_gpu_full_surveys.py -m gpu -x -q ) > gpurun_out/r2t_pytest_full_surveys.log 2>&1; tail -n 6 gpurun_out/r2t_pytest_full_surveys.log
cat gpurun_out/parity_notes.txt
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2t_pytest_all.log 2>&1; tail -n 6 gpurun_out/r2t_pytest_all.log
