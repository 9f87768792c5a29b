#!/bin/bash
# compute-sanitizer passes over the small parity tests (memcheck: out-of-bounds / misaligned; racecheck:
# shared-memory hazards between the heap warp and the stencil warp of k_fmm_duo and inside k_fmm)
mkdir -p gpurun_out
SEL="test_fmm_fields_bit_exact or test_forward_subset or test_heap_spill_path"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
DAZIM_DUO=0 DAZIM_SPC=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_racecheck_spc2.log 2>&1; echo "racecheck spc2 rc=$?" | tee -a gpurun_out/sanitize_racecheck_spc2.log
tail -n 6 gpurun_out/sanitize_*.log
