#!/bin/bash
# round 2, call 12: compute-sanitizer over the round-2 kernels (cohort eikonal kernel with 8 and 32 lanes, one-thread
# kernel, device-resident LSMR) on the small parity tests
mkdir -p gpurun_out
SEL="test_fmm_fields_bit_exact or test_forward_subset or test_heap_spill_path"
for v in "DAZIM_TPS=1" "DAZIM_TPS=1 DAZIM_COH_LANES=32" "DAZIM_TPS=1 DAZIM_COH=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/r2l_memcheck_$tag.log 2>&1; echo "memcheck $v rc=$?" | tee -a gpurun_out/r2l_memcheck_$tag.log
  tail -n 4 gpurun_out/r2l_memcheck_$tag.log
done
DAZIM_TPS=1 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact" > gpurun_out/r2l_racecheck_coh8.log 2>&1; echo "racecheck coh8 rc=$?" | tee -a gpurun_out/r2l_racecheck_coh8.log
tail -n 6 gpurun_out/r2l_racecheck_coh8.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_lsmr.py tests/test_gpu_inversion.py -m gpu -q -x -k "lsmr or plan_iterate" > gpurun_out/r2l_memcheck_lsmr.log 2>&1; echo "memcheck lsmr rc=$?" | tee -a gpurun_out/r2l_memcheck_lsmr.log
tail -n 4 gpurun_out/r2l_memcheck_lsmr.log
