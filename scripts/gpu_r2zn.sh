#!/bin/bash
# round 2, call 40: ncu --set full of the FINAL cohort kernel (records computed ahead) on the small S40 grid + full GPU suite
mkdir -p gpurun_out
DAZIM_TPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fmm_coh -c 1 -o gpurun_out/r2_prof_k_fmm_coh8_S40_final -f python bench.py --workload S40 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_prof_k_fmm_coh8_S40_final.log 2>&1
ls -la gpurun_out/r2_prof_k_fmm_coh8_S40_final.ncu-rep
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2zn_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/r2zn_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zn_smoke.log 2>&1; tail -n 2 gpurun_out/r2zn_smoke.log
