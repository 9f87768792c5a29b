#!/bin/bash
# round 2, call 15: LSMR large-n reorthogonalisation path -- tests + S200-lite LSMR bench + test4 (real data) inversion
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lsmr.py tests/test_gpu_inversion.py tests/test_fortran_abi.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -n 3 gpurun_out/r2o_pytest.log
timeout 600 python scripts/bench_lsmr.py > gpurun_out/r2o_bench_lsmr.json 2> gpurun_out/r2o_bench_lsmr.err; tail -n 1 gpurun_out/r2o_bench_lsmr.json | cut -c1-600
timeout 900 python scripts/bench_invert.py test4 0 0 > gpurun_out/r2o_invert_test4.json 2> gpurun_out/r2o_invert_test4.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_invert_test4.json").read().strip().splitlines()[-1])
print("test4", d["gpu_ms_per_iteration"], d["per_iteration_ms"]["lsmr_solve"], d["lsmr_itn"], d["vs_oracle_final_model"], d["device_s_total"])
PY
tail -3 gpurun_out/r2o_invert_test4.err
