#!/bin/bash
# round 2, call 18: final confirmation -- full suite, racecheck of the cohort kernel, mode switch at 1 504 / 2 000 solves, S200 with the cycle split
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -n 3 gpurun_out/r2r_pytest.log
DAZIM_TPS=1 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact" > gpurun_out/r2r_racecheck_coh8.log 2>&1; echo "racecheck coh8 rc=$?" | tee -a gpurun_out/r2r_racecheck_coh8.log; tail -n 3 gpurun_out/r2r_racecheck_coh8.log
for n in 188 250 300; do
  timeout 300 python bench.py --workload S200-$n --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2r_$n.log 2>&1
  python - "$n" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2r_%s.log" % sys.argv[1]).read().strip().splitlines()[-1])
print("solves %5d  fmm_ms %8.1f  kernel %s" % (int(sys.argv[1]) * 8, d["stage_ms"]["fmm_ms"], d["roofline"]["kernel"]))
PY
done
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2r_full_prof.log 2>&1; grep "coh prof" gpurun_out/r2r_full_prof.log | tail -2 | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err; python scripts/show_bench.py gpurun_out/r2r_bench_default.json | cut -c1-420
