#!/bin/bash
# round 2, call 14: LSMR with the segmented transposed product -- tests, inversions, per-kernel times
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 900 python -m pytest tests/test_lsmr.py tests/test_gpu_inversion.py tests/test_fortran_abi.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -n 4 gpurun_out/r2n_pytest.log
cat gpurun_out/parity_notes.txt
for t in test2 test3; do timeout 600 python scripts/bench_invert.py $t 0 0 > gpurun_out/r2n_invert_$t.json 2> gpurun_out/r2n_invert_$t.err; python - "$t" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2n_invert_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["gpu_ms_per_iteration"], d["per_iteration_ms"]["lsmr_solve"][:5], d["lsmr_itn"][:5], d["vs_reference_shipped_model"], d["vs_oracle_final_model"])
PY
done
DAZIM_LSMR_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_spmv|k_seg|k_reorth|k_dot|k_tail|k_scal|k_update|k_store' -s 400 -c 320 --csv --log-file gpurun_out/r2n_lsmr_launches_test3.csv python scripts/bench_invert.py test3 0 0 > gpurun_out/r2n_lsmr_launches_test3.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/r2n_lsmr_launches_test3.csv')))
h=[r for r in rows if r and r[0]=="ID"][0]; data=[r for r in rows if r and r[0].isdigit()]
ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size")
t=defaultdict(float); n=defaultdict(int)
for r in data: t[(r[ik][:40],r[ig])]+=float(r[iv]); n[(r[ik][:40],r[ig])]+=1
for k,v in sorted(t.items(),key=lambda x:-x[1]): print("%-42s grid %-16s %4d launches  avg %9.1f us"%(k[0],k[1],n[k],v/n[k]/1e3))
PY
timeout 600 python scripts/bench_lsmr.py > gpurun_out/r2n_bench_lsmr.log 2>&1; tail -n 2 gpurun_out/r2n_bench_lsmr.log | cut -c1-800
