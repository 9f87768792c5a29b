#!/bin/bash
# round 2, call 13: racecheck after the end-of-march handshake fix; parity; LSMR per-kernel times on the reference's test3
mkdir -p gpurun_out
DAZIM_TPS=1 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact" > gpurun_out/r2m_racecheck_coh8.log 2>&1; echo "racecheck coh8 rc=$?" | tee -a gpurun_out/r2m_racecheck_coh8.log
tail -n 4 gpurun_out/r2m_racecheck_coh8.log
DAZIM_TPS=1 DAZIM_COH_LANES=32 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_fmm_fields_bit_exact" > gpurun_out/r2m_racecheck_coh32.log 2>&1; echo "racecheck coh32 rc=$?" | tee -a gpurun_out/r2m_racecheck_coh32.log
tail -n 4 gpurun_out/r2m_racecheck_coh32.log
timeout 900 python -m pytest tests -m gpu -x -q -k "eikonal or s200 or fmm or forward_subset" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -n 3 gpurun_out/r2m_pytest.log
DAZIM_LSMR_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_spmv|k_reorth|k_dot|k_tail|k_scal|k_update|k_store' -s 400 -c 300 --csv --log-file gpurun_out/r2m_lsmr_launches_test3.csv python scripts/bench_invert.py test3 0 0 > gpurun_out/r2m_lsmr_launches_test3.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/r2m_lsmr_launches_test3.csv')))
h=[r for r in rows if r and r[0]=="ID"][0]; data=[r for r in rows if r and r[0].isdigit()]
ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size")
t=defaultdict(float); n=defaultdict(int)
for r in data: t[(r[ik][:40],r[ig])]+=float(r[iv]); n[(r[ik][:40],r[ig])]+=1
for k,v in sorted(t.items(),key=lambda x:-x[1]): print("%-42s grid %-16s %4d launches  avg %9.1f us"%(k[0],k[1],n[k],v/n[k]/1e3))
PY
