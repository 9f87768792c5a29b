#!/bin/bash
# round 2, call 7: full GPU suite (cohort kernel with 8/16/32 solves per heap warp, device-resident LSMR), lanes sweep, inversions
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -n 4 gpurun_out/r2g_pytest.log
for L in 8 16 32; do
  DAZIM_COH_PROF=1 DAZIM_COH_LANES=$L timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g_full_L$L.log 2>&1
  echo "== lanes $L"; grep "coh prof" gpurun_out/r2g_full_L$L.log | tail -2 | cut -c1-300; python scripts/show_bench.py gpurun_out/r2g_full_L$L.log | cut -c1-260
done
for W in T1 YN; do for L in 8 16; do
  DAZIM_COH_LANES=$L timeout 600 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g_${W}_L$L.log 2>&1; echo "== $W lanes $L"; python scripts/show_bench.py gpurun_out/r2g_${W}_L$L.log | cut -c1-250
done; done
DAZIM_TPS=0 timeout 600 python bench.py --workload T1 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g_T1_legacy.log 2>&1; echo "== T1 legacy"; python scripts/show_bench.py gpurun_out/r2g_T1_legacy.log | cut -c1-250
for t in test2 test3; do timeout 600 python scripts/bench_invert.py $t 0 0 > gpurun_out/r2g_invert_$t.json 2> gpurun_out/r2g_invert_$t.err; tail -c 900 gpurun_out/r2g_invert_$t.json; echo; done
