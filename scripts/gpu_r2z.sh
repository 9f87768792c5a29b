#!/bin/bash
# round 2, call 25: stencil threads issue all loads at once + prefetch of the predicted next node: parity + timing
mkdir -p gpurun_out
for qs in 1 2; do
  DAZIM_TPS=1 DAZIM_COH_QS=$qs timeout 600 python -m pytest tests -m gpu -x -q -k "test_fmm_fields_bit_exact or s200_eikonal or forward_subset" > gpurun_out/r2z_pytest_qs$qs.log 2>&1; echo "pytest qs=$qs rc=$?"; tail -n 2 gpurun_out/r2z_pytest_qs$qs.log
done
for cfg in "8 1" "8 2" "16 1"; do
  set -- $cfg
  DAZIM_COH_PROF=1 DAZIM_COH_LANES=$1 DAZIM_COH_QS=$2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2z_L$1_Q$2.log 2>&1
  echo "== lanes $1 qs $2"; grep "coh prof" gpurun_out/r2z_L$1_Q$2.log | tail -2 | head -1 | cut -c1-200; python scripts/show_bench.py gpurun_out/r2z_L$1_Q$2.log | cut -c1-200
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2z_plain.log 2>&1; python scripts/show_bench.py gpurun_out/r2z_plain.log | cut -c1-300
