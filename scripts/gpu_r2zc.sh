#!/bin/bash
# round 2, call 28: + prefetch of the node after the predicted one; stencil-side cycle counters
mkdir -p gpurun_out
DAZIM_TPS=1 timeout 600 python -m pytest tests -m gpu -x -q -k "test_fmm_fields_bit_exact or s200_eikonal or forward_subset or kernel_variants" > gpurun_out/r2zc_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2zc_pytest.log
DAZIM_COH_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zc_prof.log 2>&1
grep "coh prof" gpurun_out/r2zc_prof.log | tail -5 | cut -c1-260; python scripts/show_bench.py gpurun_out/r2zc_prof.log | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zc_plain.log 2>&1; python scripts/show_bench.py gpurun_out/r2zc_plain.log | cut -c1-300
