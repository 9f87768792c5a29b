#!/bin/bash
# round 2, call 41: non-alive stencil words kept out of the IEEE sqrt / div slow paths: parity subset + timings of every kernel tier
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "test_fmm_fields_bit_exact or s200_eikonal or forward_subset or kernel_variant" > gpurun_out/r2zo_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2zo_pytest.log
for wl in S200 S200-125 T1 YN; do
  timeout 300 python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zo_$wl.log 2>&1; echo "== $wl"; python scripts/show_bench.py gpurun_out/r2zo_$wl.log | cut -c1-220
done
