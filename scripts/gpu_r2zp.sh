#!/bin/bash
# round 2, call 42: last-entry cache in the shipped heap round: S200 timing first, then the full GPU suite + smoke on this code
mkdir -p gpurun_out; rm -f gpurun_out/parity_notes.txt
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zp_S200.log 2>&1; python scripts/show_bench.py gpurun_out/r2zp_S200.log | cut -c1-220
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2zp_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/r2zp_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zp_smoke.log 2>&1; tail -n 2 gpurun_out/r2zp_smoke.log
