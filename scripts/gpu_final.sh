#!/bin/bash
# Round-end check: whole GPU suite, smoke(), then the default bench line (S200) with its JSON kept under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/full_gpu_suite.log
cat gpurun_out/full_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default_final.log 2> gpurun_out/bench_default_final.err
tail -c 2500 gpurun_out/bench_default_final.log; tail -3 gpurun_out/bench_default_final.err
