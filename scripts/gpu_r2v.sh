#!/bin/bash
# round 2, call 22: fall-back test + whole GPU suite on the final library
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2v_pytest_all.log 2>&1; tail -n 6 gpurun_out/r2v_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
