#!/bin/bash
# round 2, call 9: eikonal mode sweep -- solves per GPU x kernel (what N = 2, 4, 8 ranks of the S200 job each run)
mkdir -p gpurun_out
OUT=gpurun_out/r2i_mode_sweep.txt; : > $OUT
for n in 125 188 250 375 500 750; do
  for v in "auto" "DAZIM_DUO=1" "DAZIM_DUO=1 DAZIM_DUO_MINB=16" "DAZIM_DUO=0" "DAZIM_TPS=1"; do
    e="$v"; [ "$v" = "auto" ] && e="DAZIM_NOP=1"
    env $e timeout 300 python bench.py --workload S200-$n --steps 1 --warmup 1 --no-cpu --no-e2e > /tmp/sw.log 2>&1
    python - "$n" "$v" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(open("/tmp/sw.log").read().strip().splitlines()[-1])
    print("solves %5d  %-32s fmm_ms %8.1f  kernel %s" % (int(sys.argv[1]) * 8, sys.argv[2], d["stage_ms"]["fmm_ms"], d["roofline"]["kernel"]))
except Exception as e:
    print("solves %5d  %-32s FAILED %s" % (int(sys.argv[1]) * 8, sys.argv[2], open("/tmp/sw.log").read()[-300:].replace("\n", " ")))
PY
  done
done
cat $OUT
