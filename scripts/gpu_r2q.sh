#!/bin/bash
# round 2, call 17 (8 GPUs): N=8 and N=4 bench lines on the final mode switch
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2965$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2q_bench_n$N.log 2> gpurun_out/r2q_bench_n$N.err; echo "N=$N rc=$?"
  tail -n 1 gpurun_out/r2q_bench_n$N.log > gpurun_out/r2_bench_S200_${N}gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_S200_${N}gpu.json | cut -c1-420
  python - "$N" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2_bench_S200_%sgpu.json" % sys.argv[1]).read())
print("   kernel", d["roofline"]["kernel"], "e2e ms", d["e2e"]["ms_per_step"], "gather_rows_ms", d.get("gather_rows_ms"))
PY
done
