#!/bin/bash
# round 2, call 19 (4 GPUs): N=4 bench line on the final mode switch (duo<16>, every solve resident)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2s_bench_n4.log 2> gpurun_out/r2s_bench_n4.err; echo "rc=$?"
tail -n 1 gpurun_out/r2s_bench_n4.log > gpurun_out/r2_bench_S200_4gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_S200_4gpu.json | cut -c1-420
