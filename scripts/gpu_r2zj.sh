#!/bin/bash
# round 2, call 36 (2 GPUs): row-distributed LSMR timing after a warm-up solve: 19 M-entry and 230 M-entry systems
mkdir -p gpurun_out

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_lsmr_worker.py --big > gpurun_out/r2_dist_lsmr_2gpu.json 2> gpurun_out/r2_dist_lsmr_2gpu.err; echo "big rc=$?"; tail -n 1 gpurun_out/r2_dist_lsmr_2gpu.json; tail -n 5 gpurun_out/r2_dist_lsmr_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/dist_lsmr_worker.py --huge > gpurun_out/r2_dist_lsmr_2gpu_huge.json 2> gpurun_out/r2_dist_lsmr_2gpu_huge.err; echo "huge rc=$?"; tail -n 1 gpurun_out/r2_dist_lsmr_2gpu_huge.json; tail -n 3 gpurun_out/r2_dist_lsmr_2gpu_huge.err
