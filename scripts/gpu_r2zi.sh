#!/bin/bash
# round 2, call 34 (2 GPUs): multi-GPU C-ABI tests + the N=2 bench line with the new kernel rule (4 000 solves per GPU -> cohort kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2zi_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 3 gpurun_out/r2zi_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_bench_final_n2.json 2> gpurun_out/r2_bench_final_n2.err; tail -n 1 gpurun_out/r2_bench_final_n2.json | cut -c1-900
