#!/bin/bash
# round 2, second call: thread-per-solve eikonal kernel -- parity suite, S200 bench, quick ncu counters
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -n 5 gpurun_out/r2b_pytest.log
for v in "" "DAZIM_TPS_PER_SM=3" "DAZIM_TPS_PER_SM=4"; do
  env $v timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2b_bench_tps_${v}.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_tps_${v}.log
  echo "== $v"; tail -n 2 gpurun_out/r2b_bench_tps_${v}.log | cut -c1-1200
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_lsu.sum --clock-control none -k regex:k_fmm_tps -c 1 --csv --log-file gpurun_out/r2b_ncu_tps.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2b_ncu_tps.log 2>&1
tail -n 3 gpurun_out/r2b_ncu_tps.csv | cut -c1-600
timeout 600 python bench.py --workload T1 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2b_bench_T1.log 2>&1; tail -n 1 gpurun_out/r2b_bench_T1.log | cut -c1-1000
timeout 600 python bench.py --workload YN --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2b_bench_YN.log 2>&1; tail -n 1 gpurun_out/r2b_bench_YN.log | cut -c1-1000
