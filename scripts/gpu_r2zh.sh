#!/bin/bash
# round 2, call 33: final numbers -- counters + launch list of the final cohort kernel, cohort / half-warp crossover, default bench line
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_fmm_coh -c 1 --csv --log-file gpurun_out/r2_k_fmm_coh8_S200_counters_final.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_k_fmm_coh8_S200_counters_final.log 2>&1
tail -n 8 gpurun_out/r2_k_fmm_coh8_S200_counters_final.csv | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_S200_final.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_launches_S200_final.log 2>&1
for n in 375 438 500; do
  DAZIM_TPS=1 timeout 300 python bench.py --workload S200-$n --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zh_coh_$n.log 2>&1; echo "== cohort S200-$n"; python scripts/show_bench.py gpurun_out/r2zh_coh_$n.log | cut -c1-160
done
for n in 438 500; do
  DAZIM_TPS=0 timeout 300 python bench.py --workload S200-$n --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2zh_hw_$n.log 2>&1; echo "== half-warp S200-$n"; python scripts/show_bench.py gpurun_out/r2zh_hw_$n.log | cut -c1-160
done
timeout 900 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; cat gpurun_out/r2_bench_final_n1.json
