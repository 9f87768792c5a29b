#!/bin/bash
# round 2, call 11: full GPU suite on the final code + the evidence captures for profiles/
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -n 4 gpurun_out/r2k_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -n 1 gpurun_out/r2k_smoke.log
# (1) launch list of one full S200 step (shares, cold cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_S200.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_launches_S200.log 2>&1
grep -c k_fmm gpurun_out/r2_launches_S200.csv
# (2) DRAM traffic + instruction counters of the dominant kernel on the full S200 launch (a few passes of a 7 s kernel)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_fmm_coh -c 1 --csv --log-file gpurun_out/r2_k_fmm_coh8_S200_counters.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_k_fmm_coh8_S200_counters.log 2>&1
tail -n 8 gpurun_out/r2_k_fmm_coh8_S200_counters.csv | cut -c1-300
# (3) --set full of the cohort kernel on the small S40 grid (source page; ~40 replays)
DAZIM_TPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fmm_coh -c 1 -o gpurun_out/r2_prof_k_fmm_coh8_S40 -f python bench.py --workload S40 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_prof_k_fmm_coh8_S40.log 2>&1
ls -la gpurun_out/r2_prof_k_fmm_coh8_S40.ncu-rep
# (4) K1: FP64 instruction mix of k_disp on S200 (FLOP/s against the non-tensor FP64 peak)
timeout 900 ncu --metrics gpu__time_duration.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_disp -c 1 --csv --log-file gpurun_out/r2_k_disp_S200_fp64.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_k_disp_S200_fp64.log 2>&1
tail -n 8 gpurun_out/r2_k_disp_S200_fp64.csv | cut -c1-300
# (5) the two bench arms with the driver's flags shortened (3 + 3)
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 --cpu-sources 320 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 600 gpurun_out/r2_bench_reference_arm.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default_S200_1gpu.json 2> gpurun_out/r2_bench_default.err; python scripts/show_bench.py gpurun_out/r2_bench_default_S200_1gpu.json
