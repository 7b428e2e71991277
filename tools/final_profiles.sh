#!/bin/bash
# the round-end evidence run: per-CTA / per-round clocks (debug build), thin-regime and overhead probes, ncu launch list of bench.py,
# ncu --set full captures of the solver kernel and of the cooperative value-term kernel, FP64 instruction counts.  usage: tools/final_profiles.sh (on the GPU box)
set -x
[ -f build/libigtmpc_clk.so ] || tools/build_variant.sh clk -DIGT_PHASE_CLOCKS     # debug build with the cycle counters
python tools/phase_clocks.py build/libigtmpc_clk.so 32768 > gpurun_out/r2f_clk_cfg2.log 2>&1
python tools/thin_probe.py 148 0 3 > gpurun_out/r2f_thin.log 2>&1; python tools/thin_probe.py 148 1 3 >> gpurun_out/r2f_thin.log 2>&1
python tools/overhead_probe.py >> gpurun_out/r2f_thin.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-closed-loop > gpurun_out/r2f_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-closed-loop > gpurun_out/r2f_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2f python tools/gpu_probe.py 32768 f64 1 > gpurun_out/r2f_ncu_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_coop_kernel -s 1 -c 1 -f -o gpurun_out/prof_coop_f python tools/mlp_coop_probe.py 262144 > gpurun_out/r2f_ncu_coop.log 2>&1
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:solve_kernel -s 1 -c 1 python tools/gpu_probe.py 32768 f64 1 > gpurun_out/r2f_ncu_fp64.log 2>&1
tail -3 gpurun_out/r2f_ncu_fp64.log
