#!/bin/bash
# one GPU session: full gpu tests, ncu counters of the headline kernel and of the tile-mode LIN, bench N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputests_r02b.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:ba_window_cluster -c 2 --csv --log-file gpurun_out/ba_cluster_r02b.csv python bench.py --steps 1 --warmup 1 --no-extra > gpurun_out/ncu_bench.log 2>&1
ncu --metrics $M --clock-control none -k regex:"k_lg_lin|k_lg_backsub|k_lg_solve" -c 12 --csv --log-file gpurun_out/lg_r02b.csv python scripts/lg_time.py cfg4 cfg5 > /dev/null 2>&1
python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
tail -c 1500 gpurun_out/bench_r02b.json
