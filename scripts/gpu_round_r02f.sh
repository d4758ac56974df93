#!/bin/bash
# one GPU session at the end of round 2: full gpu tests, both bench arms, the ncu launch list of the bench command,
# one `--set full` capture of the headline kernel and the counters of the tile-mode phase kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/gputests_r02f.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02f_ref.json 2> gpurun_out/bench_r02f_ref.err
python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02f.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ba_window_cluster -s 3 -c 1 -f -o gpurun_out/ba_cluster_r02f \
    python bench.py --steps 1 --warmup 3 --no-extra > gpurun_out/ncu_full.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"k_lg_lin|k_lg_backsub|k_bcr_chol|k_bcr_update" -c 16 --csv --log-file gpurun_out/lg_r02f.csv \
    python scripts/lg_time.py cfg5 > /dev/null 2>&1
tail -c 600 gpurun_out/bench_r02f.json; cat gpurun_out/gputests_r02f.txt | tail -3
