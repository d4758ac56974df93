"""Reads a `.ncu-rep` (ncu --set full) with `ncu -i ... --page raw --csv` and writes the metrics the profiles/
summaries quote as kernel,metric,unit,value rows.  Usage: ncu_extract.py gpurun_out/x.ncu-rep profiles/x.metrics.csv"""
import csv, io, re, subprocess, sys
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__throughput\.avg\.pct|sm__inst_executed_pipe_(fp64|lsu|alu|fma|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct|sm__warps_active\.avg\.pct|launch__(registers_per_thread|shared_mem_per_block_dynamic|"
                  r"occupancy_limit_\w+|grid_size|block_size|cluster\w*)|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on\.sum|"
                  r"smsp__inst_executed\.sum|sm__cycles_elapsed\.max|l1tex__data_pipe_lsu_wavefronts(_mem_shared(_op_(ld|st))?)?\.sum(\.pct_of_peak_sustained_elapsed)?|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st|ldgsts))?\.sum|smsp__sass_inst_executed_op_shared(_ld|_st)?\.sum|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|memory_l1_wavefronts_shared(_ideal)?)$")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        for i, h in enumerate(hdr):
            if KEEP.match(h) and r[i] not in ("", "0", "0.000000"):
                w.writerow([r[ki][:60], h, units[i], r[i]])
