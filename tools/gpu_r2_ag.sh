#!/bin/bash
# round 2, capture AG: ncu of the frustum kernel (camera rays of a 16-sample wave)
mkdir -p gpurun_out
BPT_PACKET=4 ncu --set full --clock-control none --import-source on -k regex:"k_extend_frustum" -s 1 -c 1 -o /tmp/r2ag python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_ae.log 2>&1
ncu -i /tmp/r2ag.ncu-rep --page raw --csv > gpurun_out/r2ag_raw.csv 2>> gpurun_out/ncu_ae.log
python tools/summarize_ncu.py source /tmp/r2ag.ncu-rep > gpurun_out/r2ag_source.md 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2ag_raw.csv')))
h=rows[0]; u=rows[1]; r=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for w in want:
    if w in h: i=h.index(w); print(w, u[i], r[i])
PY
head -45 gpurun_out/r2ag_source.md | cut -c1-210
