"""Print the ncu raw-page metrics we track, per profiled kernel launch (used to write profiles/*.md)."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__cycles_elapsed.max', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in WANT:
        if w in idx:
            print(f"{w:85s} {r[idx[w]]:>22s} {units[idx[w]]}")
