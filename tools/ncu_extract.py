#!/usr/bin/env python3
"""Key-metric extract of an `ncu --set full` report: one CSV row per captured launch.

    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv > profiles/<name>.csv
"""
import csv
import sys

COLS = """Kernel Name
gpu__time_duration.sum
launch__registers_per_thread
launch__block_size
launch__grid_size
launch__waves_per_multiprocessor
sm__warps_active.avg.pct_of_peak_sustained_active
sm__inst_issued.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio""".split("\n")


def main(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    cols = [c for c in COLS if c in h]
    w = csv.writer(sys.stdout)
    w.writerow(cols)
    w.writerow([units[h.index(c)] for c in cols])
    for r in rows[2:]:
        if len(r) == len(h):
            w.writerow([r[h.index(c)] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1])
