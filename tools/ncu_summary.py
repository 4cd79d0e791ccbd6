#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed).
  python tools/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xxx.txt"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_shared_mem', 'occ_limit_smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_active %'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_hmma_active %'),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor inst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active %'),
    ('smsp__inst_executed.sum', 'warp inst'),
    ('sm__inst_executed_pipe_xu.sum', 'xu (MUFU) inst'),
    ('sm__inst_executed_pipe_fma.sum', 'fma inst'),
    ('sm__inst_executed_pipe_alu.sum', 'alu inst'),
    ('sm__inst_executed_pipe_lsu.sum', 'lsu inst'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput %'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'stall long_scoreboard %'),
    ('smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'stall short_scoreboard %'),
    ('smsp__warp_issue_stalled_wait_per_warp_active.pct', 'stall wait %'),
    ('smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'stall barrier %'),
    ('smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'stall math_throttle %'),
    ('smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'stall mio_throttle %'),
    ('smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'stall no_instruction %'),
    ('smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'stall not_selected %'),
    ('smsp__warp_issue_stalled_sleeping_per_warp_active.pct', 'stall sleeping %'),
    ('smsp__warp_issue_stalled_membar_per_warp_active.pct', 'stall membar %'),
    ('smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'stall lg_throttle %'),
    ('smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct', 'stall tex_throttle %'),
    ('smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'stall dispatch %'),
    ('smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'stall branch %'),
    ('smsp__warp_issue_stalled_selected_per_warp_active.pct', 'selected %'),
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        if pat and not pat.search(name):
            continue
        print("== %s" % name.split('(')[0])
        for m, label in METRICS:
            if m in col:
                print("   %-28s %s %s" % (label, r[col[m]], units[col[m]]))
        print()


if __name__ == '__main__':
    main()
