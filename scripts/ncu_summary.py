"""Key metrics of every kernel in an .ncu-rep (run here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'local_load', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    for row in rows[2:]:
        rec = dict(zip(header, row))
        print('====', rec.get('Kernel Name', '')[:100])
        for key in KEYS:
            if key in rec:
                print('  %-85s %s %s' % (key, rec[key], units[header.index(key)]))


if __name__ == '__main__':
    main(sys.argv[1])
