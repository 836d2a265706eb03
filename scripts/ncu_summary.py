"""Dev tool: condense an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / profiles/ cite.
usage: python scripts/ncu_summary.py <file.ncu-rep> <out.txt> "<header comment>" """
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# %s\n# source: %s\n' % (header, rep))
        for r in rows[2:]:
            d = dict(zip(names, zip(units, r)))
            f.write('\n## kernel: %s\n' % (d.get('Kernel Name', ('', '?'))[1]))
            for k in KEYS:
                if k in d:
                    f.write('%-88s %-12s %s\n' % (k, d[k][0], d[k][1]))
    print(open(out).read())


if __name__ == '__main__':
    main()
