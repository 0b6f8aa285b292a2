"""Compact per-kernel summary of an .ncu-rep (ncu --set full): the metrics the design
discussion cites, one CSV row per metric and launch.  Usage:
    python scripts/ncu_summary.py report.ncu-rep out.csv [note]"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_write.sum.per_second',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
    'lts__t_sectors_srcunit_tex_op_write.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'kernel', 'metric', 'unit', 'value'])
        if note:
            w.writerow(['', '', 'note', '', note])
        for k, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            for m in WANT:
                if m in d:
                    w.writerow([k, d.get('Kernel Name', '')[:80], m, u[m], d[m]])


if __name__ == '__main__':
    main()
