"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py rep [pattern ...]"""
import csv
import subprocess
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
           'launch__occupancy_limit', 'launch__shared_mem_per_block', 'launch__grid_size',
           'launch__block_size', 'sm__throughput.avg.pct', 'smsp__inst_executed.sum',
           'sm__inst_executed_pipe_fp64', 'op_dfma', 'op_dmul', 'op_dadd',
           'smsp__issue_active.avg.pct', 'smsp__warp_issue_stalled', 'smsp__average_warp',
           'smsp__thread_inst_executed_per_inst_executed', 'pipe_fp64', 'pipe_xu', 'pipe_lsu',
           'pipe_alu', 'pipe_fma', 'l1tex__data_bank_conflicts', 'smsp__inst_executed_op_shared',
           'smsp__inst_executed_op_global', 'smsp__inst_executed_op_local',
           'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'gpu__dram_throughput',
           'lts__t_sector_hit_rate', 'l1tex__t_sector_hit_rate', 'smsp__warps_eligible',
           'derived__smsp__sass_thread_inst', 'local_load', 'local_store']


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[4][:90], 'grid', r[8], 'block', r[7])
        for h, u, v in zip(hdr, unit, r):
            if any(p in h for p in pats):
                if v in ('', '0', 'n/a'):
                    continue
                print('  %-95s %-14s %s' % (h.split('.', 2)[-1] if h.count('.') > 2 else h, u, v))


if __name__ == '__main__':
    main()
