"""Summarise an `ncu --page raw --csv` dump: python scripts/ncu_summary.py raw.csv [pattern ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
default = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
           'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
           'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64',
           'sm__pipe_fp64_cycles_active', 'smsp__issue_active.avg.pct', 'lts__t_sector_hit_rate.pct',
           'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
           'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
           'sm__cycles_elapsed.avg', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'sm__inst_executed_pipe_',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'shared_mem', 'smsp__thread_inst_executed_per_inst_executed']
pats = sys.argv[2:] or default
for i, h in enumerate(hdr):
    if any(p in h for p in pats):
        print("%-90s %-12s %s" % (h, units[i], [r[i] for r in data]))
