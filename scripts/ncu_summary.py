"""Summarise an `ncu --page raw --csv` dump: python scripts/ncu_summary.py raw.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
pats = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit', 'launch__waves', 'launch__shared_mem_per_block', 'sm__warps_active.avg.pct',
        'smsp__thread_inst_executed_per_inst_executed', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum$',
        r'dram__bytes_(read|write).sum$', 'pipe_fma.*pct_of_peak_sustained_active', 'pipe_alu.*pct_of_peak_sustained_active',
        'pipe_fp64.*pct_of_peak_sustained_active', 'pipe_xu.*pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu', 'bank_conflicts_pipe_lsu_mem_shared.sum', 'data_pipe_lsu_wavefronts_mem_shared.sum$',
        'data_pipe_lsu_wavefronts.sum$', r'sass_thread_inst_executed_op_f(fma|add|mul)_pred_on.sum$',
        'smsp__average_warps_issue_stalled.*_per_issue_active', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'lts__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'smsp__cycles_active.avg$', 'l1tex__data_pipe_lsu_wavefronts.avg.pct',
        'lts__t_bytes.sum$', 'sm__sass_inst_executed_op_shared', 'smsp__inst_executed_op_shared', 'l1tex__t_bytes.*global.*sum$']
for p in pats:
    for h in hdr:
        if re.search(p, h):
            vals = [d[idx[h]][:60] for d in data]
            print(f"{h[:95]:95s} {units[idx[h]][:14]:14s} {vals}")
