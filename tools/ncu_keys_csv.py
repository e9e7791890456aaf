"""Print the key metrics of a raw-page CSV (ncu -i rep --page raw --csv), first kernel.  usage: ncu_keys_csv.py file.raw.csv [extra-substring ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__icc_request_hit_rate.pct', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed_pipe_fp64.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']
for h, u, v in zip(hdr, units, vals):
    try: fv = float(v.replace(',', ''))
    except Exception: fv = 0
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and fv > 0.05) or any(e in h for e in sys.argv[2:]):
        print(f"{h} [{u}] {v}")
