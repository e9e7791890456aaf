"""Print the key metrics of an .ncu-rep (first kernel).  usage: ncu_keys.py report.ncu-rep [extra-substring ...]"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__icc_request_hit_rate.pct', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'launch__grid_size', 'launch__block_size',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'smsp__inst_executed_pipe_fp64.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
extra = sys.argv[2:]
for h, u, v in zip(hdr, units, vals):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.02) or any(e in h for e in extra):
        print(f"{h} [{u}] {v}")
