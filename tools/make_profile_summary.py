"""Summarise .ncu-rep captures into profiles/<name>.md.  usage: make_profile_summary.py out.md title rep1 [rep2 ...]"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__icc_request_hit_rate.pct', 'smsp__inst_executed.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum']
out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
lines = [f"# {title}", ""]
for rep in reps:
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else rep
    lines += [f"## {kname}", "", f"source: `{rep}` (raw metric table of one `ncu --set full` launch; scratch, not committed)", "", "| metric | value | unit |", "|---|---:|---|"]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h and v and float(v) > 0.05):
            lines.append(f"| {h} | {v} | {u} |")
    lines.append("")
open(out, "w").write("\n".join(lines))
print("wrote", out)
