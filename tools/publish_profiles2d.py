"""Copy the evidence of tools/collect_round2d.sh (final state of round 2) from gpurun_out/ (scratch) into profiles/ (tracked)."""
import csv, os, re, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
order = ["pair_wide32", "pair_tail16", "tql_small", "cta64"]
reps = [os.path.join("gpurun_out", f"r2d_full_{k}.raw.csv") for k in order if os.path.exists(os.path.join(O, f"r2d_full_{k}.raw.csv"))]
subprocess.run([sys.executable, os.path.join(R, "tools", "make_profile_summary.py"), os.path.join(P, "r02d_ncu_summary.md"),
                "ncu --set full summaries, round 2 final state (after the FP64-issue diet of the scalar chains): the three kernels of the T step "
                "(N = 32, 64^3 = 262144 k per launch) and tridiag_cta64 (C4, 32^3)"] + reps, cwd=R, check=True)
# FP64 instructions per matrix of the headline kernel, from the source page of the same capture
rows = list(csv.reader(open(os.path.join(O, "r2d_src_pair_wide32.csv"))))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
cnt = {}
for r in rows[2:]:
    op = re.sub(r'^@!?U?P\w+\s+', '', r[col['Source']].strip()).split()[0].split('.')[0]
    cnt[op] = cnt.get(op, 0) + int(r[col['Instructions Executed']] or 0)
fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
with open(os.path.join(P, "r02d_ncu_summary.md"), "a") as f:
    f.write("\n## FP64 warp instructions of pair_wide32 per matrix (source page of the capture above, 262144 matrices)\n\n")
    f.write("| opcode | executed | per matrix |\n|---|---:|---:|\n")
    for k in ("DFMA", "DMUL", "DADD", "DSETP"):
        f.write(f"| {k} | {cnt.get(k, 0)} | {cnt.get(k, 0) / 262144:.1f} |\n")
    f.write(f"| all FP64 | {fp64} | **{fp64 / 262144:.0f}** (round 1: 8 700, round 2 second session: 6 740) |\n")
    f.write(f"| all instructions | {sum(cnt.values())} | {sum(cnt.values()) / 262144:.0f} |\n")
copies = [("r2d_launches_T.csv", "r02_launches_T_bench256.csv"), ("r2d_bench_T.json", "r02_bench_T_1gpu.json"), ("r2d_bench_T_reference.json", "r02_bench_T_reference_arm.json"),
          ("r2d_secondary.log", "r02d_secondary_throughput.log"), ("r2d_nscan.log", "r02d_heev_size_scan.log"), ("r2d_pytest.log", "r02_pytest_gpu.log"), ("r2d_smoke.log", "r02_smoke.log")]
copies += [(f"r2d_bench_{w}.json", f"r02_bench_{w}_1gpu.json") for w in ("C1", "C2", "C3", "C4", "C5")]
for src, dst in copies:
    if os.path.exists(os.path.join(O, src)) and os.path.getsize(os.path.join(O, src)) > 0:
        shutil.copy(os.path.join(O, src), os.path.join(P, dst))
print("published", sorted(x for x in os.listdir(P) if x.startswith("r02d")), "FP64 per matrix", fp64 / 262144)
