"""Copy the evidence of tools/collect_round.sh from gpurun_out/ (scratch) into profiles/ (tracked): ncu summary,
launch list, bench lines, secondary throughput log, and the traffic record bench.py reads."""
import csv, io, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
order = ["pair_wide32", "pair_tail16", "tql_small", "cta64", "tqlrec", "vecapply32", "vecbt32", "tqlrec_n64", "vecapply64_n64", "vecbt64_n64", "generic512", "lorentz", "n2"]
reps = [os.path.join("gpurun_out", f"final_full_{k}.raw.csv") for k in order if os.path.exists(os.path.join(O, f"final_full_{k}.raw.csv"))]
subprocess.run([sys.executable, os.path.join(R, "tools", "make_profile_summary.py"), os.path.join(P, "r01_ncu_summary_final.md"),
                "ncu --set full summaries, round 1 final (T: N=32, 64^3 = 262144 k per launch; C4: N=64, 32^3; vec kernels: 65536 k (N=32) / 8192 k (N=64); "
                "generic: N=512, 296 matrices; lorentz: C2 2048^2 x 400 E; n2: 4096^2)"] + reps, cwd=R, check=True)
for src, dst in [("final_launches_T.csv", "r01_launches_T_bench256.csv"), ("final_bench_T.json", "r01_bench_T_1gpu_final.json"),
                 ("final_bench_T_reference.json", "r01_bench_T_reference_arm.json"), ("final_secondary.log", "r01_secondary_throughput.log"),
                 ("final_nscan.log", "r01_heev_size_scan.log")] + \
                [(f"final_bench_{w}.json", f"r01_bench_{w}_1gpu.json") for w in ("C1", "C2", "C3", "C4", "C5")]:
    if os.path.exists(os.path.join(O, src)):
        shutil.copy(os.path.join(O, src), os.path.join(P, dst))
def dram(rep):
    raw = open(rep).read()
    rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    g = lambda k: float(vals[hdr.index(k)]) * scale[units[hdr.index(k)]]
    return g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
tj = {}
for key, rep, n, nk, note in (("pair_wide32", "final_full_pair_wide32.raw.csv", 32, 262144, "algorithmic: 512 B (d,e) + 4096 B trailing block per k = 1.208e9 B"),
                              ("tridiag_cta64", "final_full_cta64.raw.csv", 64, 32768, "algorithmic: 64 KB H(k) in (lower triangle 32 KB) + 1 KB (d,e) out per k"),
                              ("heev_generic", "final_full_generic512.raw.csv", 512, 296, "panel formulation, NB = 8: 16 B read per element and step + 32/8 B read+write = 0.90 GB per matrix")):
    f = os.path.join(O, rep)
    if os.path.exists(f):
        r, w = dram(f)
        tj[key] = {"matrix_size": n, "k_points_per_launch": nk, "dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
                   "source": f"profiles/r01_ncu_summary_final.md (ncu --set full, one launch of {nk} k-points; {note})"}
if tj:
    json.dump(tj, open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
print("published", sorted(os.listdir(P)))
