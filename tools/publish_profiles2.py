"""Copy the evidence of tools/collect_round2.sh from gpurun_out/ (scratch) into profiles/ (tracked): ncu summary, launch
lists, bench lines, secondary throughput logs, FP64 peak record, and the traffic record bench.py reads."""
import csv, io, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
order = ["pair_wide32", "pair_tail16", "tql_small", "cta64", "solve_thread", "solve_thread12", "hk", "generic512", "stein", "vecbt_big", "vecbt_big256", "rowsq_big",
         "tqlrec", "vecapply32", "vecbt32", "tqlrec_n64", "vecapply64_n64", "vecbt64_n64", "lorentz", "n2"]
reps = [os.path.join("gpurun_out", f"r2_full_{k}.raw.csv") for k in order if os.path.exists(os.path.join(O, f"r2_full_{k}.raw.csv")) and os.path.getsize(os.path.join(O, f"r2_full_{k}.raw.csv")) > 1000]
subprocess.run([sys.executable, os.path.join(R, "tools", "make_profile_summary.py"), os.path.join(P, "r02_ncu_summary.md"),
                "ncu --set full summaries, round 2 (T: N=32, 64^3 = 262144 k per launch; C4: N=64, 32^3 with K0+K1 fused; C1: N=8, 128^3; C3: N=12, 1024^2; hk: C4 "
                "materialised; generic / stein / vecbt_big: N=512, 148-296 matrices; vecbt_big256: N=256; rowsq_big: C5 edge rows; vec kernels: 65536 k (N=32) / "
                "8192 k (N=64); lorentz: C2 2048^2 x 400 E; n2: 4096^2)"] + reps, cwd=R, check=True)
copies = [("r2_launches_T.csv", "r02_launches_T_bench256.csv"), ("r2_launches_C4.csv", "r02_launches_C4.csv"), ("r2_launches_C5ldos.csv", "r02_launches_C5_edge_ldos.csv"),
          ("r2_bench_T.json", "r02_bench_T_1gpu.json"), ("r2_bench_T_reference.json", "r02_bench_T_reference_arm.json"), ("r2_secondary.log", "r02_secondary_throughput.log"),
          ("r2_nscan.log", "r02_heev_size_scan.log"), ("r2_peaks_fp64.json", "peaks_fp64.json"), ("r2_pytest.log", "r02_pytest_gpu.log"), ("r2_smoke.log", "r02_smoke.log"),
          ("r2_host.txt", "r02_host.txt")] + [(f"r2_bench_{w}.json", f"r02_bench_{w}_1gpu.json") for w in ("C1", "C2", "C3", "C4", "C5")]
for src, dst in copies:
    if os.path.exists(os.path.join(O, src)) and os.path.getsize(os.path.join(O, src)) > 0:
        shutil.copy(os.path.join(O, src), os.path.join(P, dst))
def dram(rep):
    rows = list(csv.reader(io.StringIO(open(rep).read()))); hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    g = lambda k: float(vals[hdr.index(k)]) * scale[units[hdr.index(k)]]
    return g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
tj = {}
for key, rep, n, nk, note in (("pair_wide32", "r2_full_pair_wide32.raw.csv", 32, 262144, "algorithmic: 512 B (d,e) + 4096 B trailing block per k = 1.208e9 B"),
                              ("tridiag_cta64", "r2_full_cta64.raw.csv", 64, 32768, "K0+K1 fused: no H(k) read; algorithmic: 1 KB (d,e) + 16 KB trailing 32x32 block out per k"),
                              ("solve_thread", "r2_full_solve_thread.raw.csv", 8, 2097152, "algorithmic: 128 B (d,e) out per k"),
                              ("heev_generic", "r2_full_generic512.raw.csv", 512, 296, "panel formulation, NB = 8: 16 B read per element and step + 32/8 B read+write = 0.90 GB per matrix")):
    f = os.path.join(O, rep)
    if os.path.exists(f) and os.path.getsize(f) > 1000:
        r, w = dram(f)
        tj[key] = {"matrix_size": n, "k_points_per_launch": nk, "dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
                   "source": f"profiles/r02_ncu_summary.md (ncu --set full, one launch of {nk} k-points; {note})"}
if tj:
    json.dump(tj, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
print("published", sorted(f for f in os.listdir(P) if f.startswith("r02") or f.startswith("peaks")))
