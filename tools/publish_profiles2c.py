"""Copy the evidence of tools/collect_round2c.sh (third session of round 2: dense-inverse Green's function) from gpurun_out/ (scratch) into
profiles/ (tracked): ncu summary of the K5 kernels, launch lists, bench lines, throughput log, GPU test / smoke logs."""
import os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
order = ["gf_warp", "gf_tile4", "gf_cta", "gf_smem"]
reps = [os.path.join("gpurun_out", f"r2c_full_{k}.raw.csv") for k in order
        if os.path.exists(os.path.join(O, f"r2c_full_{k}.raw.csv")) and os.path.getsize(os.path.join(O, f"r2c_full_{k}.raw.csv")) > 1000]
subprocess.run([sys.executable, os.path.join(R, "tools", "make_profile_summary.py"), os.path.join(P, "r02c_ncu_summary.md"),
                "ncu --set full summaries, round 2, dense-inverse Green's function (K5): gf_inverse_warp on T (N = 32, 20^3 k x 400 E = 3.2e6 inverses per launch), "
                "gf_inverse_tile<4> / gf_inverse_cta<8,2> on C4 (N = 64, 10^3 k x 400 E = 4e5 inverses), gf_inverse_kernel (shared-memory layout) on T"] + reps, cwd=R, check=True)
copies = [("r2c_launches_gf_T.csv", "r02c_launches_gf_T.csv"), ("r2c_launches_gf_C4.csv", "r02c_launches_gf_C4.csv"),
          ("r2c_bench_T.json", "r02_bench_T_1gpu.json"), ("r2c_bench_T_reference.json", "r02_bench_T_reference_arm.json"),
          ("r2c_bench_C4.json", "r02_bench_C4_1gpu.json"), ("r2c_bench_C5.json", "r02_bench_C5_1gpu.json"),
          ("r2c_gf_throughput.log", "r02c_gf_throughput.log"), ("r2c_pytest.log", "r02_pytest_gpu.log"), ("r2c_smoke.log", "r02_smoke.log")]
for src, dst in copies:
    if os.path.exists(os.path.join(O, src)) and os.path.getsize(os.path.join(O, src)) > 0:
        shutil.copy(os.path.join(O, src), os.path.join(P, dst))
print("published", sorted(x for x in os.listdir(P) if x.startswith("r02c")))
