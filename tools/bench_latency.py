"""Wall-clock latency of small solve_kpath calls through the public Python API (band-structure use case)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
for name, builder, path, nk in (("C1", W.cubic_sp3, [[0, 0, 0], [0.5, 0.5, 0.5]], 50), ("T", W.ce_spdf_1atom, [[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0]], 50),
                                ("C4", W.ce_spdf_2atom, [[0, 0, 0], [0.5, 0, 0]], 50), ("C3", W.sb_buckled, [[0, 0, 0], [0.5, 0, 0], [1 / 3, 1 / 3, 0]], 100)):
    spec = builder()
    t0 = time.perf_counter(); _, ham = W.instantiate(spec, tb, tb.scaling); torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    kpts, dist, spl = ham.get_kpts(path, nk)
    for vec in (False, True):
        ham.solve_kpath(kpts, eig_vectors=vec)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); r = ham.solve_kpath(kpts, eig_vectors=vec); ts.append(time.perf_counter() - t0)
        print(f"{name}: {len(kpts)} k-points, vectors={vec}: median {np.median(ts)*1e3:.3f} ms, min {min(ts)*1e3:.3f} ms (model build {t_build*1e3:.1f} ms)")
