"""Tiny driver for ncu: one workload, a few mesh solves.  usage: run_solve.py [workload] [n0 n1 n2] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "T"
mesh = [int(x) for x in sys.argv[2:5]] if len(sys.argv) >= 5 else [64, 64, 64]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
builder, kw, _ = WORKLOADS[name]
spec = W.ALL[builder](**kw)
_, ham = W.instantiate(spec, tb, tb.scaling)
for _ in range(reps):
    out = ham.solve_kmesh(mesh, soc=spec["soc"], check=False)
torch.cuda.synchronize()
print(ham.last_kernel_info(), float(out.sum()))
