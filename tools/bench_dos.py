"""Time GreensFunction.dos (K1+K2+K3, or K1+K5 with method=inverse) for a workload.  usage: bench_dos.py [C2|C3|T|C4] [n0 n1 n2] [nE] [dos|ldos] [eigh|inverse|auto]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
mesh = [int(x) for x in sys.argv[2:5]] if len(sys.argv) >= 5 else [2048, 2048, 1]
nE = int(sys.argv[5]) if len(sys.argv) > 5 else 400
ldos = len(sys.argv) > 6 and sys.argv[6] == "ldos"
builder, kw, _ = WORKLOADS[name]
spec = W.ALL[builder](**kw)
_, ham = W.instantiate(spec, tb, tb.scaling)
gf = tb.GreensFunction(ham, method=sys.argv[7] if len(sys.argv) > 7 else "eigh")
E = np.linspace(-8, 8, nE)
N = len([o for _, _, orbs in spec["atoms"] for o in orbs]) * (2 if spec["soc"] else 1)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = gf.ldos(E, atom_indices=[0], nk=mesh, eta=0.1, soc=spec["soc"]).sum(axis=0) if ldos else gf.dos(E, nk=mesh, eta=0.1, soc=spec["soc"])
    torch.cuda.synchronize(); t = time.perf_counter() - t0
nk = mesh[0] * mesh[1] * mesh[2]
print(("LDOS " if ldos else "DOS ") + f"{name} N={N} mesh={mesh} nE={nE}: {t*1e3:.2f} ms, {nk/t:.3e} k/s, {nk*N*nE/t:.3e} Lorentzians/s, integral={np.trapezoid(d, E):.6f}  [{ham.last_kernel_info()[:50]}]")
