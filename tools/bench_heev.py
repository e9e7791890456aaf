"""Batched sktb_heev throughput on random Hermitian matrices.  usage: bench_heev.py N [nm] [vec]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
_, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
lib = _lib.lib()
for N in [int(x) for x in sys.argv[1].split(",")]:
    nm = int(sys.argv[2]) if len(sys.argv) > 2 else max(296, min(65536, (1 << 28) // (N * N * 16)))
    vec = len(sys.argv) > 3 and sys.argv[3] == "vec"
    M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda")
    M = M + M.conj().transpose(1, 2)
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    U = torch.empty((N, nm, N), dtype=torch.complex128, device="cuda") if vec else None
    def run():
        _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nm, ev.data_ptr(), U.data_ptr() if vec else None, None, None))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    flops = (16.0 if vec else 16.0 / 3) * N ** 3
    ref = np.linalg.eigvalsh(M[0].cpu().numpy())
    print(f"N={N} nm={nm} vec={vec}: {ms:.2f} ms  {nm/ms*1e3:.3e} matrices/s  {flops*nm/ms/1e9:.2f} TFLOP/s(conv)  err {np.abs(ev[:,0].cpu().numpy()-ref).max():.1e}  [{ham.last_kernel_info()[:80]}]")
