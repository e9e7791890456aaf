"""Phase clocks of stage1_kernel (profile build: make -C pysktb_b200/csrc TSPROF=1).  usage: ts_phase_profile.py N nm"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
_, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
lib = _lib.lib()
N, nm = int(sys.argv[1]), int(sys.argv[2])
M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda")
band = torch.zeros((nm, N, 32), dtype=torch.complex128, device="cuda")
buf = (C.c_ulonglong * 32)()
raw = C.CDLL(_lib.LIB_PATH)
for rep in range(2):
    lib.sktb_heev_twostage(ham._plan, M.data_ptr(), N, nm, band.data_ptr(), None, None)
    raw.sktb_ts_profile(buf)
v = np.array(list(buf), dtype=float)
names = ["top barrier (update imbalance)", "panel load", "QR", "band out", "Y = A V (+barrier)", "X = Y T", "G, S", "W", "update (warp 0)"]
tot = v[:9].sum() + v[12:17].sum()
for i, n in enumerate(names): print(f"{n:32s} {v[i]/nm/1e6:8.3f} Mclk/matrix  {100*v[i]/tot:5.1f}%")
print(f"  inside QR: dots {v[15]/nm/1e6:.3f} + barrier {v[12]/nm/1e6:.3f}, update {v[16]/nm/1e6:.3f} + barrier {v[13]/nm/1e6:.3f} Mclk/matrix (thread 0)")
print(f"total {tot/nm/1e6:.3f} Mclk/matrix; warp-busy Y = A V: {v[10]/16/nm/1e6:.3f}, update: {v[11]/16/nm/1e6:.3f} Mclk/matrix (mean over warps)")
