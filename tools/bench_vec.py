"""Device-resident eigenvalues + eigenvectors through sktb_solve (C ABI).  usage: bench_vec.py [workload] [nk]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "T"
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
builder, kw, _ = WORKLOADS[name]
spec = W.ALL[builder](**kw)
_, ham = W.instantiate(spec, tb, tb.scaling)
N = len([o for _, _, orbs in spec["atoms"] for o in orbs]) * (2 if spec["soc"] else 1)
k = torch.from_numpy(np.random.default_rng(0).random((nk, 3))).cuda()
ev = torch.empty((N, nk), dtype=torch.float64, device="cuda")
vec = torch.empty((N, nk, N), dtype=torch.complex128, device="cuda")
lib = _lib.lib()
def run(with_vec):
    _lib.check(lib.sktb_solve(ham._plan, k.data_ptr(), nk, int(spec["soc"]), ev.data_ptr(), vec.data_ptr() if with_vec else None, nk, 0, None, None))
for with_vec in (False, True):
    run(with_vec); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run(with_vec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name} N={N} nk={nk} vectors={with_vec}: {ms:.2f} ms  {nk/ms*1e3:.3e} k/s  [{ham.last_kernel_info()[:70]}]")
# correctness on a sample: residual and orthogonality against the device-built H(k)
H = torch.empty((64, N, N), dtype=torch.complex128, device="cuda")
_lib.check(lib.sktb_hk(ham._plan, k.data_ptr(), 64, int(spec["soc"]), H.data_ptr(), None, None))
torch.cuda.synchronize()
Hh = H.cpu().numpy(); U = vec[:, :64, :].cpu().numpy(); E = ev[:, :64].cpu().numpy()
res = orth = 0.0
for q in range(64):
    Hl = np.tril(Hh[q]); Hm = Hl + Hl.conj().T - np.diag(np.diag(Hh[q]).real) - np.diag(np.diag(Hl)) + np.diag(np.diag(Hh[q]).real)
    Uq = U[:, q, :]                        # rows = eigenvectors
    res = max(res, np.abs(Hm @ Uq.T - Uq.T * E[:, q][None, :]).max())
    orth = max(orth, np.abs(Uq.conj() @ Uq.T - np.eye(N)).max())
    assert np.abs(np.linalg.eigvalsh(Hm) - E[:, q]).max() < 1e-10
print(f"max residual {res:.2e}, max |U U^H - I| {orth:.2e}")
