import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spec = W.zigzag_ribbon(n_chains=nch)
_, ham = W.instantiate(spec, tb, tb.scaling)
N = 2 * nch
lib = _lib.lib()
k = np.array([[0.37, 0, 0], [0.5, 0, 0], [0.0, 0, 0], [0.123, 0.3, 0.9], [0.71, 0, 0]])
H = ham.get_ham_batch(k, l_soc=False)
def check(tag, ev, U):
    for q in range(len(k)):
        L = np.tril(H[q], -1); A = L + L.conj().T + np.diag(np.diag(H[q]).real)
        Uq = U[:, q, :].T
        ref = np.linalg.eigvalsh(A)
        print(f"{tag} k={k[q][0]:.3f} eig err {np.abs(ev[:, q]-ref).max():.1e} residual {np.abs(A @ Uq - Uq * ev[:, q][None, :]).max():.1e} orth {np.abs(Uq.conj().T @ Uq - np.eye(N)).max():.1e}")
Md = torch.from_numpy(np.ascontiguousarray(H)).cuda()
ev = torch.empty((N, len(k)), dtype=torch.float64, device="cuda"); U = torch.empty((N, len(k), N), dtype=torch.complex128, device="cuda")
_lib.check(lib.sktb_heev(ham._plan, Md.data_ptr(), N, len(k), ev.data_ptr(), U.data_ptr(), None, None)); torch.cuda.synchronize()
check("heev ", ev.cpu().numpy(), U.cpu().numpy())
v, u = ham.solve_kpath(list(k), eig_vectors=True, soc=False)
check("solve", v, u)
print(ham.last_kernel_info())
