import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.linalg import hessenberg
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
spec = W.zigzag_ribbon(n_chains=256)
_, ham = W.instantiate(spec, tb, tb.scaling)
N = 512
lib = _lib.lib()
k = np.array([[0.6066, 0, 0], [0.45, 0, 0], [0.4858, 0, 0]])
H = ham.get_ham_batch(k, l_soc=False)
mats = []
for q in range(len(k)):
    L = np.tril(H[q], -1); A = L + L.conj().T + np.diag(np.diag(H[q]).real)
    T, Q = hessenberg(A, calc_q=True)
    d, e = np.diag(T).real, np.abs(np.diag(T, -1))
    mats += [A, (np.diag(d) + np.diag(e, 1) + np.diag(e, -1)).astype(complex), (np.diag(d) + np.diag(-e, 1) + np.diag(-e, -1)).astype(complex)]
M = np.array(mats)
Md = torch.from_numpy(np.ascontiguousarray(M)).cuda()
nm = len(M)
ev = torch.empty((N, nm), dtype=torch.float64, device="cuda"); U = torch.empty((N, nm, N), dtype=torch.complex128, device="cuda")
_lib.check(lib.sktb_heev(ham._plan, Md.data_ptr(), N, nm, ev.data_ptr(), U.data_ptr(), None, None)); torch.cuda.synchronize()
ev, U = ev.cpu().numpy(), U.cpu().numpy()
for q in range(nm):
    A = M[q]; Uq = U[:, q, :].T
    R = np.abs(A @ Uq - Uq * ev[:, q][None, :]).max(axis=0); O = np.abs(Uq.conj().T @ Uq - np.eye(N))
    bad = np.nonzero(R > 1e-10)[0]
    print(["dense H(k)", "T(+e)", "T(-e)"][q % 3], "k=%.4f" % k[q // 3][0], "residual %.2e orth %.2e" % (R.max(), O.max()), "bad cols", bad[:12], "ev there", ev[bad[:4], q],
          "norms", np.linalg.norm(Uq[:, bad[:4]], axis=0))
