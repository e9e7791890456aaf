"""Residual / orthogonality of sktb_heev eigenvectors for N > 64: dense random Hermitian and real tridiagonal inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
_, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
lib = _lib.lib()
for N in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "96,256,257,288,320,384,512").split(",")]:
    r = np.random.default_rng(N)
    nm = 3
    M = r.standard_normal((nm, N, N)) + 1j * r.standard_normal((nm, N, N))
    M = M + M.conj().transpose(0, 2, 1)
    d, e = r.standard_normal(N), r.standard_normal(N - 1)
    M[1] = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)              # real tridiagonal: trivial reflectors, U = Z
    M[2] = np.kron(M[2][: N // 2, : N // 2], np.eye(2))[:N, :N] if N % 2 == 0 else M[2]
    Md = torch.from_numpy(np.ascontiguousarray(M)).cuda()
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    U = torch.empty((N, nm, N), dtype=torch.complex128, device="cuda")
    _lib.check(lib.sktb_heev(ham._plan, Md.data_ptr(), N, nm, ev.data_ptr(), U.data_ptr(), None, None))
    torch.cuda.synchronize()
    evh, Uh = ev.cpu().numpy(), U.cpu().numpy()
    for q, tag in enumerate(("dense", "tridiagonal", "kron-degenerate")):
        A = M[q]
        Uq = Uh[:, q, :].T
        ref = np.linalg.eigvalsh(A)
        print(f"N={N} {tag:16s} eig err {np.abs(evh[:, q] - ref).max():.1e}  residual {np.abs(A @ Uq - Uq * evh[:, q][None, :]).max():.1e}  "
              f"orth {np.abs(Uq.conj().T @ Uq - np.eye(N)).max():.1e}")
print(ham.last_kernel_info())
