import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
spec = W.zigzag_ribbon(n_chains=256)
_, ham = W.instantiate(spec, tb, tb.scaling)
N = 512
lib = _lib.lib()
for nk in (64, 200, 592):
    k = torch.from_numpy(np.random.default_rng(0).random((nk, 3))).cuda()
    ev = torch.empty((N, nk), dtype=torch.float64, device="cuda")
    vec = torch.empty((N, nk, N), dtype=torch.complex128, device="cuda")
    for rep in range(2):
        vec.zero_()
        _lib.check(lib.sktb_solve(ham._plan, k.data_ptr(), nk, 0, ev.data_ptr(), vec.data_ptr(), nk, 0, None, None)); torch.cuda.synchronize()
        H = torch.empty((nk, N, N), dtype=torch.complex128, device="cuda")
        _lib.check(lib.sktb_hk(ham._plan, k.data_ptr(), nk, 0, H.data_ptr(), None, None)); torch.cuda.synchronize()
        L = torch.tril(H, -1); A = L + L.conj().transpose(1, 2) + torch.diag_embed(torch.diagonal(H, dim1=1, dim2=2).real.to(torch.complex128))
        U = vec.permute(1, 2, 0)                                   # [k][comp][band]
        R = (A @ U - U * ev.T[:, None, :]).abs().amax(dim=(1, 2)).cpu().numpy()
        O = (U.conj().transpose(1, 2) @ U - torch.eye(N, device="cuda", dtype=torch.complex128)).abs().amax(dim=(1, 2)).cpu().numpy()
        bad = np.nonzero((R > 1e-10) | (O > 1e-10))[0]
        print(f"nk={nk} rep={rep}: max residual {R.max():.2e} orth {O.max():.2e}; {len(bad)} bad k; first bad {bad[:10]} kx {k[bad[:10], 0].cpu().numpy()}")
        if len(bad):
            q = int(bad[0])
            colnorm = U[q].abs().pow(2).sum(dim=0).cpu().numpy()
            print("   column norms of first bad k: min %.3e max %.3e; #zero cols %d" % (colnorm.min(), colnorm.max(), (colnorm < 1e-20).sum()))
