"""Two-stage tridiagonalisation on the GPU against LAPACK, stage by stage.  usage: test_twostage.py [N,N,...] [nm]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
_, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
lib = _lib.lib()
Ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 40, 64, 97, 130, 256, 512]
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def band_to_dense(AB):
    N, w = AB.shape
    A = np.zeros((N, N), complex)
    for j in range(N):
        n = min(w, N - j)
        A[j:j + n, j] = AB[j, :n]
    return np.tril(A) + np.tril(A, -1).conj().T


for N in Ns:
    torch.manual_seed(N)
    M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda")     # NOT Hermitian: only the lower triangle may be read
    band = torch.zeros((nm, N, 32), dtype=torch.complex128, device="cuda")
    de = torch.zeros((nm, 2, N), dtype=torch.float64, device="cuda")
    t0 = time.time()
    rc = lib.sktb_heev_twostage(ham._plan, M.data_ptr(), N, nm, band.data_ptr(), de.data_ptr(), None)
    torch.cuda.synchronize()
    if rc:
        print(N, "rc", rc, _lib.lib().sktb_last_error().decode()); continue
    Mh, Bh, Dh = M.cpu().numpy(), band.cpu().numpy(), de.cpu().numpy()
    e1 = e2 = 0.0
    for i in range(min(nm, 3)):
        ref = np.linalg.eigvalsh(Mh[i], UPLO="L")
        b = np.linalg.eigvalsh(band_to_dense(Bh[i]))
        d, e = Dh[i, 0], Dh[i, 1, :N - 1]
        t = np.linalg.eigvalsh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
        e1 = max(e1, np.abs(b - ref).max()); e2 = max(e2, np.abs(t - ref).max())
    wide = np.abs(Bh[:, :, 17:]).max()
    print(f"N={N} nm={nm}: stage1 err {e1:.2e}  stage2 err {e2:.2e}  fill beyond d=16 after stage 1: {wide:.1e}  ({time.time()-t0:.2f} s)", flush=True)
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    os.environ["SKTB_TWOSTAGE_MIN"] = "32"
    _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nm, ev.data_ptr(), None, None, None))
    torch.cuda.synchronize()
    ref = np.linalg.eigvalsh(Mh[nm - 1], UPLO="L")
    print(f"     sktb_heev: err {np.abs(ev[:, nm - 1].cpu().numpy() - ref).max():.2e}  [{ham.last_kernel_info()[:100]}]", flush=True)

# timing: stage 1 alone, stage 1 + 2, whole sktb_heev (two-stage) - nm = 296 unless given
if os.environ.get("TS_TIME"):
    for N in Ns:
        nmt = int(os.environ["TS_TIME"])
        M = torch.randn((nmt, N, N), dtype=torch.complex128, device="cuda")
        de = torch.zeros((nmt, 2, N), dtype=torch.float64, device="cuda")
        band = torch.zeros((nmt, N, 32), dtype=torch.complex128, device="cuda")
        ev = torch.empty((N, nmt), dtype=torch.float64, device="cuda")
        def timed(f, reps=2):
            f(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): f()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        t1 = timed(lambda: lib.sktb_heev_twostage(ham._plan, M.data_ptr(), N, nmt, band.data_ptr(), None, None))
        t12 = timed(lambda: lib.sktb_heev_twostage(ham._plan, M.data_ptr(), N, nmt, band.data_ptr(), de.data_ptr(), None))
        tall = timed(lambda: lib.sktb_heev(ham._plan, M.data_ptr(), N, nmt, ev.data_ptr(), None, None, None))
        fl = 16.0 / 3 * N ** 3 * nmt
        print(f"N={N} nm={nmt}: stage1 {t1:.2f} ms ({nmt/t1*1e3:.3e}/s, {fl/t1/1e9:.2f} TF conv)  stage2 {t12-t1:.2f} ms ({nmt/(t12-t1)*1e3:.3e}/s)  "
              f"sktb_heev {tall:.2f} ms ({nmt/tall*1e3:.3e}/s, {fl/tall/1e9:.2f} TF conv) [{ham.last_kernel_info()[:40]}]", flush=True)

# adversarial inputs through sktb_heev (two-stage path forced): zero / diagonal / decoupled blocks / exact degeneracies / scales
if os.environ.get("TS_ADV"):
    rng = np.random.default_rng(7)
    for N in [int(x) for x in os.environ["TS_ADV"].split(",")]:
        cases = {}
        cases["zero"] = np.zeros((N, N), complex)
        cases["diag_repeated"] = np.diag(np.repeat(np.array([-1.5, 0.0, 0.0, 2.0]), N // 4 + 1)[:N]).astype(complex)
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)); A = A + A.conj().T
        B = A.copy(); h = N // 2; B[h:, :h] = 0; B[:h, h:] = 0
        cases["two_blocks"] = B
        C = np.zeros((N, N), complex)                                  # bipartite hopping, zero diagonal: spectrum symmetric about 0, zero modes
        for i in range(N - 1): C[i + 1, i] = C[i, i + 1] = 1.0 if i % 2 == 0 else 0.3
        cases["ssh_chain"] = C
        cases["tiny_scale"] = A * 1e-7
        cases["big_scale"] = A * 1e3
        Q, _ = np.linalg.qr(A)
        lam = np.repeat(np.arange(N // 8 + 1, dtype=float), 8)[:N]       # 8-fold exact degeneracies
        cases["degenerate8"] = (Q * lam) @ Q.conj().T
        M = torch.tensor(np.stack(list(cases.values())), device="cuda")
        nmc = M.shape[0]
        ev = torch.empty((N, nmc), dtype=torch.float64, device="cuda")
        _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nmc, ev.data_ptr(), None, None, None))
        torch.cuda.synchronize()
        got = ev.cpu().numpy()
        for i, (name, mat) in enumerate(cases.items()):
            ref = np.linalg.eigvalsh(mat, UPLO="L")
            sc = max(1e-300, np.abs(ref).max())
            print(f"N={N} {name:14s} max err {np.abs(got[:, i] - ref).max():.2e} (rel {np.abs(got[:, i] - ref).max() / sc:.1e})  [{ham.last_kernel_info()[-45:]}]", flush=True)
