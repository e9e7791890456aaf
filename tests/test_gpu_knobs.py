"""The A/B knobs select alternative kernels for the same result: every alternative path must stay
parity-green, so each is run in a subprocess (the library reads the environment once) on random Hermitian batches
against LAPACK."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_ready():
    try:
        import torch
        return torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "pysktb_b200", "libsktb.so"))
    except Exception:
        return False


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _gpu_ready(), reason="needs a CUDA device and the built libsktb.so")]

CHECK = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
_, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
lib = _lib.lib()
worst = 0.0
for N, nm, vec in %s:
    g = torch.Generator(device="cuda").manual_seed(N)
    M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda", generator=g)
    Mh = M.cpu().numpy()
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    U = torch.empty((N, nm, N), dtype=torch.complex128, device="cuda") if vec else None
    _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nm, ev.data_ptr(), U.data_ptr() if vec else None, None, None))
    torch.cuda.synchronize()
    e = ev.cpu().numpy()
    for q in (0, nm // 2, nm - 1):
        L = np.tril(Mh[q]); A = L + np.tril(Mh[q], -1).conj().T; A[np.diag_indices(N)] = A.diagonal().real
        worst = max(worst, np.abs(e[:, q] - np.linalg.eigvalsh(A)).max() / max(1.0, np.abs(A).max()))
        if vec:
            Uq = U[:, q, :].cpu().numpy().T
            worst = max(worst, np.abs(A @ Uq - Uq * e[:, q][None, :]).max() / max(1.0, np.abs(A).max()))
assert ham.ql_failures() == 0
print("WORST", worst)
assert worst < 1e-11, worst
"""

CASES = {
    "SKTB_CTA64_FRONT=0": [(40, 300, False), (57, 300, False)],
    "SKTB_CTA64_SPLIT=0": [(48, 300, False), (64, 300, True), (40, 70, True)],
    "SKTB_GENERIC_NB=0": [(100, 40, False), (200, 20, False), (130, 10, True)],
    "SKTB_GENERIC_NB=2": [(100, 40, False), (330, 10, False)],
    "SKTB_PEEL_MAX=0": [(65, 60, False), (128, 40, False)],
    "SKTB_GENERIC_REMAP_MIN=32": [(100, 20, False), (330, 10, False), (130, 10, True)],
    "SKTB_GENERIC_REMAP_MIN=100000": [(512, 5, False)],
    "SKTB_PEEL_QL_MIN=1": [(65, 60, False), (100, 70, False), (200, 33, False)],
    "SKTB_GENERIC_SPLIT=0": [(80, 20, True), (130, 10, True)],
    "SKTB_VEC_LIST=0": [(32, 200, True), (64, 100, True)],
    "SKTB_VEC32_SPLIT=0": [(32, 200, True), (20, 100, True), (32, 300, False), (9, 50, False)],
    "SKTB_NO_PAIR=1": [(32, 300, False), (26, 300, False)],
    "SKTB_HEEV_GENERIC=1": [(8, 50, False), (32, 50, True), (64, 30, False)],
}


@pytest.mark.parametrize("knob", sorted(CASES))
def test_alternative_kernel_paths(knob):
    env = dict(os.environ)
    k, v = knob.split("=")
    env[k] = v
    r = subprocess.run([sys.executable, "-c", CHECK % (ROOT, repr(CASES[knob]))], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
