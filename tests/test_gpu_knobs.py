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
    "SKTB_TWOSTAGE_MIN=65": [(65, 40, False), (100, 30, False), (200, 20, False), (333, 7, False), (130, 10, True)],   # two-stage from N = 65 (vectors: unaffected)
    "SKTB_TWOSTAGE_MIN=100000": [(352, 8, False), (512, 5, False)],                                                    # one-stage path for every N
    "SKTB_TWOSTAGE_BISECT=0": [(352, 8, False), (400, 6, False)],                                                      # quotient-form Sturm count
}


@pytest.mark.parametrize("knob", sorted(CASES))
def test_alternative_kernel_paths(knob):
    env = dict(os.environ)
    k, v = knob.split("=")
    env[k] = v
    r = subprocess.run([sys.executable, "-c", CHECK % (ROOT, repr(CASES[knob]))], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


FUSE_CHECK = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
out = {}
for name, spec in (("C4", W.ce_spdf_2atom()), ("zz20", W.zigzag_ribbon(n_chains=20)), ("zz31", W.zigzag_ribbon(n_chains=31))):
    _, ham = W.instantiate(spec, tb, tb.scaling)
    k = np.random.default_rng(5).random((37, 3))
    out[name] = ham.solve_kpath(list(k), soc=spec["soc"])
    v, U = ham.solve_kpath(list(k[:5]), soc=spec["soc"], eig_vectors=True)
    out[name + "_v"] = v
    out[name + "_w"] = np.abs(U) ** 2
    out[name + "_info"] = np.array(ham.last_kernel_info())
np.savez(sys.argv[1], **out)
"""


def test_fused_hk_equals_materialised_hk(tmp_path):
    """32 < N <= 64: K0 + K1 fused into tridiag_cta64 (default) against the hk_kernel -> workspace -> tridiag_cta64 path
    (SKTB_CTA64_FUSE=0): same bond order and arithmetic, so eigenvalues are bit-identical."""
    import numpy as np
    res = {}
    for fuse in ("1", "0"):
        env = dict(os.environ, SKTB_CTA64_FUSE=fuse)
        f = str(tmp_path / f"fuse{fuse}.npz")
        r = subprocess.run([sys.executable, "-c", FUSE_CHECK % ROOT, f], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[fuse] = np.load(f)
    assert "fused" in str(res["1"]["C4_info"]) and "fused" not in str(res["0"]["C4_info"]), (str(res["1"]["C4_info"]), str(res["0"]["C4_info"]))
    for key in ("C4", "zz20", "zz31", "C4_v", "zz20_v", "zz31_v"):
        assert np.array_equal(res["1"][key], res["0"][key]), key
    for key in ("C4_w", "zz20_w", "zz31_w"):
        np.testing.assert_allclose(res["1"][key], res["0"][key], rtol=0, atol=1e-12)


THREAD_CHECK = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
out = {}
for name, spec in (("C1", W.cubic_sp3()), ("spchain", W.sp_chain()), ("cubic_s", W.cubic_s())):
    _, ham = W.instantiate(spec, tb, tb.scaling)
    k = np.random.default_rng(11).random((1001, 3)) * 2 - 0.5
    out[name] = ham.solve_kpath(list(k), soc=spec["soc"])
    out[name + "_nosoc"] = ham.solve_kpath(list(k[:77]), soc=False)
    out[name + "_mesh"] = ham.solve_kmesh([7, 5, 3], soc=spec["soc"]).cpu().numpy()
    out[name + "_info"] = np.array(ham.last_kernel_info())
np.savez(sys.argv[1], **out)
"""


def test_thread_per_matrix_equals_lane_group_kernels(tmp_path):
    """N <= 8: the one-matrix-per-thread kernel (default) against the lane-group kernels (SKTB_NO_THREAD=1): same
    eigenvalues to 1e-12 (different summation order of the bond phases, sincospi instead of sincos(2 pi x))."""
    import numpy as np
    res = {}
    for knob in ("0", "1"):
        env = dict(os.environ, SKTB_NO_THREAD=knob)
        f = str(tmp_path / f"thread{knob}.npz")
        r = subprocess.run([sys.executable, "-c", THREAD_CHECK % ROOT, f], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[knob] = np.load(f)
    assert "solve_thread" in str(res["0"]["C1_info"]) and "solve_thread" not in str(res["1"]["C1_info"]), (str(res["0"]["C1_info"]), str(res["1"]["C1_info"]))
    for key in res["0"].files:
        if not key.endswith("_info"):
            np.testing.assert_allclose(res["0"][key], res["1"][key], rtol=0, atol=1e-12, err_msg=key)


TAIL8_CHECK = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pysktb_b200 as tb
from pysktb_b200 import workloads as W
out = {}
for name, spec in (("T", W.ce_spdf_1atom()), ("ce13", W.ce_partial(n_orb=13)), ("zz16", W.zigzag_ribbon(n_chains=16))):
    _, ham = W.instantiate(spec, tb, tb.scaling)
    k = np.random.default_rng(2).random((777, 3)) * 2 - 0.5
    out[name] = ham.solve_kpath(list(k), soc=spec["soc"])
    out[name + "_mesh"] = ham.solve_kmesh([9, 7, 5], soc=spec["soc"]).cpu().numpy()
    out[name + "_info"] = np.array(ham.last_kernel_info())
np.savez(sys.argv[1], **out)
"""


def test_tail8_thread_stage_equals_lane_group_tail(tmp_path):
    """24 < N <= 32 eigenvalue path: the one-matrix-per-thread stage for the trailing 8 x 8 block (default) against the
    lane-group tail kernel doing all 15 steps (SKTB_TAIL8=0)."""
    import numpy as np
    res = {}
    for knob in ("1", "0"):
        env = dict(os.environ, SKTB_TAIL8=knob)
        f = str(tmp_path / f"tail8_{knob}.npz")
        r = subprocess.run([sys.executable, "-c", TAIL8_CHECK % ROOT, f], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[knob] = np.load(f)
    assert "tail8_thread" in str(res["1"]["T_info"]) and "tail8_thread" not in str(res["0"]["T_info"]), (str(res["1"]["T_info"]), str(res["0"]["T_info"]))
    for key in res["0"].files:
        if not key.endswith("_info"):
            np.testing.assert_allclose(res["1"][key], res["0"][key], rtol=0, atol=1e-12, err_msg=key)
