"""The numpy model of the panel formulation (tools/proto_panel.py; algorithm of the generic kernel's `NB > 0` branch and
of the peeled run in csrc/kernels_generic.cuh / csrc/sktb.cu): deferred rank-2NB updates with on-the-fly corrections keep
the spectrum, for panel widths 2/4/8, sizes that are not multiples of the width, and when the run stops after P steps and
hands the trailing block over (d, e of the peeled steps + the block reproduce the spectrum)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_panel", os.path.join(ROOT, "tools", "proto_panel.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


def _herm(n, seed):
    r = np.random.default_rng(seed)
    M = r.standard_normal((n, n)) + 1j * r.standard_normal((n, n))
    return M + M.conj().T


@pytest.mark.parametrize("N,NB", [(9, 2), (33, 4), (40, 8), (67, 8), (96, 4)])
def test_panel_tridiagonalisation_keeps_the_spectrum(N, NB):
    A = _herm(N, N)
    d, e = proto.tridiag_panel(A, NB)
    T = np.diag(d) + np.diag(e[:N - 1], 1) + np.diag(e[:N - 1], -1)
    assert np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(T)).max() < 1e-11


@pytest.mark.parametrize("N", [65, 72, 100, 161])
def test_peeled_run_hands_over_a_consistent_block(N):
    """P = the multiple of the panel width that leaves a trailing block of 57..64 (sktb.cu, launch_heev_peel)."""
    nb = 8 if N >= 160 else 4
    P = (N - 64 + nb - 1) // nb * nb
    A = _herm(N, 1000 + N)
    d, e, blk = proto.tridiag_panel(A, nb, peel=P)
    assert blk.shape == (N - P, N - P) and 57 <= N - P <= 64
    assert np.abs(blk - blk.conj().T).max() < 1e-12
    # head tridiagonal + coupling e[P-1] to the first row of the block + the block itself
    Tfull = np.zeros((N, N), complex)
    Tfull[:P, :P] = np.diag(d) + np.diag(e[:P - 1], 1) + np.diag(e[:P - 1], -1)
    Tfull[P:, P:] = blk
    Tfull[P, P - 1] = Tfull[P - 1, P] = e[P - 1]
    assert np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(Tfull)).max() < 1e-11


@pytest.mark.parametrize("N,NB", [(6, 2), (32, 4), (45, 8)])
def test_reflector_store_and_back_transformation(N, NB):
    """Eigenvector path: tridiagonalise keeping (v_j, tau_j), eigenvectors Z of the real tridiagonal matrix,
    U = H_0 ... H_{N-2} Z diagonalises the input (conventions of the vecbt32 / vecbt64 kernels)."""
    A = _herm(N, 7 * N)
    refl = {}
    d, e = proto.tridiag_panel(A, NB, reflectors=refl)
    T = np.diag(d) + np.diag(e[:N - 1], 1) + np.diag(e[:N - 1], -1)
    lam, Z = np.linalg.eigh(T)
    U = proto.back_transform(refl, Z)
    assert np.abs(U.conj().T @ U - np.eye(N)).max() < 1e-12
    assert np.abs(A @ U - U * lam[None, :]).max() < 1e-11 * max(1.0, np.abs(A).max())
    # v_j has zeros up to row j and a unit entry at j+1, as the kernels store it explicitly
    for j in range(N - 1):
        assert np.all(refl["V"][:j + 1, j] == 0) and refl["V"][j + 1, j] == 1


@pytest.mark.parametrize("N", [33, 40, 57, 63])
def test_front_padding_makes_the_leading_steps_trivial(N):
    """tridiag_cta64 pads 32 < N < 64 at the FRONT of its 64-wide frame and skips the phases before
    J0 = 8 * ((64 - N) // 8): the steps over the zero padding are identity reflectors with d = e = 0, so the tridiagonal
    matrix of the frame is 0_(64-N) (+) T(A) and starting at J0 changes nothing."""
    off = 64 - N
    A = _herm(N, 50 + N)
    frame = np.zeros((64, 64), complex)
    frame[off:, off:] = A
    refl = {}
    d, e = proto.tridiag_panel(frame, 4, reflectors=refl)
    assert np.all(d[:off] == 0.0) and np.all(e[:off] == 0.0) and np.all(refl["tau"][:off] == 0.0)
    T = np.diag(d[off:]) + np.diag(e[off:63], 1) + np.diag(e[off:63], -1)
    assert np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(T)).max() < 1e-11
