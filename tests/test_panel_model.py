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
