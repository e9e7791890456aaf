"""CPU test of the numpy model of the two-stage tridiagonalisation (tools/proto_twostage.py: same blocking, reflector conventions and
band indexing as csrc/heev_twostage.cuh) against LAPACK."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import proto_twostage as P  # noqa: E402


@pytest.mark.parametrize("N", [17, 32, 33, 40, 47, 64, 97])
def test_band_and_tridiagonal_have_the_spectrum(N):
    rng = np.random.default_rng(N)
    M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ref = np.linalg.eigvalsh(M, UPLO="L")
    AB = P.stage1(M)
    assert np.abs(AB[:, P.B + 1:]).max() == 0.0
    assert np.abs(np.linalg.eigvalsh(P.band_to_dense(AB)) - ref).max() < 1e-12 * N
    d, e = P.stage2(AB)
    assert np.abs(e.imag).max() < 1e-13
    T = np.diag(d) + np.diag(e.real, 1) + np.diag(e.real, -1)
    assert np.abs(np.linalg.eigvalsh(T) - ref).max() < 1e-12 * N


def test_decoupled_and_diagonal_inputs():
    N = 48
    A = np.diag(np.arange(N, dtype=float)).astype(complex)
    d, e = P.stage2(P.stage1(A))
    assert np.allclose(np.sort(d), np.arange(N)) and np.abs(e).max() == 0.0
    rng = np.random.default_rng(1)
    M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)); M = M + M.conj().T
    M[24:, :24] = 0; M[:24, 24:] = 0
    d, e = P.stage2(P.stage1(M))
    T = np.diag(d) + np.diag(e.real, 1) + np.diag(e.real, -1)
    assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(M)).max() < 1e-11
