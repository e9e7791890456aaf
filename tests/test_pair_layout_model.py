"""The numpy lane-level model of the pair-layout tridiagonalisation (tools/proto_pair_layout.py, the algorithm of
csrc/heev_pair.cuh: lane = 2r+h, rotating register window, look-ahead mat-vec, A -= v p^H + w' v^H) must produce a
tridiagonal matrix with the spectrum of its input - full, zero-padded (N < 32) and already-tridiagonal inputs."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_pair_layout", os.path.join(ROOT, "tools", "proto_pair_layout.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


def _check(A):
    d, e = proto.tridiag_pair(A.copy())
    T = np.diag(d) + np.diag(e[:31], 1) + np.diag(e[:31], -1)
    assert np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(T)).max() < 1e-12 * max(1.0, np.abs(A).max()) * 32


def test_pair_layout_model_random_and_padded():
    rng = np.random.default_rng(7)
    for n in (32, 26, 17, 5):
        M = rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))
        A = M + M.conj().T
        A[n:, :] = 0
        A[:, n:] = 0
        _check(A)


def test_pair_layout_model_structured_inputs():
    d = np.linspace(-2, 2, 32)
    _check(np.diag(d).astype(complex))                                   # diagonal: every reflector is trivial
    T = np.diag(d).astype(complex) + np.diag(np.full(31, 0.3 + 0.1j), -1)
    _check(T + np.tril(T, -1).conj().T)                                  # already tridiagonal, complex sub-diagonal
    _check(np.ones((32, 32), complex))                                   # rank one
