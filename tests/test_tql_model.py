"""The Python model of the per-lane QL state machine (tools/proto_tql.py = tql_lane in csrc/heev_small.cuh, including
the in-sweep deflation bookkeeping) against LAPACK: random, clustered, split, graded and Wilkinson tridiagonals; the sweep
count stays near the textbook ~1.7 per eigenvalue and nothing hits the iteration limit."""
import importlib.util
import os

import numpy as np
import pytest
from scipy.linalg import eigvalsh_tridiagonal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_tql", os.path.join(ROOT, "tools", "proto_tql.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


def _cases():
    r = np.random.default_rng(11)
    for n in (2, 5, 32, 64):
        yield f"random{n}", r.standard_normal(n), r.standard_normal(n - 1)
    yield "clustered", np.full(32, -0.4) + 1e-12 * r.standard_normal(32), 1e-8 * r.standard_normal(31)
    e = r.standard_normal(31)
    e[[0, 9, 10, 30]] = 0.0
    yield "split", r.standard_normal(32), e
    yield "graded", 10.0 ** -np.arange(16.0), 10.0 ** -(np.arange(15.0) + 0.5)
    yield "wilkinson", np.abs(np.arange(-10, 11)).astype(float), np.ones(20)
    yield "diagonal", r.standard_normal(12), np.zeros(11)


@pytest.mark.parametrize("name,d,e", list(_cases()), ids=[c[0] for c in _cases()])
def test_tql_state_machine_vs_lapack(name, d, e):
    ev, sweeps, fail = proto.tql_lane(d, e)
    ref = eigvalsh_tridiagonal(d, e)
    assert fail == 0
    assert sweeps <= 3 * len(d)
    scale = max(np.abs(d).max(), np.abs(e).max() if len(e) else 0.0, 1e-300)
    assert np.abs(np.sort(ev) - ref).max() <= 40 * len(d) * 2.3e-16 * scale


def test_tiny_rotation_guard_deflates():
    """The CUDA fast path treats h2 < 1e-290 as an exact zero (deflation); the model with the same guard agrees with
    LAPACK on eV-scale input."""
    r = np.random.default_rng(3)
    d, e = r.standard_normal(32), r.standard_normal(31)
    ev, _, fail = proto.tql_lane(d, e, tiny=1e-290)
    assert fail == 0 and np.abs(np.sort(ev) - eigvalsh_tridiagonal(d, e)).max() < 1e-13


@pytest.mark.parametrize("n", [2, 7, 32, 64])
def test_record_and_replay_gives_the_eigenvectors(n):
    """Record-and-replay QL (tqlrec + vecapply kernels): replaying the recorded (c, s) lists on Z = I diagonalises T."""
    r = np.random.default_rng(100 + n)
    d, e = r.standard_normal(n), r.standard_normal(n - 1)
    e[n // 2 - 1] = 0.0 if n > 8 else e[n // 2 - 1]                    # one exact split for the larger sizes
    rec = []
    ev, _, fail = proto.tql_lane(d, e, record=rec)
    Z = proto.replay_rotations(n, rec)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert fail == 0
    assert np.abs(Z.T @ Z - np.eye(n)).max() < 1e-13
    assert np.abs(T @ Z - Z * ev[None, :]).max() < 1e-12 * max(1.0, np.abs(T).max()) * n
