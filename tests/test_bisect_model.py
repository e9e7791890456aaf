"""The numpy model of the Sturm-count bisection (tools/proto_bisect.py = the eigenvalue stage of the generic path above
N = 48 and of bisect_merge_kernel) against LAPACK on random, clustered, split (zero off-diagonal) and scaled tridiagonals."""
import importlib.util
import os

import numpy as np
import pytest
from scipy.linalg import eigvalsh_tridiagonal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_bisect", os.path.join(ROOT, "tools", "proto_bisect.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


def _cases():
    r = np.random.default_rng(5)
    yield "random", r.standard_normal(40), r.standard_normal(39)
    yield "clustered", np.full(30, 1.5) + 1e-13 * r.standard_normal(30), 1e-9 * r.standard_normal(29)
    e = r.standard_normal(31)
    e[[4, 5, 17]] = 0.0
    yield "split", r.standard_normal(32), e
    yield "scaled_small", 1e-7 * r.standard_normal(20), 1e-7 * r.standard_normal(19)
    yield "scaled_large", 1e3 * r.standard_normal(20), 1e3 * r.standard_normal(19)
    yield "wilkinson", np.abs(np.arange(-10, 11)).astype(float), np.ones(20)
    yield "single", np.array([0.7]), np.array([])


@pytest.mark.parametrize("name,d,e", list(_cases()), ids=[c[0] for c in _cases()])
def test_bisection_model_vs_lapack(name, d, e):
    got = proto.bisect_eigenvalues(d, e)
    ref = eigvalsh_tridiagonal(d, e) if len(d) > 1 else d.copy()
    scale = max(1e-300, np.abs(ref).max())
    assert np.all(np.diff(got) >= 0.0)
    assert np.abs(got - ref).max() <= 64 * 2.3e-16 * max(scale, np.abs(d).max())
