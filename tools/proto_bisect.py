"""numpy model of the bisection used for the eigenvalues of the tridiagonal matrices on the generic path
(csrc/kernels_generic.cuh: in-kernel bisection and bisect_merge_kernel): thread n brackets the n-th smallest eigenvalue
with the Sturm count of LAPACK's dstebz recurrence  q_r = d_r - x - e_{r-1}^2 / q_{r-1}  (pivmin safeguard), starting
from the Gershgorin interval, so the result is ascending by construction.  Not part of the product path."""
import numpy as np

EPS = 2.220446049250313e-16
SAFMIN = 2.2250738585072014e-308


def sturm_count(d, e2, x, pivmin):
    q = d[0] - x
    if abs(q) < pivmin:
        q = -pivmin
    cnt = int(q < 0.0)
    for r in range(1, len(d)):
        q = d[r] - x - e2[r - 1] / q
        if abs(q) < pivmin:
            q = -pivmin
        cnt += int(q < 0.0)
    return cnt


def bisect_eigenvalues(d, e):
    """d (N,), e (N-1,) -> ascending eigenvalues, one independent bisection per index like one thread per eigenvalue."""
    d = np.asarray(d, float)
    e = np.asarray(e, float)
    N = len(d)
    el = np.concatenate([[0.0], np.abs(e)])
    er = np.concatenate([np.abs(e), [0.0]])
    gl, gu = float(np.min(d - el - er)), float(np.max(d + el + er))
    tn = float(np.max(np.maximum(np.abs(d), np.maximum(el, er))))
    pivmin = max(SAFMIN * max(1.0, tn * tn), 1e-290)
    gl -= 2.0 * EPS * tn * N + 2.0 * pivmin
    gu += 2.0 * EPS * tn * N + 2.0 * pivmin
    e2 = e * e
    out = np.empty(N)
    for n in range(N):
        lo, hi = gl, gu
        for _ in range(128):
            x = 0.5 * (lo + hi)
            if not (hi - lo > 2.0 * EPS * max(abs(lo), abs(hi)) + 2.0 * pivmin) or x <= lo or x >= hi:
                break
            if sturm_count(d, e2, x, pivmin) > n:
                hi = x
            else:
                lo = x
        out[n] = 0.5 * (lo + hi)
    return out
