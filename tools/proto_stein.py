"""Scalar model of stein_kernel (kernels_big_vec.cuh): eigenvectors of a real symmetric tridiagonal matrix by inverse
iteration on given eigenvalues (LAPACK dstein's scheme: LU with partial pivoting of T - lambda I (dlagtf), solves with
perturbed tiny pivots (dlagts, job -1), modified Gram-Schmidt inside clusters of close eigenvalues), without the
splitting into unreduced blocks - negligible off-diagonals are simply kept."""
import numpy as np

EPS = 2.220446049250313e-16


def lagtf(d, e, lam):
    n = len(d)
    a = d - lam
    b = e.copy()            # super-diagonal (n-1)
    c = e.copy()            # sub-diagonal -> multipliers
    d2 = np.zeros(max(n - 2, 0))
    pin = np.zeros(n, dtype=np.int8)
    tl = max(EPS, 0.0)
    scale1 = abs(a[0]) + (abs(b[0]) if n > 1 else 0.0)
    for k in range(n - 1):
        a[k + 1] = a[k + 1]
        scale2 = abs(c[k]) + abs(a[k + 1]) + (abs(b[k + 1]) if k < n - 2 else 0.0)
        piv1 = 0.0 if a[k] == 0 else abs(a[k]) / scale1
        if c[k] == 0.0:
            pin[k] = 0
            piv2 = 0.0
            scale1 = scale2
            if k < n - 2:
                d2[k] = 0.0
        else:
            piv2 = abs(c[k]) / scale2
            if piv2 <= piv1:
                pin[k] = 0
                scale1 = scale2
                c[k] = c[k] / a[k]
                a[k + 1] -= c[k] * b[k]
                if k < n - 2:
                    d2[k] = 0.0
            else:
                pin[k] = 1
                mult = a[k] / c[k]
                a[k] = c[k]
                temp = a[k + 1]
                a[k + 1] = b[k] - mult * temp
                if k < n - 2:
                    d2[k] = b[k + 1]
                    b[k + 1] = -mult * d2[k]
                b[k] = temp
                c[k] = mult
    return a, b, c, d2, pin


def lagts(a, b, c, d2, pin, y, tol):
    n = len(a)
    y = y.copy()
    for k in range(1, n):
        if pin[k - 1] == 0:
            y[k] -= c[k - 1] * y[k - 1]
        else:
            t = y[k - 1]
            y[k - 1] = y[k]
            y[k] = t - c[k - 1] * y[k]
    for k in range(n - 1, -1, -1):
        t = y[k]
        if k <= n - 2:
            t -= b[k] * y[k + 1]
        if k <= n - 3:
            t -= d2[k] * y[k + 2]
        ak = a[k]
        if abs(ak) < tol:
            ak = tol if ak >= 0 else -tol
        t = t / ak
        if abs(t) > 1e150:          # exponentially localised vectors (edge states) span hundreds of decades: rescale the column
            y *= 1e-150
            t *= 1e-150
        y[k] = t
    return y


def stein(d, e, w, seed=0, tight_rel=1e-10):
    """Two-level scheme of the kernel: TIGHT clusters (gaps <= tight_rel ||T||) are computed one vector after the other
    with modified Gram-Schmidt inside the iteration (dstein); everything else independently (one thread each), followed by
    ONE first-order symmetric (Loewdin) correction x_n -= 1/2 sum_m (x_m.x_n) x_m over the LOOSE cluster (gaps <= 1e-3
    ||T||): independent inverse iterations are orthogonal to ~eps ||T|| / gap <= 2e-9 there, the correction squares that."""
    n = len(d)
    d = np.asarray(d, float); e = np.asarray(e, float)[: n - 1]
    ee = np.concatenate([e, [0.0]])
    onenrm = max(abs(d[i]) + (abs(ee[i - 1]) if i > 0 else 0.0) + abs(ee[i]) for i in range(n))
    ortol, tight = 1e-3 * onenrm, tight_rel * onenrm
    tol = EPS * onenrm if onenrm > 0 else EPS
    Z = np.zeros((n, n))
    rng = np.random.default_rng(seed)
    head = 0
    xjm = 0.0
    for j in range(n):
        xj = w[j]
        if j > 0 and w[j] - w[j - 1] > tight:
            head = j
        if j > head:
            pertol = 10.0 * EPS * max(abs(xj), onenrm)      # ABSOLUTE floor: bisection to relative precision can return a pair +-1e-38
            if xj - xjm < pertol:
                xj = xjm + pertol
        xjm = xj
        a, b, c, d2, pin = lagtf(d, e, xj)
        x = rng.uniform(-1, 1, n)
        nrmchk = 0
        for it in range(8):
            scl = n * max(onenrm, 1e-300) * max(EPS, abs(a[n - 1])) / max(np.abs(x).sum(), 1e-300)
            x = lagts(a, b, c, d2, pin, x * scl, tol)
            for i in range(head, j):
                x -= (x @ Z[:, i]) * Z[:, i]
            if np.abs(x).max() >= np.sqrt(0.1 / n):
                nrmchk += 1
                if nrmchk >= 3:
                    break
        mx = np.abs(x).max()
        Z[:, j] = (x / mx) / np.linalg.norm(x / mx)
    ptol = 2e-4 * onenrm          # PAIRWISE window: the overlap of two independent vectors is ~eps ||T|| / |gap|, below 1e-12 beyond it
    for sweep in range(2):        # two passes: E -> ~(c E)^2 -> ~(c E)^4 for a window of c vectors
        Z2 = Z.copy()
        for j in range(n):
            lo = j
            while lo > 0 and w[j] - w[lo - 1] <= ptol:
                lo -= 1
            hi = j
            while hi < n - 1 and w[hi + 1] - w[j] <= ptol:
                hi += 1
            for m in range(lo, hi + 1):
                if m != j:
                    Z2[:, j] -= 0.5 * (Z[:, m] @ Z[:, j]) * Z[:, m]
        Z = Z2
    return Z2


if __name__ == "__main__":
    r = np.random.default_rng(1)
    worst = (0, 0)
    for trial in range(72):
        n = int(r.integers(2, 90)) if trial < 60 else 300
        kind = trial % 6
        d = r.standard_normal(n); e = r.standard_normal(n - 1)
        if kind == 1: d[:] = 0.0
        if kind == 2: e[n // 2 - 1] = 0.0; d[n // 2:] = d[: n - n // 2][: len(d[n // 2:])] if n % 2 == 0 else d[n // 2:]; 
        if kind == 2 and n % 2 == 0: e[n // 2:] = e[: n // 2 - 1]           # two identical blocks: exact double degeneracy
        if kind == 3: e[::4] = 1e-18
        if kind == 4: d[:] = 1.0; e[:] = 1e-7
        if kind == 5: d = np.round(d); e = np.round(e)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w = np.linalg.eigvalsh(T)
        Z = stein(d, e, w)
        res = np.abs(T @ Z - Z * w).max() / max(1.0, np.abs(T).max())
        orth = np.abs(Z.T @ Z - np.eye(n)).max()
        worst = (max(worst[0], res), max(worst[1], orth))
    print("worst residual %.2e orthogonality %.2e" % worst)
