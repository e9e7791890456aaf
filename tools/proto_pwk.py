"""Scalar model of the square-root-free (Pal-Walker-Kahan) implicit QL used by tql_lane_pwk (heev_small.cuh): the
eigenvalues-only recurrence of LAPACK dsterf's QL branch (the routine numpy's eigvalsh ends in), as the flat state
machine of the kernel: one trip = one sweep over the unreduced block [l..m], deflation tests folded into the sweep."""
import numpy as np

EPS = 2.220446049250313e-16


def pwk_eigenvalues(d, e, max_iter=80):
    d = np.array(d, dtype=float)
    n = len(d)
    e2 = np.zeros(n)
    e2[: n - 1] = np.asarray(e, dtype=float)[: n - 1] ** 2
    anorm = max(np.abs(d).max(), np.sqrt(e2.max())) if n else 0.0
    eps2 = EPS * EPS
    floor2 = eps2 * anorm * anorm * 1e-6
    small = lambda m: e2[m] <= eps2 * abs(d[m] * d[m + 1]) + floor2

    def scan(l):
        m = l
        while m < n - 1 and not small(m):
            m += 1
        return m
    fail, l, it, sweeps = 0, 0, 0, 0
    m = scan(0)
    while True:
        if m == l or it > max_iter:
            if m != l:
                fail = 1
            l += 1
            it = 0
            if l >= n:
                break
            m = scan(l)
            continue
        it += 1
        sweeps += 1
        p = d[l]
        rte = np.sqrt(e2[l])
        sg = (d[l + 1] - p) / (2.0 * rte)
        r = np.sqrt(sg * sg + 1.0)
        sigma = p - rte / (sg + (r if sg >= 0 else -r))
        c, s = 1.0, 0.0
        gamma = d[m] - sigma
        pp = gamma * gamma
        mb, dnext = m, 0.0
        for i in range(m - 1, l - 1, -1):
            bb = e2[i]
            r = pp + bb
            e2new = s * r
            oldc = c
            c = pp / r
            s = bb / r
            oldgam = gamma
            alpha = d[i]
            gamma = c * (alpha - sigma) - s * oldgam
            dnew = oldgam + (alpha - gamma)
            d[i + 1] = dnew
            pp = gamma * gamma / c if c > 1e-290 else oldc * bb
            adn = abs(dnew)
            if i != m - 1:
                e2[i + 1] = e2new
                if e2new <= eps2 * adn * dnext + floor2:
                    mb = i + 1
            dnext = adn
        e2l = s * pp
        dl = sigma + gamma
        e2[l], d[l] = e2l, dl
        e2[m] = 0.0
        if e2l <= eps2 * abs(dl) * dnext + floor2:
            l += 1
            it = 0
        m = mb
    return np.sort(d), fail, sweeps


if __name__ == "__main__":
    r = np.random.default_rng(0)
    worst = 0
    for trial in range(400):
        n = int(r.integers(1, 65))
        kind = trial % 6
        d = r.standard_normal(n) * (10 if kind == 1 else 1)
        e = r.standard_normal(max(n - 1, 0))
        if kind == 2: d[:] = 0.0
        if kind == 3: e[::3] = 0.0
        if kind == 4: d[:] = 1.0; e[:] = 1e-9
        if kind == 5: d = np.round(d); e = np.round(e)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ref = np.linalg.eigvalsh(T)
        ev, fail, sw = pwk_eigenvalues(d, e)
        assert not fail
        worst = max(worst, np.abs(ev - ref).max() / max(1.0, np.abs(ref).max()))
    print("worst", worst)
