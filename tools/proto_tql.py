"""Python model of the per-lane implicit-shift QL (csrc/heev_small.cuh: tql_lane / tql_scan): a flat state machine - every
trip of the single loop is one sweep over the lane's current unreduced block [l..m] or a bookkeeping step - with the
deflation tests of the NEXT sweep folded into the rotation loop (mb = smallest index whose new off-diagonal is
negligible) so that rescans happen only at block ends.  Returns (eigenvalues unsorted, sweeps, failed).  Not part of the
product path; the CUDA code additionally replaces sqrt/division by Newton-refined hardware seeds."""
import numpy as np

EPS = 2.220446049250313e-16


def tql_scan(d, e, l, n):
    m = l
    while m < n - 1:
        if abs(e[m]) <= EPS * (abs(d[m]) + abs(d[m + 1])):
            break
        m += 1
    return m


def tql_lane(d, e_in, tiny=0.0, record=None):
    """record: optional list that receives the rotation list of the record-and-replay eigenvector path
    (csrc/heev_vec32.cuh, tql_lane_rec): per sweep a header (m, count) followed by `count` pairs (c, s) for the rotations
    i = m-1, m-2, ... in the order they are generated; replay_rotations() applies them to Z."""
    d = np.array(d, float)
    n = len(d)
    e = np.zeros(n)
    e[:n - 1] = np.asarray(e_in, float)[:n - 1]
    fail, l, it, sweeps = 0, 0, 0, 0
    m = tql_scan(d, e, 0, n)
    while True:
        if m == l or it > 80:
            if m != l:
                fail = 1
            l += 1
            it = 0
            if l >= n:
                break
            m = tql_scan(d, e, l, n)
            continue
        it += 1
        sweeps += 1
        el, dl = e[l], d[l]
        g = (d[l + 1] - dl) / (2.0 * el)
        r = np.sqrt(g * g + 1.0)
        g = d[m] - dl + el / (g + (r if g >= 0.0 else -r))
        s, c, p = 1.0, 1.0, 0.0
        under = False
        di1, dnext, mb = d[m], 0.0, m
        i = m - 1
        rots = []
        while i >= l:
            ei, di = e[i], d[i]
            f, b = s * ei, c * ei
            h2 = f * f + g * g
            if h2 <= tiny:
                e[i + 1] = 0.0
                d[i + 1] = di1 - p
                e[m] = 0.0
                under = True
                break
            r = np.sqrt(h2)
            s, c = f / r, g / r
            rots.append((c, s))
            g = di1 - p
            rr = (di - g) * s + 2.0 * c * b
            p = s * rr
            dnew = g + p
            d[i + 1] = dnew
            an = abs(dnew)
            if i + 1 < m:
                e[i + 1] = r
                if r <= EPS * (an + dnext):
                    mb = i + 1
            dnext = an
            g = c * rr - b
            di1 = di
            i -= 1
        if record is not None and rots:
            record.append((m, rots))
        if under:
            m = tql_scan(d, e, l, n)
            continue
        dl_new = di1 - p
        d[l], e[l], e[m] = dl_new, g, 0.0
        if abs(g) <= EPS * (abs(dl_new) + dnext):
            l += 1
            it = 0
        m = mb
    return d, sweeps, fail


def replay_rotations(n, record):
    """Z <- product of the recorded plane rotations (the vecapply kernels: one row of Z per lane, a running `up` value, one
    load and one store per rotation).  Column q of the result is the eigenvector of eigenvalue d[q] of tql_lane."""
    Z = np.eye(n)
    for m, rots in record:
        up = Z[:, m].copy()
        i = m - 1
        for c, s in rots:
            g = Z[:, i].copy()
            Z[:, i + 1] = s * g + c * up
            up = c * g - s * up
            i -= 1
        Z[:, i + 1] = up
    return Z
