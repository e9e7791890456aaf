"""Lane-level numpy prototype of the 'pair layout' Householder tridiagonalisation (tridiag_pair32 kernel).

32 lanes per matrix: lane = 2 r + h, r in [0,16) row pair (slot A: row r, slot B: row 31-r), h = column parity.
Register window B[slot][m] holds column base + 2 m + h.  Look-ahead: the mat-vec of step j+1 is accumulated
while step j's rank-2 update streams over the window.  This file only validates the index/conjugation
conventions used by csrc/heev_pair.cuh; it is not part of the product path.
"""
import numpy as np

N = 32


def householder(alpha, xn):
    if xn == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0 + 0.0j, 0.0 + 0.0j
    nrm = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xn)
    beta = -nrm if alpha.real >= 0 else nrm
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    scale = 1.0 / (alpha - beta)
    return beta, tau, scale


def tridiag_pair(Afull, S=4):
    lanes = 32
    r = np.arange(lanes) // 2
    h = np.arange(lanes) % 2
    rows = np.stack([r, 31 - r], axis=1)          # [lane][slot]
    base, W = 0, 32
    B = np.zeros((lanes, 2, 16), complex)
    for l in range(lanes):
        for s in range(2):
            for m in range(16):
                B[l, s, m] = Afull[rows[l, s], 2 * m + h[l]]
    V = np.zeros((2, 64), complex); P = np.zeros(64, complex); X = np.zeros(64, complex)
    ALPHA = [0j]; D = np.zeros(32); E = np.zeros(32)

    def extract(jn, m_e, h_e):
        # owner lanes (parity h_e) publish the new pivot column jn
        for l in range(lanes):
            if h[l] != h_e:
                continue
            for s in range(2):
                i = rows[l, s]; x = B[l, s, m_e]
                X[i] = x if i > jn + 1 else 0.0
                if i == jn + 1: ALPHA[0] = x
                if i == jn: D[jn] = x.real

    def epilogue(jn, z, m_a, h_a, vbuf):
        # all lanes: reflector of step jn, p, w' ; publishes V[vbuf], P
        xo = X[rows]                                     # [lane][slot]
        part = (np.abs(xo) ** 2).sum(axis=1)
        xn = part.reshape(16, 2)[:, 0].sum()             # reduce over r within parity
        beta, tau, scale = householder(ALPHA[0], xn)
        E[jn] = beta
        v = np.where(rows == jn + 1, 1.0, xo * scale)
        v = np.where(rows <= jn, 0.0, v)
        ypart = scale * z + np.where((h == h_a)[:, None], B[:, :, m_a], 0.0)
        y = ypart + ypart.reshape(16, 2, 2)[:, ::-1, :].reshape(32, 2)     # xor-1 combine
        y = np.where(rows <= jn, 0.0, y)
        p = tau * y
        g = (np.conj(y) * v).real.sum(axis=1)
        G = g.reshape(16, 2)[:, 0].sum()
        coef = -abs(tau) ** 2 * G
        wp = p + coef * v
        for l in range(lanes):
            for s in range(2):
                V[vbuf, rows[l, s]] = v[l, s]; P[rows[l, s]] = p[l, s]
        return v, wp

    # prologue: pivot column 0, plain mat-vec
    extract(0, 0, 0)
    z = np.zeros((lanes, 2), complex)
    for m in range(16):
        c = base + 2 * m + h
        z += B[:, :, m] * X[c][:, None]
    v, wp = epilogue(0, z, 0, 1, 0)      # column 1: parity 1, register 0
    nslot = 2
    j = 0
    while j <= 30:
        if j % S == 0:
            W = 32 - j; assert base == j
        half = W // 2
        even = (j == base)
        vb = j & 1
        m_e, h_e = (0, 1) if even else (1, 0)
        m_a, h_a = (1, 0) if even else (1, 1)

        def upd(m, part2=True):
            c = base + 2 * m + h
            B[:, :, m] -= v * np.conj(P[c])[:, None]
            B[:, :, m] -= wp * np.conj(V[vb, c])[:, None]
        if m_e < half: upd(m_e)
        extract(j + 1, m_e, h_e) if m_e < half else None
        z = np.zeros((lanes, 2), complex)
        for m in range(m_e + 1, half):
            upd(m)
            c = base + 2 * m + h
            z += B[:, :, m] * X[c][:, None]
        if m_a >= half:     # window too narrow (only at the very end): column j+2 does not exist
            Bsave = None
        v, wp = epilogue(j + 1, z, min(m_a, half - 1) if m_a < half else 0, h_a if m_a < half else 2, (j + 1) & 1)
        if not even:
            B[:, :, :-1] = B[:, :, 1:]; B[:, :, -1] = 0.0; base += 2
        j += 1
    return D, E


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for trial in range(5):
        M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        A = M + M.conj().T
        if trial == 4:
            A[20:, :] = 0; A[:, 20:] = 0          # zero padding (N < NC)
        d, e = tridiag_pair(A.copy())
        T = np.diag(d) + np.diag(e[:31], 1) + np.diag(e[:31], -1)
        ref = np.linalg.eigvalsh(A)
        got = np.linalg.eigvalsh(T)
        print(trial, np.abs(ref - got).max())
