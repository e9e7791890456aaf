"""numpy model of the two-stage tridiagonalisation (csrc/heev_twostage.cuh): dense Hermitian (lower triangle read) ->
band of width B by blocked Householder QR panels + two-sided block update (stage 1), band -> tridiagonal by bulge chasing
(stage 2).  Same index conventions, reflector conventions (zlarfg: H^H x = beta e1, H = I - tau v v^H) and band layout
AB[j][d] = A[j+d][j], d = 0..2B-1, as the kernels.  Test infrastructure / design model only."""
import numpy as np

B = 16


def zlarfg(alpha, x):
    xn2 = float(np.vdot(x, x).real)
    if xn2 == 0.0 and alpha.imag == 0.0:
        return alpha.real, np.zeros_like(x), 0.0 + 0.0j
    beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xn2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    return beta, x / (alpha - beta), tau


def stage1(A, b=B):
    """A: full Hermitian copy (only the lower triangle is read / written).  Returns AB [N][2b]."""
    A = np.tril(A).astype(complex)                        # make sure nothing above the diagonal is used
    idx = np.arange(A.shape[0])
    A[idx, idx] = A[idx, idx].real
    N = A.shape[0]
    AB = np.zeros((N, 2 * b), complex)
    c0 = 0
    while True:
        m = N - c0 - b
        if m <= 0:
            break
        P = A[c0 + b:, c0:c0 + b].copy()                 # m x b panel
        V = np.zeros((m, b), complex); R = np.zeros((b, b), complex); T = np.zeros((b, b), complex)
        taus = np.zeros(b, complex)
        for c in range(b):
            if c >= m:                                     # no such row: identity
                continue
            raw = P[c + 1:, c].conj() @ P[c + 1:, :]      # raw_w = sum_{r>c} conj(P[r][c]) P[r][w]   (all w)
            beta, v, tau = zlarfg(P[c, c], P[c + 1:, c])
            scale = 0.0 if tau == 0 else 1.0 / (P[c, c] - beta)
            for w in range(c + 1, b):
                s = P[c, w] + np.conj(scale) * raw[w]
                P[c, w] -= np.conj(tau) * s
                P[c + 1:, w] -= np.conj(tau) * s * scale * P[c + 1:, c]
            R[:c, c] = P[:c, c]; R[c, c] = beta
            z = np.conj(V[c, :c]) + scale * np.conj(raw[:c]) if c else np.zeros(0, complex)
            P[c + 1:, c] *= scale
            V[:, c] = 0; V[c, c] = 1; V[c + 1:, c] = P[c + 1:, c]
            P[:, c] = V[:, c]                             # the panel buffer now holds V in column c (raw dots of later columns use it)
            T[:c, c] = -tau * (T[:c, :c] @ z); T[c, c] = tau
            taus[c] = tau
        for c in range(min(m, b), b):                     # columns without a reflector keep their (updated) entries as R
            R[:m, c] = P[:m, c] if c >= m else R[:m, c]
        # band output of the b columns: diagonal block (lower) + R
        for c in range(b):
            j = c0 + c
            for d in range(0, b - c):
                AB[j, d] = A[j + d, j]
            for r in range(0, min(c + 1, m)):
                AB[j, b + r - c] = R[r, c]
        # two-sided update of the trailing matrix A22 (lower triangle)
        A22 = A[c0 + b:, c0 + b:]
        full = np.tril(A22) + np.tril(A22, -1).conj().T
        Vt = V @ T
        X = full @ Vt
        G = V.conj().T @ X
        S = T.conj().T @ G
        W = X - 0.5 * V @ S
        upd = V @ W.conj().T + W @ V.conj().T
        A[c0 + b:, c0 + b:] = np.tril(full - upd)
        c0 += b
    for j in range(c0, N):                                # remaining columns are inside the band already
        for d in range(0, N - j):
            AB[j, d] = A[j + d, j]
    return AB


def band_to_dense(AB):
    N, w = AB.shape
    A = np.zeros((N, N), complex)
    for j in range(N):
        for d in range(min(w, N - j)):
            A[j + d, j] = AB[j, d]
    return np.tril(A) + np.tril(A, -1).conj().T


def stage2(AB, b=B):
    """Bulge chasing on the band (lower, fill kept in d = b+1..2b-1).  Returns d, e (real)."""
    AB = AB.copy()
    N = AB.shape[0]

    def get(i, j):
        return AB[j, i - j]

    def put(i, j, x):
        AB[j, i - j] = x

    def two_sided(r1, L, v, tau):
        D = np.zeros((L, L), complex)
        for i in range(L):
            for j in range(i + 1):
                D[i, j] = get(r1 + i, r1 + j)
        D = np.tril(D) + np.tril(D, -1).conj().T
        D[np.arange(L), np.arange(L)] = D[np.arange(L), np.arange(L)].real
        x = tau * (D @ v)
        alpha = -0.5 * tau * np.vdot(x, v)
        w = x + alpha * v
        D = D - np.outer(v, w.conj()) - np.outer(w, v.conj())
        for i in range(L):
            for j in range(i + 1):
                put(r1 + i, r1 + j, D[i, j])

    for s in range(N - 1):
        L = min(b, N - 1 - s)
        x = np.array([get(s + 1 + r, s) for r in range(L)])
        beta, vt, tau = zlarfg(x[0], x[1:])
        v = np.concatenate([[1.0 + 0j], vt])
        put(s + 1, s, beta)
        for r in range(1, L):
            put(s + 1 + r, s, 0.0)
        r0 = s + 1
        two_sided(r0, L, v, tau)
        while True:
            r1 = r0 + L
            if r1 >= N:
                break
            L2 = min(b, N - r1)
            Bk = np.array([[get(r1 + r, r0 + c) if (r1 + r) - (r0 + c) < 2 * b else 0.0 for c in range(L)] for r in range(L2)], complex)
            Bk = Bk - tau * np.outer(Bk @ v, v.conj())
            beta2, vt2, tau2 = zlarfg(Bk[0, 0], Bk[1:, 0])
            v2 = np.concatenate([[1.0 + 0j], vt2])
            u = v2.conj() @ Bk
            Bk = Bk - np.conj(tau2) * np.outer(v2, u)
            Bk[0, 0] = beta2; Bk[1:, 0] = 0
            for r in range(L2):
                for c in range(L):
                    if (r1 + r) - (r0 + c) < 2 * b:
                        put(r1 + r, r0 + c, Bk[r, c])
            two_sided(r1, L2, v2, tau2)
            r0, L, v, tau = r1, L2, v2, tau2
    return AB[:, 0].real.copy(), AB[:-1, 1].copy()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (40, 64, 97, 130):
        M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        A = M + M.conj().T
        ref = np.linalg.eigvalsh(A)
        AB = stage1(A)
        e1 = np.linalg.eigvalsh(band_to_dense(AB))
        d, e = stage2(AB)
        assert np.abs(e.imag).max() < 1e-13, np.abs(e.imag).max()
        Tm = np.diag(d) + np.diag(e.real, 1) + np.diag(e.real, -1)
        e2 = np.linalg.eigvalsh(Tm)
        print(N, "stage1", np.abs(e1 - ref).max(), "stage2", np.abs(e2 - ref).max())
