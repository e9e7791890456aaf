"""numpy model of the panel formulation of the generic tridiagonalisation (csrc/kernels_generic.cuh, `a.NB > 0`).

Inside a panel of NB steps the stored matrix stays M0; the reflector pairs (v_s, w_s) accumulate and
    M_t = M0 - sum_{s<t} (v_s w_s^H + w_s v_s^H)
is never formed: the pivot column is corrected on the fly, y = M_t v = M0 v - sum_s (v_s (w_s^H v) + w_s (v_s^H v)) costs
one READ-ONLY pass over M0, and the rank-2NB update is applied once per panel.  `peel` stops after that many steps and
returns the updated trailing block as the peeled run hands it to the register-resident kernels.
This file only validates the algebra and index conventions; it is not part of the product path.
"""
import numpy as np


def householder(alpha, xnorm2):
    """zlarfg: beta real, H = I - tau v v^H with v = (1, scale * x)."""
    if xnorm2 == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0j, 0.0j
    nrm = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2)
    beta = -nrm if alpha.real >= 0 else nrm
    return beta, complex((beta - alpha.real) / beta, -alpha.imag / beta), 1.0 / (alpha - beta)


def tridiag_panel(A, NB=8, peel=0, reflectors=None):
    """A Hermitian (N, N).  Returns (d, e) - and the trailing block after `peel` steps when peel > 0.
    reflectors: optional dict that receives "V" (N, N-1; column j = v_j with the unit entry at j+1) and "tau" (N-1,), the
    store the eigenvector kernels keep (csrc/heev_pair.cuh / heev_cta64.cuh, VEC = true)."""
    M0 = np.array(A, dtype=complex)
    N = M0.shape[0]
    d, e = np.zeros(N), np.zeros(N)
    if reflectors is not None:
        reflectors["V"] = np.zeros((N, max(N - 1, 0)), complex)
        reflectors["tau"] = np.zeros(max(N - 1, 0), complex)
    stop = peel if peel > 0 else N
    j0 = 0
    V = W = np.zeros((N, 0), complex)
    while j0 + 1 < N and j0 < stop:
        tmax = min(NB, N - 1 - j0)
        V = np.zeros((N, tmax), complex)
        W = np.zeros((N, tmax), complex)
        for t in range(tmax):
            j = j0 + t
            col = M0[:, j] - V[:, :t] @ W[j, :t].conj() - W[:, :t] @ V[j, :t].conj()      # column j of M_t
            d[j] = col[j].real
            x = col[j + 1:].copy()
            beta, tau, scale = householder(x[0], float(np.sum(np.abs(x[1:]) ** 2)))
            e[j] = beta
            v = np.zeros(N, complex)
            v[j + 1] = 1.0
            v[j + 2:] = x[1:] * scale
            y = M0 @ v                                                       # the read-only pass (rows/cols > j matter)
            y -= V[:, :t] @ (W[:, :t].conj().T @ v) + W[:, :t] @ (V[:, :t].conj().T @ v)
            y[:j + 1] = 0.0
            p = tau * y
            w = p - 0.5 * tau * np.vdot(p, v) * v
            V[:, t], W[:, t] = v, w
            if reflectors is not None:
                reflectors["V"][:, j], reflectors["tau"][j] = v, tau
        jn = j0 + tmax
        if tmax == NB and jn + 1 < N:
            M0 -= V @ W.conj().T + W @ V.conj().T                            # deferred rank-2NB update
            V = W = np.zeros((N, 0), complex)
        j0 = jn
    if peel > 0:
        return d[:peel], e[:peel], M0[peel:, peel:]
    last = M0[N - 1, N - 1] - V[N - 1] @ W[N - 1].conj() - W[N - 1] @ V[N - 1].conj()
    d[N - 1] = last.real
    return d, e


def back_transform(reflectors, Z):
    """U = H_0 H_1 ... H_{N-2} Z with H_j = I - tau_j v_j v_j^H (the vecbt kernels: per eigenvector and reflector one dot
    product t = tau_j (v_j^H u) and one update u -= t v_j, reflectors applied from j = N-2 down to 0)."""
    U = np.array(Z, dtype=complex)
    V, tau = reflectors["V"], reflectors["tau"]
    for j in range(V.shape[1] - 1, -1, -1):
        t = tau[j] * (V[:, j].conj() @ U)
        U -= np.outer(V[:, j], t)
    return U


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, NB in ((40, 8), (37, 4), (96, 8)):
        M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        A = M + M.conj().T
        d, e = tridiag_panel(A, NB)
        T = np.diag(d) + np.diag(e[:N - 1], 1) + np.diag(e[:N - 1], -1)
        print(N, NB, np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(T)).max())
