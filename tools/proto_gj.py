"""numpy model of gf_inverse_kernel (csrc/gf_inverse.cuh): in-place Gauss-Jordan inversion with IMPLICIT row pivoting (no row
swaps: step c picks the unused row p_c with the largest |Re|+|Im| entry in column c).  With P the permutation that puts row p_c
at position c the sweep computes (P A)^-1 = A^-1 P^T in P A's row numbering, so physically R[p_c'][j] = A^-1[c'][p_j] and
    A^-1[i][m] = R[p_i][q_m],   q = inverse permutation (q[p_c] = c);   diagonal: A^-1[i][i] = R[p_i][q_i]."""
import numpy as np


def gj_inverse_implicit(A):
    R = np.array(A, dtype=complex)
    N = len(R)
    used = np.zeros(N, bool)
    p_of_step = np.zeros(N, int)
    step_of_row = np.zeros(N, int)
    for c in range(N):
        mag = np.where(used, -1.0, np.abs(R[:, c].real) + np.abs(R[:, c].imag))
        p = int(np.argmax(mag))                      # first maximum (LAPACK izamax)
        used[p] = True
        p_of_step[c], step_of_row[p] = p, c
        inv = 1.0 / R[p, c]
        prow = R[p, :] * inv
        prow[c] = inv
        fcol = R[:, c].copy()
        fcol[p] = 0.0
        R[:, c] = 0.0
        R -= np.outer(fcol, prow)
        R[p, :] = prow
    return R, p_of_step, step_of_row


def inverse(A):
    R, p, q = gj_inverse_implicit(A)
    return R[np.ix_(p, q)]


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (1, 2, 5, 32, 64, 97):
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        G = inverse(A)
        print(N, np.abs(G - np.linalg.inv(A)).max())


def warp_kernel_model(A):
    """Lane-level model of gf_inverse_warp_kernel (N <= 32, identity padding): lane r owns row r in a ROTATING 32-register window
    - step c reads its column from register 0, every update writes one register down, the finished inverse column enters at
    register 31 -, the pivot row is broadcast UNSCALED and 1/piv is folded into the factors (g_r = f_r / piv, g_p = 1 - 1/piv),
    and lane r (pivot row of step i = my_step[r]) finds G_ii in register my_step[i].  Returns diag(A^-1)."""
    N = len(A)
    B = np.eye(32, dtype=complex)
    B[:N, :N] = A
    used = np.zeros(32, bool)
    my_step = np.zeros(32, int)
    for c in range(32):
        f = B[:, 0].copy()
        mag = np.where(used, -1.0, np.abs(f.real) + np.abs(f.imag))
        p = int(np.argmax(mag))
        used[p] = True
        my_step[p] = c
        inv = 1.0 / f[p]
        g = f * inv
        g[p] = 1.0 - inv
        prow = B[p, :].copy()                          # raw pivot row, registers 1..31 are broadcast
        new = np.empty_like(B)
        new[:, :31] = B[:, 1:] - np.outer(g, prow[1:])  # results land one register down
        new[:, 31] = -g
        new[p, 31] = inv
        B = new
    q = my_step[my_step]                               # lane r: i = my_step[r], needed register = my_step[i]
    diag = np.zeros(32, complex)
    diag[my_step] = B[np.arange(32), q]
    return diag[:N]
