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
