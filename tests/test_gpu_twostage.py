"""Two-stage tridiagonalisation (csrc/heev_twostage.cuh) against LAPACK, stage by stage: the Hermitian band after stage 1 and the
tridiagonal matrix after stage 2 must have the eigenvalues of the input matrix (lower-triangle reading, numpy's default UPLO='L',
pysktb/hamiltonian.py:119); sktb_heev above the size threshold must return them ascending.  Tolerance: 1e-10 absolute on O(10)
entries (north-star eigenvalue tolerance), checked relative to the matrix norm."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_ready():
    try:
        import torch
        return torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "pysktb_b200", "libsktb.so"))
    except Exception:
        return False


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _gpu_ready(), reason="needs a CUDA device and the built libsktb.so")]


@pytest.fixture(scope="module")
def ctx():
    import torch
    import pysktb_b200 as tb
    from pysktb_b200 import workloads as W, _lib
    _, ham = W.instantiate(W.cubic_sp3(), tb, tb.scaling)
    return torch, ham, _lib.lib(), _lib


def _band_to_dense(AB):
    N, w = AB.shape
    A = np.zeros((N, N), complex)
    for j in range(N):
        n = min(w, N - j)
        A[j:j + n, j] = AB[j, :n]
    return np.tril(A) + np.tril(A, -1).conj().T


@pytest.mark.parametrize("N,nm", [(32, 5), (33, 3), (47, 4), (64, 3), (97, 3), (130, 2), (352, 2), (528, 2)])
def test_band_and_tridiagonal_keep_the_spectrum(ctx, N, nm):
    torch, ham, lib, _lib = ctx
    g = torch.Generator(device="cuda").manual_seed(1000 + N)
    M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda", generator=g)      # not Hermitian: only the lower triangle may be read
    band = torch.zeros((nm, N, 32), dtype=torch.complex128, device="cuda")
    de = torch.zeros((nm, 2, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.sktb_heev_twostage(ham._plan, M.data_ptr(), N, nm, band.data_ptr(), de.data_ptr(), None))
    torch.cuda.synchronize()
    Mh, Bh, Dh = M.cpu().numpy(), band.cpu().numpy(), de.cpu().numpy()
    assert np.abs(Bh[:, :, 17:]).max() == 0.0                                             # stage 1 leaves a band of width 16
    for i in range(nm):
        ref = np.linalg.eigvalsh(Mh[i], UPLO="L")
        scale = np.abs(ref).max()
        assert np.abs(np.linalg.eigvalsh(_band_to_dense(Bh[i])) - ref).max() < 1e-12 * scale * N
        d, e = Dh[i, 0], Dh[i, 1, :N - 1]
        tri = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        assert np.abs(np.linalg.eigvalsh(tri) - ref).max() < 1e-12 * scale * N


def test_unsupported_size_is_an_error(ctx):
    torch, ham, lib, _lib = ctx
    M = torch.zeros((1, 16, 16), dtype=torch.complex128, device="cuda")
    assert lib.sktb_heev_twostage(ham._plan, M.data_ptr(), 16, 1, None, None, None) != 0


@pytest.mark.parametrize("N", [352, 400, 512])
def test_heev_large_eigenvalues(ctx, N):
    """sktb_heev takes the two-stage path from N = 352: random batches (more matrices than resident CTAs) and adversarial inputs."""
    torch, ham, lib, _lib = ctx
    rng = np.random.default_rng(N)
    A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)); A = A + A.conj().T
    cases = [np.zeros((N, N), complex),
             np.diag(np.repeat(np.array([-1.5, 0.0, 0.0, 2.0]), N // 4 + 1)[:N]).astype(complex),
             A * 1e-7, A * 1e3]
    B = A.copy(); h = N // 2; B[h:, :h] = 0; B[:h, h:] = 0
    cases.append(B)                                                                       # two decoupled blocks
    C = np.zeros((N, N), complex)
    for i in range(N - 1):
        C[i + 1, i] = C[i, i + 1] = 1.0 if i % 2 == 0 else 0.3                            # bipartite chain: symmetric spectrum, zero diagonal
    cases.append(C)
    Q, _ = np.linalg.qr(A)
    cases.append((Q * np.repeat(np.arange(N // 8 + 1, dtype=float), 8)[:N]) @ Q.conj().T)  # 8-fold exact degeneracies
    for _ in range(5):
        R = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        cases.append(R)                                                                   # non-Hermitian: lower-triangle reading
    M = torch.tensor(np.stack(cases), device="cuda")
    nm = M.shape[0]
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nm, ev.data_ptr(), None, None, None))
    torch.cuda.synchronize()
    assert "two-stage" in ham.last_kernel_info()
    got = ev.cpu().numpy()
    for i, mat in enumerate(cases):
        ref = np.linalg.eigvalsh(mat, UPLO="L")
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.all(np.diff(got[:, i]) >= 0)
        assert np.abs(got[:, i] - ref).max() <= 1e-11 * scale + 1e-30, (i, np.abs(got[:, i] - ref).max())


def test_heev_large_batch_matches_small_batch(ctx):
    """Persistent CTAs loop over matrices and the bulge chase pipelines sweeps through shared-memory counters: a batch larger than the
    resident grid must give bit-identical results to the same matrices solved alone."""
    torch, ham, lib, _lib = ctx
    N, nm = 352, 500
    g = torch.Generator(device="cuda").manual_seed(5)
    M = torch.randn((nm, N, N), dtype=torch.complex128, device="cuda", generator=g)
    ev = torch.empty((N, nm), dtype=torch.float64, device="cuda")
    _lib.check(lib.sktb_heev(ham._plan, M.data_ptr(), N, nm, ev.data_ptr(), None, None, None))
    for q in (0, 147, 148, 443, 444, 499):
        e1 = torch.empty((N, 1), dtype=torch.float64, device="cuda")
        _lib.check(lib.sktb_heev(ham._plan, M[q:q + 1].contiguous().data_ptr(), N, 1, e1.data_ptr(), None, None, None))
        torch.cuda.synchronize()
        assert torch.equal(e1[:, 0], ev[:, q])
        ref = np.linalg.eigvalsh(M[q].cpu().numpy(), UPLO="L")
        assert np.abs(ev[:, q].cpu().numpy() - ref).max() < 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("n_chains", [120, 256])
def test_ldos_rows_after_two_stage_against_lapack(ctx, n_chains):
    """Projections for N >= 208 take selected rows of U = Q1 Q2 Z from the two-stage reduction (block reflectors of stage 1, chase
    reflectors of stage 2, inverse iteration on the tridiagonal): LDOS of edge and bulk atoms of zigzag ribbons (240 / 512 pz orbitals,
    nearly degenerate edge states) against a Lorentzian sum over LAPACK eigenpairs of the device-built H(k); 1e-8 relative
    (north-star tolerance for eigenvector-derived projections)."""
    import pysktb_b200 as tb
    from pysktb_b200 import workloads as W
    _, ham = W.instantiate(W.zigzag_ribbon(n_chains), tb, tb.scaling)
    N = 2 * n_chains
    nk = 5
    k = np.stack([np.arange(nk) / nk, np.zeros(nk), np.zeros(nk)], axis=1)
    H = ham.get_ham_batch(k, l_soc=False)
    E = np.linspace(-3.2, 3.2, 45)
    eta = 0.05
    atoms = [0, 1, N // 2, N - 2, N - 1]
    got = tb.GreensFunction(ham).ldos(E, atom_indices=atoms, nk=[nk, 1, 1], eta=eta, soc=False)
    assert "rows_kernel" in ham.last_kernel_info(), ham.last_kernel_info()
    ref = np.zeros((len(atoms), len(E)))
    for q in range(nk):
        w, U = np.linalg.eigh(H[q])
        lor = (eta / np.pi) / ((E[None, :] - w[:, None]) ** 2 + eta ** 2)               # [n, E]
        for j, a in enumerate(atoms):
            ref[j] += (np.abs(U[a, :]) ** 2) @ lor
    ref /= nk
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-12)
    edge = tb.SurfaceGreensFunction(ham, [0, 1, N - 2, N - 1]).edge_dos(E, k_points=k, eta=eta, soc=False)
    np.testing.assert_allclose(edge, ref[[0, 1, 3, 4]].sum(axis=0), rtol=1e-8, atol=1e-12)


def test_ldos_rows_multi_chunk_sums_to_dos(ctx):
    """More k-points than one workspace chunk of the projection path (chunks are whole waves of the two-stage grids): the LDOS of ALL atoms
    (240 selected rows per k through rows_kernel) must add up to the eigenvalue-only DOS (two-stage + bisection), chunk boundaries included."""
    import pysktb_b200 as tb
    from pysktb_b200 import workloads as W
    _, ham = W.instantiate(W.zigzag_ribbon(120), tb, tb.scaling)
    gf = tb.GreensFunction(ham)
    E = np.linspace(-3.0, 3.0, 9)
    nk = [3001, 1, 1]
    ld = gf.ldos(E, atom_indices=list(range(240)), nk=nk, eta=0.07, soc=False)
    assert "rows_kernel" in ham.last_kernel_info()
    dos = gf.dos(E, nk=nk, eta=0.07, soc=False)
    assert "two-stage" in ham.last_kernel_info()
    np.testing.assert_allclose(ld.sum(axis=0), dos, rtol=1e-9)


def test_spectral_maps_after_two_stage_against_lapack(ctx):
    """A(k,E) maps with projections (edge_spectral_kpath: sktb_spectral with an index list) on the 256-orbital ribbon against LAPACK."""
    import pysktb_b200 as tb
    from pysktb_b200 import workloads as W
    _, ham = W.instantiate(W.zigzag_ribbon(128), tb, tb.scaling)
    N = 256
    kv = np.array([0.0, 0.21, 0.5, 0.77])
    k = np.stack([kv, np.zeros(4), np.zeros(4)], axis=1)
    H = ham.get_ham_batch(k, l_soc=False)
    E = np.linspace(-3.0, 3.0, 31)
    eta = 0.05
    atoms = [0, 1, N - 2, N - 1]
    got = tb.SurfaceGreensFunction(ham, atoms).edge_spectral_kpath(kv, E, eta=eta, soc=False)
    assert "rows_kernel" in ham.last_kernel_info(), ham.last_kernel_info()
    ref = np.zeros((4, len(E)))
    for q in range(4):
        w, U = np.linalg.eigh(H[q])
        lor = (eta / np.pi) / ((E[None, :] - w[:, None]) ** 2 + eta ** 2)
        ref[q] = (np.abs(U[atoms, :]) ** 2).sum(axis=0) @ lor
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-12)
