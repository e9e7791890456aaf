"""CPU: the numpy oracle (oracle/) against golden vectors produced by the real reference
(tests/golden/*.npz, made by oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden, oracle_model, spec_for, GOLDEN_SPECS
from oracle.sk_oracle import sk_table, V_NAMES
from oracle.sktb_oracle import OracleModel, lorentz_dos


def test_sk_table_matches_reference():
    g = golden("sk_table")
    assert list(g["V_names"]) == list(V_NAMES)
    for V, d, T in zip(g["V"], g["lmn"], g["T"]):
        mine = sk_table({n: V[i] for i, n in enumerate(V_NAMES)}, *d)
        np.testing.assert_allclose(mine, T, rtol=0, atol=2e-15 * max(1.0, np.abs(T).max()))


def test_sk_table_unknown_key_raises():
    with pytest.raises(TypeError):
        sk_table({"V_xyz": 1.0}, 1, 0, 0)


def test_kmesh_bit_exact():
    g = golden("kmesh")
    for i in range(5):
        assert np.array_equal(OracleModel.kmesh(list(g[f"nk{i}"])), g[f"mesh{i}"])


@pytest.mark.parametrize("name", list(GOLDEN_SPECS))
def test_model_tables_and_spectrum(name):
    g = golden(name)
    m = oracle_model(spec_for(name))
    assert np.array_equal(m.bond_mat, g["bond_mat"])
    assert np.array_equal(np.array(m.bonds, dtype=np.int32).reshape(-1, 3), g["bond_list"])
    bv = np.array([m.dist_mat_vec[i, a, b] for i, a, b in m.bonds]).reshape(-1, 3)
    assert np.array_equal(bv, g["bond_vec"])                      # same np.dot per pair -> same bits here
    H = np.zeros(tuple(g["H_wo_g_shape"]), dtype=complex)
    H[tuple(g["H_wo_g_idx"].T)] = g["H_wo_g_val"]
    np.testing.assert_allclose(m.H_wo_g, H, rtol=0, atol=1e-14)
    assert np.array_equal(m.soc_mat, g["soc_mat"])
    soc = bool(g["soc"])
    k = g["k_test"]
    for i in range(3):
        np.testing.assert_allclose(m.get_ham(k[i], l_soc=soc), g["ham_k"][i], rtol=0, atol=1e-13)
    if bool(g["raises"]):
        with pytest.raises(Exception, match="not hermitian"):
            m.solve_kpath(list(k[:2]), soc=soc)
        return
    ev = m.solve_kpath(list(k), soc=soc)
    np.testing.assert_allclose(ev, g["evals"], rtol=0, atol=1e-11)
    if "evals_nosoc" in g:
        np.testing.assert_allclose(m.solve_kpath(list(k[:8]), soc=False), g["evals_nosoc"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("name", ["C1_cubic_sp3", "C2_graphene_pz", "C3_sb_buckled", "sp_chain", "C5_zigzag_8",
                                  "T_ce_spdf_1atom"])
def test_kpath_bit_exact(name):
    g = golden(name)
    m = oracle_model(spec_for(name))
    i = 0
    while f"path{i}_in" in g:
        pts, lens, spl = m.get_kpts(g[f"path{i}_in"].tolist(), int(g[f"path{i}_nk"]))
        assert np.array_equal(np.array([np.asarray(p, float) for p in pts]), g[f"path{i}_k"])
        assert np.array_equal(lens, g[f"path{i}_len"])
        assert np.array_equal(spl, g[f"path{i}_spl"])
        i += 1
    assert i > 0


def test_known_answers_quickstart():
    """Quick Start bands at Gamma / R and the path length."""
    g = golden("C1_cubic_sp3")
    m = oracle_model(spec_for("C1_cubic_sp3"))
    pts, lens, _ = m.get_kpts([[0, 0, 0], [0.5, 0.5, 0.5]], 50)
    assert len(pts) == 51 and lens[-1] == 0.24743582965269698
    ev = m.solve_kpath([pts[0], pts[-1]])
    np.testing.assert_allclose(ev[:, 0], [-3, -3] + [2.3] * 6, atol=1e-12)
    np.testing.assert_allclose(ev[:, 1], [0.7] * 6 + [3, 3], atol=1e-12)


@pytest.mark.parametrize("name", ["C1_cubic_sp3", "C2_graphene_pz", "C3_sb_buckled", "C5_zigzag_8", "lowsym_spd_laws"])
def test_dos_ldos_inverse_path(name):
    g = golden(name)
    m = oracle_model(spec_for(name))
    soc = bool(g["soc"])
    if name == "C2_graphene_pz":   # 400 energies x 64 k inverses is slow in pure python: subsample energies
        sel = slice(None, None, 16)
    else:
        sel = slice(None)
    d = m.dos(g["dos_E"][sel], list(g["dos_nk"]), float(g["dos_eta"]), soc)
    np.testing.assert_allclose(d, g["dos"][sel], rtol=1e-10)
    if "ldos" in g:
        orbs = list(g["ldos_orbs"]) or None
        l = m.ldos(g["ldos_E"], list(g["ldos_atoms"]), orbs, list(g["ldos_nk"]), float(g["ldos_eta"]), soc)
        np.testing.assert_allclose(l, g["ldos"], rtol=1e-10)
    if "edge_dos" in g:
        e = m.edge_dos(g["edge_E"], list(g["edge_atoms"]), nk=int(g["edge_nk"]), eta=float(g["edge_eta"]), soc=soc)
        np.testing.assert_allclose(e, g["edge_dos"], rtol=1e-10)


def test_lorentz_closed_form_equals_reference_dos():
    """The identity the CUDA path relies on, checked against the REFERENCE's inverse-based DOS."""
    for name in ["C2_graphene_pz", "C3_sb_buckled", "C5_zigzag_8"]:
        g = golden(name)
        m = oracle_model(spec_for(name))
        ev = m.solve_kpath(list(m.kmesh(list(g["dos_nk"]))), soc=bool(g["soc"]))
        np.testing.assert_allclose(lorentz_dos(ev, g["dos_E"], float(g["dos_eta"])), g["dos"], rtol=1e-10)


# Gaussian get_dos and Forces.get_total_energy of the real reference (tests/golden/next_rows.npz)
NEXT_MODELS = {"C1": "C1_cubic_sp3", "C3": "C3_sb_buckled", "C2": "C2_graphene_pz", "spchain": "sp_chain"}


def test_next_rows_oracle_vs_reference():
    from oracle import sktb_oracle as O
    g = golden("next_rows")
    E = g["energy"]
    np.testing.assert_allclose(O.gauss_dos(g["eig_in"], E, 0.2), g["eig_gdos"], rtol=1e-12, atol=1e-300)
    for tag, name in NEXT_MODELS.items():
        spec = spec_for(name)
        m = oracle_model(spec)
        ne, nk = int(g[tag + "_etot_args"][0]), [int(x) for x in g[tag + "_etot_args"][1:]]
        e_band = O.total_energy_band(m, ne, nk=nk, soc=spec["soc"])
        assert abs(e_band - g[tag + "_etot"][1]) < 1e-10, (tag, e_band, g[tag + "_etot"])
        assert g[tag + "_etot"][2] == 0.0                                    # no repulsive potential in these models
        if tag + "_gdos" in g:
            d = O.get_dos(m, E, w=0.15, nk=[int(x) for x in g[tag + "_gdos_nk"]])
            np.testing.assert_allclose(d, g[tag + "_gdos"], rtol=1e-9, atol=1e-12)
        if tag + "_emean" in g:                                              # Hamiltonian.total_energy (ARPACK in the reference)
            fb, nk1, dim = (int(x) for x in g[tag + "_emean_args"])
            assert abs(O.total_energy_mean(m, fb, nk1, dim, soc=spec["soc"]) - float(g[tag + "_emean"])) < 1e-9



def _fuzz_moment(ev, pw):
    """Rotation-invariant (within degenerate subspaces) summary of eigenvectors: sum_n |U_nkc|^2 / (1 + eps_n^2)."""
    return np.einsum("nkc,nk->kc", pw, 1.0 / (1.0 + np.asarray(ev) ** 2))


def test_fuzz_oracle_vs_reference():
    """48 seeded random models (orbital subsets, partial shells, S orbital, laws, periodicities): the oracle against the
    reference's eigenvalues / eigenvector weights, and the same Hermiticity-gate exceptions."""
    from pysktb_b200 import workloads as W
    g = golden("fuzz")
    for seed in range(int(g["n_seeds"])):
        spec = W.random_spec(seed)
        om = oracle_model(spec)
        ks = g[f"k_{seed}"]
        if int(g[f"raises_{seed}"]):
            with pytest.raises(Exception, match="not hermitian"):
                om.solve_kpath(list(ks), soc=spec["soc"])
            continue
        ev, vec = om.solve_kpath(list(ks), eig_vectors=True, soc=spec["soc"])
        np.testing.assert_allclose(ev, g[f"ev_{seed}"], rtol=0, atol=1e-10, err_msg=spec["name"])
        np.testing.assert_allclose(_fuzz_moment(ev, np.abs(vec) ** 2), _fuzz_moment(g[f"ev_{seed}"], g[f"pw_{seed}"]),
                                   rtol=1e-8, atol=1e-10, err_msg=spec["name"])


NONHERM = {"ce_spf_example": ("ce_spf_example", {}), "T": ("ce_spdf_1atom", {}), "C4": ("ce_spdf_2atom", {})}
# max_E |DOS(lower-triangle spectrum) / DOS(reference's dense inverse) - 1| on the fixture's mesh: the documented deviation
# for f-orbital models, whose H(k) is NOT Hermitian (|H - H^dagger| ~ 1.4 eV, purely imaginary)
NONHERM_DEVIATION = {"ce_spf_example": 0.9004073575305005, "T": 0.9486637184619495, "C4": 0.6988516850186401}


@pytest.mark.parametrize("tag", ["ce_spf_example", "T"])
def test_nonhermitian_dos_deviation_recorded(tag):
    """For f-orbital models the reference's inverse-based DOS (greens.py:70) is not an eigenvalue sum; the fixture holds
    the reference's numbers for both readings.  The oracle follows the reference on both; the deviation between them is
    the number DESIGN.md records (the CUDA path returns the lower-triangle reading, tests/test_gpu_parity.py)."""
    from pysktb_b200 import workloads as W
    g = golden("nonherm_dos")
    fn, kw = NONHERM[tag]
    spec = W.ALL[fn](**kw)
    m = oracle_model(spec)
    E, nk, eta = g[tag + "_E"], [int(x) for x in g[tag + "_nk"]], float(g[tag + "_eta"])
    np.testing.assert_allclose(m.dos(E, nk, eta, spec["soc"]), g[tag + "_dos_inv"], rtol=1e-10)          # the reference's inverse
    ev = m.solve_kpath(list(m.kmesh(nk)), soc=spec["soc"])
    low = lorentz_dos(ev, E, eta)
    np.testing.assert_allclose(low, g[tag + "_dos_lower"], rtol=1e-10)                                    # its lower-triangle spectrum
    dev = float(np.abs(low / g[tag + "_dos_inv"] - 1).max())
    assert abs(dev - NONHERM_DEVIATION[tag]) < 1e-8 and float(g[tag + "_asym"]) > 1.0


NONHERM_GF = {"ce_spf_example": ("ce_spf_example", {}), "T": ("ce_spdf_1atom", {}), "C4": ("ce_spdf_2atom", {}),
              "d_soc": ("d_soc_nonhermitian", {})}


@pytest.mark.parametrize("tag", ["ce_spf_example", "T", "d_soc"])
def test_oracle_inverse_greens_function_vs_reference_on_nonhermitian_models(tag):
    """The oracle's dense-inverse Green's function (greens.py:70 restated) against the reference's numbers on models whose
    H(k) is not Hermitian (tests/golden/nonherm_gf.npz) - the checker of the CUDA path's `method="inverse"`."""
    from pysktb_b200 import workloads as W
    g = golden("nonherm_gf")
    fn, kw = NONHERM_GF[tag]
    spec = W.ALL[fn](**kw)
    m = oracle_model(spec)
    E, nk, eta, soc = g[tag + "_E"], [int(x) for x in g[tag + "_nk"]], float(g[tag + "_eta"]), spec["soc"]
    np.testing.assert_allclose(m.dos(E, nk, eta, soc), g[tag + "_dos"], rtol=1e-10)
    n_at = len(spec["atoms"])
    np.testing.assert_allclose(m.ldos(E, atom_indices=list(range(n_at))[::-1], nk=nk, eta=eta, soc=soc), g[tag + "_ldos"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.spectral(-0.4, g["kpt"], eta=0.05, soc=soc), g[tag + "_spectral"], rtol=0,
                               atol=1e-11 * np.abs(g[tag + "_spectral"]).max())
    assert float(g[tag + "_asym"]) > 0.1


def test_gauss_jordan_model_of_the_inverse_kernel():
    """tools/proto_gj.py, the numpy model of csrc/gf_inverse.cuh (implicit row pivoting, A^-1[i][m] = R[p_i][q_m]), against LAPACK -
    including matrices whose leading block is zero (no elimination order without row exchanges)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("proto_gj", os.path.join(os.path.dirname(__file__), "..", "tools", "proto_gj.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    r = np.random.default_rng(3)
    for N in (1, 2, 7, 32, 65):
        A = r.standard_normal((N, N)) + 1j * r.standard_normal((N, N))
        A[: N // 2, : N // 2] = 0.0
        assert np.abs(mod.inverse(A) - np.linalg.inv(A)).max() < 1e-11 * max(1.0, np.abs(np.linalg.inv(A)).max()) * N
    # lane-level model of the row-per-lane kernel: rotating register window, unscaled pivot-row broadcast with 1/piv folded into the
    # factors, read-out of G_ii from register my_step[my_step[r]] of lane r
    for N in (1, 5, 22, 32):
        A = r.standard_normal((N, N)) + 1j * r.standard_normal((N, N))
        A[: N // 2, : N // 2] = 0.0
        ref = np.diag(np.linalg.inv(A))
        assert np.abs(mod.warp_kernel_model(A) - ref).max() < 1e-11 * max(1.0, np.abs(np.linalg.inv(A)).max()) * N


def cluster_projectors(ev, vec, tol=1e-7):
    """Spectral projectors of one k-point: [(mean eigenvalue, P)] with P = sum over a degenerate cluster of u u^H
    (rows of `vec` are eigenvectors, the reference's layout); gauge- and basis-independent within a cluster."""
    out, start = [], 0
    for n in range(1, len(ev) + 1):
        if n == len(ev) or ev[n] - ev[n - 1] > tol:
            U = vec[start:n]
            out.append((float(np.mean(ev[start:n])), U.T @ U.conj()))
            start = n
    return out


def test_fuzz_projectors_oracle_vs_reference():
    """Every 4th fuzz model carries the reference's full eigenvectors: spectral projectors of every degenerate cluster."""
    from pysktb_b200 import workloads as W
    g = golden("fuzz")
    n_checked = 0
    for seed in range(0, int(g["n_seeds"]), 4):
        if int(g[f"raises_{seed}"]):
            continue
        spec = W.random_spec(seed)
        ev, vec = oracle_model(spec).solve_kpath(list(g[f"k_{seed}"]), eig_vectors=True, soc=spec["soc"])
        for q in range(ev.shape[1]):
            ref = cluster_projectors(g[f"ev_{seed}"][:, q], g[f"vec_{seed}"][:, q, :])
            got = cluster_projectors(ev[:, q], vec[:, q, :])
            assert len(ref) == len(got), spec["name"]
            for (e0, P0), (e1, P1) in zip(ref, got):
                assert abs(e0 - e1) < 1e-10 and np.abs(P0 - P1).max() < 1e-8, spec["name"]
                n_checked += 1
    assert n_checked > 100
