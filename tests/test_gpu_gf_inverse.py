"""GPU (B200): the dense-inverse Green's function (csrc/gf_inverse.cuh, K1+K5) through GreensFunction / SurfaceGreensFunction ->
C ABI, against the REFERENCE's own numbers on models whose H(k) is not Hermitian (tests/golden/nonherm_gf.npz, nonherm_dos.npz:
f shells - T and C4, the headline models, among them - and d-shell SOC), against the numpy oracle's inverse on seeded fuzz
models, and against the eigenpair route on Hermitian models (the two device paths must agree where both are exact).
Tolerance: 1e-8 relative (BASELINE.json north_star, projections / DOS / LDOS)."""
import numpy as np
import pytest

from conftest import golden, oracle_model

pytestmark = pytest.mark.gpu
RTOL = 1e-8


@pytest.fixture(scope="module")
def tb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pysktb_b200
    return pysktb_b200


_cache = {}


def model(tb, tag):
    from pysktb_b200 import workloads as W
    from test_oracle_golden import NONHERM_GF
    if tag not in _cache:
        fn, kw = NONHERM_GF[tag]
        spec = W.ALL[fn](**kw)
        _cache[tag] = (spec,) + tuple(W.instantiate(spec, tb, tb.scaling))
    return _cache[tag]


TAGS = ["ce_spf_example", "T", "C4", "d_soc"]


@pytest.mark.parametrize("tag", TAGS)
def test_auto_method_takes_the_inverse_for_nonhermitian_models(tb, tag):
    spec, st, ham = model(tb, tag)
    g = golden("nonherm_gf")
    assert ham.herm_defect > 1e-3 and float(g[tag + "_asym"]) > 1e-3
    assert tb.GreensFunction(ham)._use_inverse()
    assert not tb.GreensFunction(ham, method="eigh")._use_inverse()


def test_d_soc_model_without_soc_is_hermitian_and_takes_the_eigenpair_route(tb):
    """The d-SOC model's only defect is soc_mat: with soc=False its H(k) is Hermitian, `auto` takes the eigenpair route, and both
    routes agree."""
    spec, st, ham = model(tb, "d_soc")
    assert ham.herm_defect > 0.5 and ham.herm_defect_nosoc <= 1e-12
    gf = tb.GreensFunction(ham)
    assert gf._use_inverse(True) and not gf._use_inverse(False)
    E = np.linspace(-2, 2, 9)
    a = gf.dos(E, nk=[3, 2, 2], eta=0.08, soc=False)
    b = tb.GreensFunction(ham, method="inverse").dos(E, nk=[3, 2, 2], eta=0.08, soc=False)
    np.testing.assert_allclose(a, b, rtol=1e-9)


def test_hermitian_models_keep_the_eigenpair_route(tb):
    from pysktb_b200 import workloads as W
    for fn in ("graphene_pz", "cubic_sp3", "sb_buckled"):
        _, ham = W.instantiate(W.ALL[fn](), tb, tb.scaling)
        assert ham.herm_defect <= 1e-12, (fn, ham.herm_defect)
        assert not tb.GreensFunction(ham)._use_inverse()


@pytest.mark.parametrize("tag", TAGS)
def test_dos_ldos_vs_reference_inverse(tb, tag):
    spec, st, ham = model(tb, tag)
    g = golden("nonherm_gf")
    E, nk, eta, soc = g[tag + "_E"], [int(x) for x in g[tag + "_nk"]], float(g[tag + "_eta"]), spec["soc"]
    gf = tb.GreensFunction(ham)
    n_at = len(st.atoms)
    np.testing.assert_allclose(gf.dos(E, nk=nk, eta=eta, soc=soc), g[tag + "_dos"], rtol=RTOL)
    np.testing.assert_allclose(gf.ldos(E, atom_indices=list(range(n_at))[::-1], nk=nk, eta=eta, soc=soc), g[tag + "_ldos"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(gf.ldos(E, atom_indices=[0], orbital_indices=[0, 2, 5], nk=nk, eta=eta, soc=soc), g[tag + "_ldos_orb"],
                               rtol=RTOL, atol=1e-12)
    by = gf.ldos_by_orbital(E, atom_index=n_at - 1, nk=nk, eta=eta, soc=soc)
    got = np.array([by[o] for o in st.atoms[n_at - 1].orbitals])
    np.testing.assert_allclose(got, g[tag + "_by_orbital"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("tag", TAGS)
def test_retarded_and_spectral_matrix_vs_reference(tb, tag):
    spec, st, ham = model(tb, tag)
    g = golden("nonherm_gf")
    gf = tb.GreensFunction(ham)
    G = gf.retarded(0.3, g["kpt"], eta=0.05, soc=spec["soc"])
    ref = g[tag + "_retarded"]
    assert G.shape == ref.shape
    assert np.abs(G - ref).max() < RTOL * np.abs(ref).max()
    A = gf.spectral(-0.4, g["kpt"], eta=0.05, soc=spec["soc"])
    assert np.abs(A - g[tag + "_spectral"]).max() < RTOL * np.abs(g[tag + "_spectral"]).max()


@pytest.mark.parametrize("tag", TAGS)
def test_spectral_maps_and_surface_vs_reference(tb, tag):
    spec, st, ham = model(tb, tag)
    g = golden("nonherm_gf")
    E, eta, soc = g[tag + "_E"], float(g[tag + "_eta"]), spec["soc"]
    kd, amap, spl = tb.GreensFunction(ham).spectral_kpath([list(p) for p in g["path"]], E, nk=3, eta=eta, soc=soc)
    np.testing.assert_array_equal(kd, g[tag + "_kpath_dist"])
    np.testing.assert_allclose(amap, g[tag + "_kpath_map"], rtol=RTOL, atol=1e-12)
    sg = tb.SurfaceGreensFunction(ham, surface_atoms=[len(st.atoms) - 1])
    kp = np.array([[k, 0.0, 0.0] for k in g["kvals"]])
    np.testing.assert_allclose(sg.edge_dos(E, k_points=kp, eta=eta, soc=soc), g[tag + "_edge_dos"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(sg.edge_spectral_kpath(g["kvals"], E, eta=eta, soc=soc), g[tag + "_edge_map"], rtol=RTOL, atol=1e-12)
    assert abs(sg.surface_spectral(0.2, g["kpt"], eta=0.05, soc=soc) - float(g[tag + "_surface_spectral"])) < RTOL * abs(float(g[tag + "_surface_spectral"]))


@pytest.mark.parametrize("tag", ["ce_spf_example", "T", "C4"])
def test_headline_models_dos_equals_the_reference_inverse(tb, tag):
    """The round-1 fixture of the reference's inverse-based DOS on the f-orbital models: the default path now returns it."""
    from pysktb_b200 import workloads as W
    from test_oracle_golden import NONHERM
    g = golden("nonherm_dos")
    fn, kw = NONHERM[tag]
    spec = W.ALL[fn](**kw)
    _, ham = W.instantiate(spec, tb, tb.scaling)
    E, nk, eta = g[tag + "_E"], [int(x) for x in g[tag + "_nk"]], float(g[tag + "_eta"])
    np.testing.assert_allclose(tb.GreensFunction(ham).dos(E, nk=nk, eta=eta, soc=spec["soc"]), g[tag + "_dos_inv"], rtol=RTOL)


def test_inverse_equals_eigenpair_route_on_hermitian_models(tb):
    """Where H(k) is Hermitian both device paths are exact: DOS, LDOS, maps and G itself agree to 1e-9."""
    from pysktb_b200 import workloads as W
    for fn, nk in (("graphene_pz", [7, 5, 1]), ("cubic_sp3", [3, 2, 3]), ("sb_buckled", [4, 3, 1])):
        spec = W.ALL[fn]()
        st, ham = W.instantiate(spec, tb, tb.scaling)
        soc = spec["soc"]
        E = np.linspace(-4, 4, 23)
        a, b = tb.GreensFunction(ham, method="eigh"), tb.GreensFunction(ham, method="inverse")
        np.testing.assert_allclose(b.dos(E, nk=nk, eta=0.06, soc=soc), a.dos(E, nk=nk, eta=0.06, soc=soc), rtol=1e-9)
        np.testing.assert_allclose(b.ldos(E, nk=nk, eta=0.06, soc=soc), a.ldos(E, nk=nk, eta=0.06, soc=soc), rtol=1e-9, atol=1e-13)
        k = [0.21, 0.4, 0.0]
        Ga, Gb = a.retarded(0.37, k, eta=0.05, soc=soc), b.retarded(0.37, k, eta=0.05, soc=soc)
        assert np.abs(Ga - Gb).max() < 1e-9 * np.abs(Ga).max()


def test_fuzz_f_shell_models_vs_oracle_inverse(tb):
    """Seeded random models WITH f shells (the ones the eigenpair DOS test has to skip): DOS and LDOS against the oracle's dense
    inverse (greens.py:70 restated in oracle/sktb_oracle.py, pinned on the reference's fixtures)."""
    from pysktb_b200 import workloads as W
    n = 0
    for seed in range(64):
        spec = W.random_spec(seed)
        if not any(o.startswith("f") for _, _, orbs in spec["atoms"] for o in orbs):
            continue
        try:
            st, ham = W.instantiate(spec, tb, tb.scaling)
        except Exception:
            continue
        if ham.herm_defect <= 1e-12:
            continue
        om, E = oracle_model(spec), np.linspace(-3, 3, 7)
        nk = [2, 2, 1]
        gf = tb.GreensFunction(ham)
        assert gf._use_inverse()
        np.testing.assert_allclose(gf.dos(E, nk=nk, eta=0.1, soc=spec["soc"]), om.dos(E, nk, 0.1, spec["soc"]), rtol=RTOL, err_msg=spec["name"])
        at = list(range(len(st.atoms)))
        np.testing.assert_allclose(gf.ldos(E, atom_indices=at, nk=nk, eta=0.1, soc=spec["soc"]),
                                   om.ldos(E, atom_indices=at, nk=nk, eta=0.1, soc=spec["soc"]), rtol=RTOL, atol=1e-12, err_msg=spec["name"])
        n += 1
        if n >= 6:
            break
    assert n >= 3


@pytest.mark.parametrize("N", [1, 3, 31, 32, 33, 48, 64, 65, 96, 97, 118, 130])
def test_dense_inverse_kernels_on_random_matrices_vs_lapack(tb, N):
    """The three inverse kernels alone on random NON-Hermitian matrices against numpy's LAPACK inverse: the full G from the
    shared-memory kernel (matrix in shared memory up to N = 118, the global-workspace fallback above), -(1/pi) Im Tr G from the
    register-resident kernels (warp per matrix N <= 32, CTA per matrix N <= 96); half of the matrices have a zero leading block
    (no elimination order without row exchanges)."""
    from pysktb_b200 import workloads as W
    _, ham = W.instantiate(W.graphene_pz(), tb, tb.scaling)
    r = np.random.default_rng(N)
    nm = 5
    H = r.standard_normal((nm, N, N)) + 1j * r.standard_normal((nm, N, N))
    H[1::2, : max(1, N // 2), : max(1, N // 2)] = 0.0
    E, eta = np.array([0.2, -1.3, 0.0]), 0.05
    G = ham._gf_inverse_raw(H, E, eta, want="G")
    tr = ham._gf_inverse_raw(H, E, eta, want="trace")
    for m in range(nm):
        for ie, e in enumerate(E):
            ref = np.linalg.inv((e + 1j * eta) * np.eye(N) - H[m])
            assert np.abs(G[m, ie] - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()) * N, (N, m, ie)
            t_ref = -np.trace(ref).imag / np.pi
            assert abs(tr[m, ie] - t_ref) < 1e-10 * max(1.0, np.abs(ref).max()) * N, (N, m, ie, tr[m, ie], t_ref)


@pytest.mark.parametrize("tag,nk", [("T", [7, 6, 5]), ("C4", [4, 3, 3])])
def test_inverse_path_size_independent_properties(tb, tag, nk):
    """Larger meshes than the reference fixtures reach, through properties of the inverse itself: the atom-resolved LDOS rows
    (all orbitals, both spins) sum to the trace DOS; the per-orbital LDOS of every atom sums to that atom's row; a mesh given as an
    explicit k list (edge_dos route, un-normalised sum / n) equals the mesh route; splitting the k range in two and adding the
    partial sums reproduces the whole (the multi-GPU sharding arithmetic)."""
    import torch
    from pysktb_b200 import _lib
    spec, st, ham = model(tb, tag)
    soc = spec["soc"]
    E = np.linspace(-5, 5, 41)
    gf = tb.GreensFunction(ham)
    dos = gf.dos(E, nk=nk, eta=0.07, soc=soc)
    ld = gf.ldos(E, nk=nk, eta=0.07, soc=soc)
    np.testing.assert_allclose(ld.sum(axis=0), dos, rtol=1e-10)
    for ai in range(len(st.atoms)):
        by = gf.ldos_by_orbital(E, atom_index=ai, nk=nk, eta=0.07, soc=soc)
        np.testing.assert_allclose(sum(by.values()), ld[ai], rtol=1e-10)
    kmesh = gf._generate_kmesh(nk)
    n = ham.n_orbitals * (2 if soc else 1)
    res, nkp = gf._reduce(E, soc, 0.07, [list(range(n))], k_points=kmesh)
    np.testing.assert_allclose(res[0] / nkp, dos, rtol=1e-12)
    half = len(kmesh) // 2
    a, _ = gf._reduce(E, soc, 0.07, None, k_points=kmesh[:half])
    b, _ = gf._reduce(E, soc, 0.07, None, k_points=kmesh[half:])
    np.testing.assert_allclose((a[0] + b[0]) / len(kmesh), dos, rtol=1e-12)
    assert np.all(np.isfinite(dos))
