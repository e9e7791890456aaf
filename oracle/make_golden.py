"""Generate tests/golden/*.npz by running the REAL reference (build container only).

    python -m oracle.make_golden            # from the repo root; needs /root/reference

The fixtures pin (a) the numpy oracle (tests/test_oracle_golden.py, CPU) and (b) the CUDA path
(tests/test_gpu_parity.py, GPU box - where /root/reference does not exist).  Everything stored here is an
output of the reference's own code (`Hamiltonian(..., numba=0)`, `GreensFunction`, `get_hop_int`, the
reference's scaling laws) on the inputs stored beside it.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.sk_oracle import V_NAMES  # noqa: E402
from pysktb_b200 import workloads as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_kpoints(n, seed):
    r = np.random.default_rng(seed)
    special = np.array([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0.5], [1 / 3, 1 / 3, 0], [2 / 3, 1 / 3, 0], [0.1, 0.2, 0.3]])
    return np.concatenate([special, r.random((n - len(special), 3)) * 2 - 0.5])


def model_fixture(ref, spec, nk_test=24, n_vec=4, paths=(), dos=None, ldos=None, edge=None):
    st, ham = W.instantiate(spec, ref, ref.scaling)
    soc = spec["soc"]
    out = {}
    out["images"] = np.array(list(__import__("itertools").product(
        *[(np.arange(3) - 1) if p else [0] for p in st.periodicity])))
    out["bond_mat"] = st.bond_mat
    bl = []
    mi = st.max_image
    for a1 in range(len(st.atoms)):                 # the visiting order of calc_ham_wo_k
        for a2 in range(len(st.atoms)):
            for im in range(mi):
                if st.bond_mat[im, a1, a2]:
                    bl.append((im, a1, a2))
    out["bond_list"] = np.array(bl, dtype=np.int32).reshape(-1, 3)
    out["bond_vec"] = np.array([st.dist_mat_vec[i, a, b] for i, a, b in bl]).reshape(-1, 3)
    nz = np.nonzero(ham.H_wo_g)
    out["H_wo_g_idx"] = np.stack(nz, axis=1).astype(np.int32)
    out["H_wo_g_val"] = ham.H_wo_g[nz]
    out["H_wo_g_shape"] = np.array(ham.H_wo_g.shape)
    out["soc_mat"] = ham.soc_mat
    out["n_orbitals"] = np.array(ham.n_orbitals)
    out["soc"] = np.array(soc)
    k = test_kpoints(nk_test, seed=len(spec["name"]))
    out["k_test"] = k
    out["ham_k"] = np.array([ham.get_ham(kk, l_soc=soc) for kk in k[:3]])
    out["g_mat0_nz"] = np.array(ham.calc_g(k[5])[nz])
    try:
        vals, vecs = quiet(ham.solve_kpath, list(k[:n_vec]), eig_vectors=True, soc=soc, parallel=0)
        out["evals"] = quiet(ham.solve_kpath, list(k), eig_vectors=False, soc=soc, parallel=0)
        out["evecs"] = vecs
        out["evals_vec"] = vals      # eigh and eigvalsh take different LAPACK paths: equal to ~1e-15, not bitwise
        assert np.allclose(vals, out["evals"][:, :n_vec], rtol=0, atol=1e-12)
        out["raises"] = np.array(False)
    except Exception as e:  # the reference's Hermiticity gate
        assert "not hermitian" in str(e)
        out["raises"] = np.array(True)
    if soc and ham.n_orbitals <= 16:   # also the no-SOC reading of the same model
        out["evals_nosoc"] = quiet(ham.solve_kpath, list(k[:8]), eig_vectors=False, soc=False, parallel=0)
    for pi, (path, nk) in enumerate(paths):
        kp, kl, sp = ham.get_kpts(path, nk)
        out[f"path{pi}_in"] = np.array(path, dtype=float)
        out[f"path{pi}_nk"] = np.array(nk)
        out[f"path{pi}_k"] = np.array([np.asarray(x, dtype=float) for x in kp])
        out[f"path{pi}_len"] = kl
        out[f"path{pi}_spl"] = sp
    gf = ref.GreensFunction(ham)
    if dos is not None:
        en, nk, eta = dos
        out["dos_E"], out["dos_nk"], out["dos_eta"] = en, np.array(nk), np.array(eta)
        out["dos"] = gf.dos(en, nk=nk, eta=eta, soc=soc, parallel=False)
    if ldos is not None:
        en, nk, eta, atoms_sel, orb_sel = ldos
        out["ldos_E"], out["ldos_nk"], out["ldos_eta"] = en, np.array(nk), np.array(eta)
        out["ldos_atoms"] = np.array(atoms_sel)
        out["ldos_orbs"] = np.array([] if orb_sel is None else orb_sel, dtype=np.int64)
        out["ldos"] = gf.ldos(en, atom_indices=atoms_sel, orbital_indices=orb_sel, nk=nk, eta=eta, soc=soc,
                              parallel=False)
        byorb = gf.ldos_by_orbital(en, atom_index=atoms_sel[0], nk=nk, eta=eta, soc=soc, parallel=False)
        out["ldos_by_orbital"] = np.array([byorb[o] for o in st.atoms[atoms_sel[0]].orbitals])
    if edge is not None:
        en, nk, eta, surf = edge
        sgf = ref.SurfaceGreensFunction(ham, surface_atoms=surf)
        out["edge_E"], out["edge_nk"], out["edge_eta"], out["edge_atoms"] = en, np.array(nk), np.array(eta), np.array(surf)
        out["edge_dos"] = sgf.edge_dos(en, nk=nk, eta=eta, soc=soc, parallel=False)
        kv = np.linspace(0, 1, 7)
        out["edge_spec_k"] = kv
        out["edge_spec"] = sgf.edge_spectral_kpath(kv, en[::4], eta=eta, soc=soc)
        kd, spec_map, sp = gf.spectral_kpath([[0, 0, 0], [0.5, 0, 0]], en[::4], nk=5, eta=eta, soc=soc)
        out["spec_kpath"] = spec_map
    np.savez_compressed(os.path.join(OUT, spec["name"] + ".npz"), **out)
    print(spec["name"], "N =", ham.n_orbitals * (2 if soc else 1), "bonds =", len(bl),
          "raises" if out["raises"] else "ok")


def sk_fixture(ref):
    r = np.random.default_rng(2024)
    dirs = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1], [-1, 0, 0], [1, 1, 0], [1, -1, 0], [1, 1, 1], [-1, 1, -1],
            [0, 1, 1], [1, 0, 1], [1e-6, 0, 1], [3e-6, 4e-6, 1]]
    dirs = [np.array(d, float) / np.linalg.norm(d) for d in dirs]
    dirs += [v / np.linalg.norm(v) for v in r.standard_normal((51, 3))]
    V = np.zeros((len(dirs), len(V_NAMES)))
    T = np.zeros((len(dirs), 17, 17))
    for b, d in enumerate(dirs):
        mask = r.random(len(V_NAMES)) < (1.0 if b % 3 else 0.5)
        V[b] = np.where(mask, r.uniform(-1.5, 1.5, len(V_NAMES)), 0.0)
        kw = {n: V[b, i] for i, n in enumerate(V_NAMES)}
        with np.errstate(all="ignore"):
            T[b] = np.array(ref.get_hop_int(**kw, l=d[0], m=d[1], n=d[2]), dtype=float)
    np.savez_compressed(os.path.join(OUT, "sk_table.npz"), V=V, lmn=np.array(dirs), T=T, V_names=np.array(V_NAMES))
    print("sk_table", T.shape)


def kmesh_fixture(ref):
    st, ham = W.instantiate(W.graphene_pz(), ref, ref.scaling)
    gf = ref.GreensFunction(ham)
    out = {}
    for i, nk in enumerate([(7, 5, 3), (12, 1, 1), (3, 49, 1), (1, 1, 1), (6, 6, 6)]):
        out[f"nk{i}"] = np.array(nk)
        out[f"mesh{i}"] = gf._generate_kmesh(list(nk))
    np.savez_compressed(os.path.join(OUT, "kmesh.npz"), **out)
    print("kmesh ok")


def next_rows_fixture(ref):
    """Hamiltonian.get_dos (Gaussian) and Forces.get_total_energy of the real reference."""
    from pysktb.forces import Forces
    out = {}
    E = np.linspace(-4.0, 4.0, 33)
    for name, spec, nk, ne in (("C1", W.cubic_sp3(), [3, 3, 3], 4), ("C3", W.sb_buckled(), [4, 3, 1], 6),
                               ("C2", W.graphene_pz(), [5, 4, 1], 2), ("spchain", W.sp_chain(), [6, 1, 1], 1)):
        st, ham = W.instantiate(spec, ref, ref.scaling)
        # solve_k -> solve_kpath(parallel=1) would spawn loky workers that cannot import the reference shell
        # (ref_loader): run the identical serial branch (parallel=0, hamiltonian.py:186-209) instead
        orig = ham.solve_kpath
        ham.solve_kpath = lambda k_list=None, eig_vectors=False, soc=True, parallel=0, _o=orig: _o(k_list, eig_vectors, soc, 0)
        if spec["soc"]:                                       # get_dos always solves with soc (solve_k default)
            out[name + "_gdos"] = np.asarray(quiet(ham.get_dos, E, w=0.15, nk=nk))
            out[name + "_gdos_nk"] = np.array(nk)
        f = Forces(ham)
        tot = quiet(f.get_total_energy, ne, nk=nk, soc=spec["soc"])
        out[name + "_etot"] = np.array(tot, dtype=float)
        out[name + "_etot_args"] = np.array([ne] + nk)
    # Hamiltonian.total_energy -> utils/energitics.get_totalenergy (ARPACK per k, fanned out with joblib: run the same
    # calls serially - loky workers cannot import the reference shell)
    from pysktb.utils import energitics
    energitics.Parallel = lambda n_jobs=None: (lambda it: [f(*a, **kw) for f, a, kw in it])
    for name, spec, args in (("C1", W.cubic_sp3(), dict(filled_band=3, nk=3, dim=3)), ("C3", W.sb_buckled(), dict(filled_band=0, nk=4, dim=2)),
                             ("spchain", W.sp_chain(), dict(filled_band=0, nk=7, dim=1))):   # (N = 2 models: ARPACK needs k < N - 1, the reference raises)
        st, ham = W.instantiate(spec, ref, ref.scaling)
        out[name + "_emean"] = np.array(quiet(energitics.get_totalenergy, ham, soc=spec["soc"], **args))
        out[name + "_emean_args"] = np.array([args["filled_band"], args["nk"], args["dim"]])
    r = np.random.default_rng(7)
    ev = r.standard_normal(57) * 2
    D = 0
    for i in ev:                                              # the reference's loop on caller-supplied eigenvalues
        D = D + np.exp(-((E - i) ** 2) / (2 * 0.2 ** 2)) / (np.pi * 0.2 * np.sqrt(2))
    out["eig_in"] = ev
    out["eig_gdos"] = D
    out["energy"] = E
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **out)
    print("next_rows", {k: v.shape for k, v in out.items()})


FUZZ_SEEDS = 48


def fuzz_fixture(ref):
    """Seeded random models (workloads.random_spec): the reference's eigenvalues (and eigenvector Gram residuals) at
    three random k each, or a flag when its Hermiticity gate raises.  tests/golden/fuzz.npz"""
    out = {"n_seeds": np.array(FUZZ_SEEDS)}
    for seed in range(FUZZ_SEEDS):
        spec = W.random_spec(seed)
        ks = np.random.default_rng(seed).random((3, 3)) * 2 - 0.5
        st, ham = quiet(W.instantiate, spec, ref, ref.scaling, numba=0)
        out[f"k_{seed}"] = ks
        try:
            ev, vec = quiet(ham.solve_kpath, list(ks), eig_vectors=True, soc=spec["soc"], parallel=0)
            out[f"ev_{seed}"] = np.asarray(ev)
            out[f"pw_{seed}"] = np.abs(np.asarray(vec)) ** 2              # |U|^2 per (band, k, component)
            if seed % 4 == 0:
                out[f"vec_{seed}"] = np.asarray(vec)                      # full eigenvectors [band, k, comp]: projector tests
            out[f"raises_{seed}"] = np.array(0)
        except Exception as exc:                                         # noqa: BLE001 - the reference raises bare Exception
            assert "not hermitian" in str(exc)
            out[f"raises_{seed}"] = np.array(1)
    np.savez_compressed(os.path.join(OUT, "fuzz.npz"), **out)
    print("fuzz.npz:", FUZZ_SEEDS, "models,", sum(int(out[f"raises_{s}"]) for s in range(FUZZ_SEEDS)), "raise")


from oracle.make_golden_cases import LAW_CASES, LAW_DISTANCES  # noqa: E402


def laws_fixture(ref):
    """scaling.py:312-447 - Polynomial / Tabulated / Custom (and Constant) through the reference's own `__call__`, scalar
    calls as `System.get_hop_params` makes them (system.py:206-210).  tests/golden/laws_misc.npz"""
    out = {"d": LAW_DISTANCES}
    for tag, (kind, kw) in LAW_CASES.items():
        lawobj = getattr(ref.scaling, kind)(**kw)
        out[tag] = np.array([float(lawobj(float(x))) for x in LAW_DISTANCES])
    np.savez_compressed(os.path.join(OUT, "laws_misc.npz"), **out)
    print("laws_misc", {k: v.shape for k, v in out.items()})


def nonherm_dos_fixture(ref):
    """f-orbital models: H(k) is not Hermitian, so the reference's inverse-based DOS (greens.py:70) is not an
    eigenvalue sum.  The fixture records the reference's numbers; the tests assert the size of the deviation of the
    lower-triangle-spectrum DOS this package returns.  tests/golden/nonherm_dos.npz"""
    out = {}
    E = np.linspace(-4.0, 4.0, 17)
    for tag, spec, nk, eta in (("ce_spf_example", W.ce_spf_example(), [3, 3, 3], 0.1), ("T", W.ce_spdf_1atom(), [3, 3, 2], 0.1),
                               ("C4", W.ce_spdf_2atom(), [2, 2, 2], 0.1)):
        st, ham = quiet(W.instantiate, spec, ref, ref.scaling, numba=0)
        gf = ref.GreensFunction(ham)
        out[tag + "_E"], out[tag + "_nk"], out[tag + "_eta"] = E, np.array(nk), np.array(eta)
        out[tag + "_dos_inv"] = quiet(gf.dos, E, nk=nk, eta=eta, soc=spec["soc"], parallel=False)
        mesh = gf._generate_kmesh(nk)
        ev = quiet(ham.solve_kpath, list(mesh), eig_vectors=False, soc=spec["soc"], parallel=0)   # the reference's own lower-triangle spectrum
        lor = ((eta / np.pi) / ((E[:, None, None] - ev[None]) ** 2 + eta ** 2)).sum(axis=(1, 2)) / len(mesh)
        out[tag + "_dos_lower"] = lor
        out[tag + "_asym"] = np.array(max(np.abs(h - h.conj().T).max() for h in (ham.get_ham(k, l_soc=spec["soc"]) for k in mesh)))
        print(tag, "max |H - H^dagger| =", float(out[tag + "_asym"]), " max rel |dos_lower / dos_inv - 1| =",
              float(np.abs(lor / out[tag + "_dos_inv"] - 1).max()))
    np.savez_compressed(os.path.join(OUT, "nonherm_dos.npz"), **out)


NONHERM_GF = {"ce_spf_example": ("ce_spf_example", {}), "T": ("ce_spdf_1atom", {}), "C4": ("ce_spdf_2atom", {}),
              "d_soc": ("d_soc_nonhermitian", {})}


def nonherm_gf_fixture(ref):
    """Every Green's-function entry point of the reference (greens.py:46-376, 431-557) on models whose H(k) is not Hermitian
    (f shells: imaginary asymmetry; d-shell SOC: the model solve_kpath refuses): the dense-inverse numbers the CUDA path's
    `method="inverse"` (csrc/gf_inverse.cuh) has to reproduce.  tests/golden/nonherm_gf.npz"""
    out = {}
    E = np.linspace(-3.0, 3.0, 9)
    kpt = np.array([0.1, 0.2, 0.3])
    path = [[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.5, 0.5, 0.0]]
    kvals = np.array([0.0, 0.13, 0.5, 0.77])
    for tag, (fn, kw) in NONHERM_GF.items():
        spec = W.ALL[fn](**kw)
        st, ham = quiet(W.instantiate, spec, ref, ref.scaling, numba=0)
        soc = spec["soc"]
        gf = ref.GreensFunction(ham)
        n_at = len(st.atoms)
        nk = [2, 2, 2] if tag != "C4" else [2, 1, 2]
        out[tag + "_E"], out[tag + "_nk"], out[tag + "_eta"] = E, np.array(nk), np.array(0.07)
        out[tag + "_dos"] = quiet(gf.dos, E, nk=nk, eta=0.07, soc=soc, parallel=False)
        out[tag + "_ldos"] = quiet(gf.ldos, E, atom_indices=list(range(n_at))[::-1], nk=nk, eta=0.07, soc=soc, parallel=False)
        out[tag + "_ldos_orb"] = quiet(gf.ldos, E, atom_indices=[0], orbital_indices=[0, 2, 5], nk=nk, eta=0.07, soc=soc, parallel=False)
        byorb = quiet(gf.ldos_by_orbital, E, atom_index=n_at - 1, nk=nk, eta=0.07, soc=soc, parallel=False)
        out[tag + "_by_orbital"] = np.array([byorb[o] for o in st.atoms[n_at - 1].orbitals])
        out[tag + "_retarded"] = gf.retarded(0.3, kpt, eta=0.05, soc=soc)
        out[tag + "_spectral"] = gf.spectral(-0.4, kpt, eta=0.05, soc=soc)
        kd, amap, spl = quiet(gf.spectral_kpath, path, E, nk=3, eta=0.07, soc=soc)
        out[tag + "_kpath_map"], out[tag + "_kpath_dist"] = amap, kd
        sg = ref.SurfaceGreensFunction(ham, surface_atoms=[n_at - 1])
        kp = np.array([[k, 0.0, 0.0] for k in kvals])
        out[tag + "_edge_dos"] = quiet(sg.edge_dos, E, k_points=kp, eta=0.07, soc=soc, parallel=False)
        out[tag + "_edge_map"] = quiet(sg.edge_spectral_kpath, kvals, E, eta=0.07, soc=soc)
        out[tag + "_surface_spectral"] = np.array(sg.surface_spectral(0.2, kpt, eta=0.05, soc=soc))
        Hk = ham.get_ham(kpt, l_soc=soc)
        out[tag + "_asym"] = np.array(np.abs(Hk - Hk.conj().T).max())
        print(tag, "N =", Hk.shape[0], "max |H - H^dagger| =", float(out[tag + "_asym"]))
    out["kpt"], out["kvals"], out["path"] = kpt, kvals, np.array(path)
    np.savez_compressed(os.path.join(OUT, "nonherm_gf.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    if "--round2-only" in sys.argv:
        laws_fixture(ref)
        nonherm_dos_fixture(ref)
        model_fixture(ref, W.lowsym_spd(laws="misc"), nk_test=12, dos=(np.linspace(-3, 3, 13), [2, 2, 2], 0.1),
                      ldos=(np.linspace(-3, 3, 7), [2, 2, 1], 0.1, [2, 1], None))
        fuzz_fixture(ref)
        return
    if "--nonherm-gf-only" in sys.argv:
        nonherm_gf_fixture(ref)
        return
    if "--next-rows-only" in sys.argv:
        next_rows_fixture(ref)
        return
    if "--fuzz-only" in sys.argv:
        fuzz_fixture(ref)
        return
    fuzz_fixture(ref)
    next_rows_fixture(ref)
    sk_fixture(ref)
    kmesh_fixture(ref)
    E = np.linspace
    gpath = [[0.0, 0.0, 0], [2.0 / 3.0, 1.0 / 3.0, 0], [0.5, 0.5, 0], [0.0, 0.0, 0]]
    model_fixture(ref, W.cubic_sp3(), paths=[([[0, 0, 0], [0.5, 0.5, 0.5]], 50)],
                  dos=(E(-4, 4, 17), [3, 3, 3], 0.1), ldos=(E(-4, 4, 9), [3, 3, 2], 0.1, [0], [1, 3]))
    model_fixture(ref, W.graphene_pz(), paths=[([[0, 0, 0], [0.5, 0, 0], [1 / 3, 1 / 3, 0], [0, 0, 0]], 4), (gpath, 100)],
                  dos=(E(-8, 8, 400), [8, 8, 1], 0.1), ldos=(E(-8, 8, 41), [6, 6, 1], 0.1, [0, 1], None))
    model_fixture(ref, W.sb_buckled(), paths=[(gpath, 100)], dos=(E(-2.5, 2.5, 21), [4, 4, 1], 0.05),
                  ldos=(E(-2.5, 2.5, 11), [3, 3, 1], 0.05, [1, 0], [0, 2]))
    model_fixture(ref, W.ce_spdf_1atom(), nk_test=16, paths=[([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0]], 10)])
    model_fixture(ref, W.ce_spdf_2atom(), nk_test=10, n_vec=2)
    model_fixture(ref, W.ce_spf_example(), nk_test=12)
    model_fixture(ref, W.zigzag_ribbon(8), nk_test=12, paths=[([[0, 0, 0], [0.5, 0, 0], [1, 0, 0]], 20)],
                  dos=(E(-3, 3, 31), [9, 1, 1], 0.08), edge=(E(-3, 3, 40), 11, 0.08, [0, 1, 14, 15]))
    model_fixture(ref, W.zigzag_ribbon(32), nk_test=8, n_vec=2)
    model_fixture(ref, W.sp_chain(), paths=[([[0.5, 0.0, 0.0], [0.0, 0.0, 0.0], [-0.5, 0.0, -0.0]], 100)])
    model_fixture(ref, W.lowsym_spd(laws=True), nk_test=12, dos=(E(-3, 3, 13), [2, 2, 2], 0.1),
                  ldos=(E(-3, 3, 7), [2, 2, 1], 0.1, [2, 1], None))
    model_fixture(ref, W.lowsym_spd(laws=False), nk_test=10)
    model_fixture(ref, W.lowsym_spd(laws="misc"), nk_test=12, dos=(E(-3, 3, 13), [2, 2, 2], 0.1),
                  ldos=(E(-3, 3, 7), [2, 2, 1], 0.1, [2, 1], None))
    laws_fixture(ref)
    nonherm_dos_fixture(ref)
    nonherm_gf_fixture(ref)
    model_fixture(ref, W.d_soc_nonhermitian(), nk_test=8)


if __name__ == "__main__":
    main()
