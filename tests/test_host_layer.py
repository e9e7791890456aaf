"""CPU: the product's host-side setup layer (Lattice/Atom/Structure/System) against golden vectors from the
real reference, the C-ABI library's symbol table against include/sktb.h, and the loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden, spec_for, GOLDEN_SPECS
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib


def build_structure(spec):
    lat = tb.Lattice(spec["lattice"], spec["a"])
    atoms = [tb.Atom(el, list(pos), list(orbs)) for el, pos, orbs in spec["atoms"]]
    st = tb.Structure(lat, atoms, periodicity=spec["periodicity"], bond_cut=spec["bond_cut"])
    sy = tb.System(st, {a.element: a.orbitals for a in st.atoms}, W.resolve(spec["params"], tb.scaling))
    return st, sy


@pytest.mark.parametrize("name", list(GOLDEN_SPECS))
def test_structure_and_system_tables(name):
    g = golden(name)
    st, sy = build_structure(spec_for(name))
    assert np.array_equal(st.images, g["images"])
    assert np.array_equal(st.bond_mat, g["bond_mat"])                          # bond indexing: bit-exact
    assert np.array_equal(st.bond_list(), g["bond_list"])                      # and in the reference's visiting order
    bv = np.array([st.dist_mat_vec[i, a, b] for i, a, b in st.bond_list()]).reshape(-1, 3)
    np.testing.assert_allclose(bv, g["bond_vec"], rtol=0, atol=4e-16 * max(1.0, np.abs(g["bond_vec"]).max()))
    assert np.array_equal(sy.get_soc_mat(), g["soc_mat"])
    assert len(sy.all_iter) == int(g["n_orbitals"])
    # on-site diagonal against the reference's H_wo_g central-image diagonal
    H = np.zeros(tuple(g["H_wo_g_shape"]), dtype=complex)
    H[tuple(g["H_wo_g_idx"].T)] = g["H_wo_g_val"]
    onsite = np.concatenate([np.diag(sy.get_onsite_term(a)) for a in range(len(st.atoms))])
    assert np.array_equal(onsite, np.diag(H[int(st.max_image / 2)]).real)


@pytest.mark.parametrize("name", ["C1_cubic_sp3", "C2_graphene_pz", "C3_sb_buckled", "sp_chain", "C5_zigzag_8", "T_ce_spdf_1atom"])
def test_get_kpts_bit_exact(name):
    g = golden(name)
    _, sy = build_structure(spec_for(name))
    i = 0
    while f"path{i}_in" in g:
        pts, lens, spl = sy.get_kpts(g[f"path{i}_in"].tolist(), int(g[f"path{i}_nk"]))
        assert isinstance(pts[0], list)                                        # first point stays the caller's list
        assert np.array_equal(np.array([np.asarray(p, float) for p in pts]), g[f"path{i}_k"])
        assert np.array_equal(lens, g[f"path{i}_len"])
        assert np.array_equal(spl, g[f"path{i}_spl"])
        i += 1


def test_scaling_laws_match_reference_hoppings():
    """lowsym_spd_laws: V(d) through the product's ScalingLaw classes == the hoppings the reference produced."""
    from oracle.sk_oracle import sk_table
    g = golden("lowsym_spd_laws")
    st, sy = build_structure(spec_for("lowsym_spd_laws"))
    H = np.zeros(tuple(g["H_wo_g_shape"]), dtype=complex)
    H[tuple(g["H_wo_g_idx"].T)] = g["H_wo_g_val"]
    starts = np.cumsum([0] + [len(a.orbitals) for a in st.atoms])
    for im, i, j in st.bond_list():
        V = sy.get_hop_params(i, j, im)
        l, m, n = st.get_dir_cos(im, i, j)
        T = sk_table(V, l, m, n)
        ki = [tb.Atom.ORBITALS_ALL.index(o) for o in st.atoms[i].orbitals]
        kj = [tb.Atom.ORBITALS_ALL.index(o) for o in st.atoms[j].orbitals]
        np.testing.assert_allclose(T[np.ix_(ki, kj)], H[im, starts[i]:starts[i + 1], starts[j]:starts[j + 1]].real, rtol=0, atol=1e-14)


def test_polynomial_tabulated_custom_laws_bit_exact():
    """scaling.py:312-447 through the product's classes == the reference's own `__call__` values (laws_misc.npz, made by
    oracle/make_golden.py:laws_fixture), and the oracle's scalar restatement of the same laws."""
    from oracle.make_golden_cases import LAW_CASES
    from oracle import laws_oracle
    g = golden("laws_misc")
    for tag, (kind, kw) in LAW_CASES.items():
        lawobj = getattr(tb.scaling, kind)(**kw)
        got = np.array([lawobj(float(x)) for x in g["d"]])
        assert np.array_equal(got, g[tag]), (tag, np.abs(got - g[tag]).max())
        assert isinstance(lawobj(float(g["d"][0])), float)
        np.testing.assert_array_equal(lawobj(g["d"]), g[tag])                       # array calls agree with scalar calls
        if kind != "Constant":
            okw = {k: v for k, v in kw.items()}
            orc = laws_oracle.make(kind, **okw)
            np.testing.assert_allclose([orc(float(x)) for x in g["d"]], g[tag], rtol=1e-15, atol=0)
    assert repr(tb.scaling.ensure_scaling(-0.37)) == "Constant(-0.37)"


def test_error_conventions():
    with pytest.raises(AssertionError, match="wrong orbitals"):
        tb.Atom("X", [0, 0, 0], ["s", "pq"])
    spec = W.cubic_sp3()
    lat = tb.Lattice(spec["lattice"], spec["a"])
    st = tb.Structure(lat, [tb.Atom("X", [0, 0, 0], ["s"])], bond_cut=spec["bond_cut"])
    with pytest.raises(AssertionError, match="wrong parameter set"):
        tb.System(st, {"X": ["s"]}, {"X": {"e_s": 0}})
    with pytest.raises(TypeError):                       # missing bond_cut pair -> None["NN"] in the reference
        tb.Structure(lat, [tb.Atom("X", [0, 0, 0], ["s"])], bond_cut={"YY": {"NN": 1.0}})


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sktb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sktb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = _lib.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/sktb.h but not exported by libsktb.so"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.v_names() == list(__import__("oracle.sk_oracle", fromlist=["V_NAMES"]).V_NAMES)


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    spec = W.cubic_sp3()
    with pytest.raises(_lib.SktbError, match="no CPU fallback"):
        W.instantiate(spec, tb, tb.scaling)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pysktb_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "/root/reference" not in src, f


def test_greens_function_method_selection_without_a_device():
    """Host logic of GreensFunction(ham, method=): "auto" takes the dense inverse exactly when the plan's Hermiticity defect of the
    H(k) that get_ham(k, l_soc=soc) returns exceeds 1e-12 eV (soc=False ignores a non-Hermitian soc_mat); explicit methods override;
    unknown methods raise.  No device is touched: the Hamiltonian is a stub carrying the two defects."""
    import types
    import pytest as _pt
    from pysktb_b200.greens import GreensFunction, SurfaceGreensFunction, HERM_DEFECT_TOL

    def stub(defect, defect_nosoc):
        st = types.SimpleNamespace(atoms=[types.SimpleNamespace(orbitals=["s", "px"]), types.SimpleNamespace(orbitals=["s"])])
        return types.SimpleNamespace(n_orbitals=3, structure=st, system=None, herm_defect=defect, herm_defect_nosoc=defect_nosoc)

    assert HERM_DEFECT_TOL == 1e-12
    herm, fshell, dsoc = stub(3e-17, 3e-17), stub(0.82, 0.82), stub(1.7, 0.0)
    assert not GreensFunction(herm)._use_inverse(True) and not GreensFunction(herm)._use_inverse(False)
    assert GreensFunction(fshell)._use_inverse(True) and GreensFunction(fshell)._use_inverse(False)
    assert GreensFunction(dsoc)._use_inverse(True) and not GreensFunction(dsoc)._use_inverse(False)
    assert not GreensFunction(fshell, method="eigh")._use_inverse(True)
    assert GreensFunction(herm, method="inverse")._use_inverse(False)
    with _pt.raises(ValueError):
        GreensFunction(herm, method="lu")
    sg = SurfaceGreensFunction(fshell, surface_atoms=[1], method="eigh")
    assert sg._gf.method == "eigh" and sg.surface_orbital_indices == [2] and sg._rows(True) == [2, 5]
