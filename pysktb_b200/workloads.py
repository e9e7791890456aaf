"""Named tight-binding models used by tests, golden-vector generation and bench.py.

Each builder returns a neutral *spec* dict (plain lists/floats) so that the same model can be
instantiated with this package, with the reference (oracle/make_golden.py) or with the numpy oracle.
Sources: BASELINE.json `configs` as made concrete in SURVEY.md §8(d); constructor arguments follow the
reference's README / example scripts cited per builder.
"""
import numpy as np

F_ORBS = ["fz3", "fxz2", "fyz2", "fz(x2-y2)", "fxyz", "fx(x2-3y2)", "fy(3x2-y2)"]
SPDF = ["s", "px", "py", "pz", "dxy", "dyz", "dxz", "dx2-y2", "dz2"] + F_ORBS


def law(kind, **kw):
    """Placeholder for a distance law; resolved against a namespace's scaling classes by `resolve`."""
    return ("law", kind, kw)


def cubic_sp3():
    """C1 - Quick Start simple cubic sp3 (reference README.md:43-56). N=8 with SOC doubling."""
    return dict(
        name="C1_cubic_sp3", lattice=np.eye(3).tolist(), a=3.5,
        atoms=[("X", [0, 0, 0], ["s", "px", "py", "pz"])], periodicity=None,
        bond_cut={"XX": {"NN": 3.6}},
        params={"X": {"e_s": 0, "e_p": 1.5},
                "XX": {"V_sss": -0.5, "V_sps": 0.5, "V_pps": 0.8, "V_ppp": -0.2}},
        soc=True,
    )


def graphene_pz():
    """C2 - graphene pz (docs/source/examples/greens_dos_example.py:55-95). N=2, soc=False."""
    a = 2.46
    return dict(
        name="C2_graphene_pz", lattice=[[1, 0, 0], [0.5, np.sqrt(3) / 2, 0], [0, 0, 10]], a=a,
        atoms=[("C", [0, 0, 0], ["pz"]), ("C", [1 / 3, 1 / 3, 0], ["pz"])],
        periodicity=[True, True, False], bond_cut={"CC": {"NN": a / np.sqrt(3) * 1.1}},
        params={"C": {"e_p": 0.0}, "CC": {"V_ppp": -2.7}}, soc=False,
    )


def cubic_s():
    """One s orbital on a simple cubic lattice with the default SOC doubling (N = 2, soc_mat = 0): the second
    case of the closed-form N = 2 kernel.  Not a reference example."""
    return dict(name="cubic_s", lattice=np.eye(3).tolist(), a=3.0, atoms=[("X", [0, 0, 0], ["s"])], periodicity=None,
                bond_cut={"XX": {"NN": 3.1}}, params={"X": {"e_s": 0.3}, "XX": {"V_sss": -0.7}}, soc=True)


def sb_buckled(theta_deg=16.0):
    """C3 - buckled group-V p-orbitals + SOC (docs/source/examples/examples.ipynb cell 33). N=12."""
    th = np.deg2rad(theta_deg)
    h = np.sin(th) * np.cos(th)
    d = 0.4
    return dict(
        name="C3_sb_buckled",
        lattice=[[1.0, 0.0, 0], [0.5, np.sqrt(3.0) / 2.0, 0], [0, 0, 10 / np.cos(th)]], a=np.cos(th),
        atoms=[("a", [1.0 / 3.0, 1.0 / 3.0, 2 + h / 10], ["px", "py", "pz"]),
               ("b", [2.0 / 3.0, 2.0 / 3.0, 2 - h / 10], ["px", "py", "pz"])],
        periodicity=None, bond_cut={"ab": {"NN": 1}, "aa": {"NN": 0}, "bb": {"NN": 0}},
        params={"a": {"e_p": [d, d, -d], "lambda": 0.19}, "b": {"e_p": [d, d, -d], "lambda": 0.2},
                "ab": {"V_ppp": -0.3, "V_pps": 1.3}, "bb": {"V_ppp": 0, "V_pps": 0},
                "aa": {"V_ppp": 0, "V_pps": 0.0}},
        soc=True,
    )


def _ce_params():
    # on-site / hoppings of docs/source/examples/f_orbital_example.py:76-112 plus the d-shell terms of
    # SURVEY.md §8(d) row C4; V_df* = 0 and no lambda_d so the reference's Hermiticity gate passes (Q5)
    return {
        "Ce": {"e_s": 2.0, "e_p": 0.5, "e_d": 1.0, "e_f": -1.0, "lambda": 0.05, "lambda_f": 0.25},
        "CeCe": {"V_sss": -0.8, "V_sps": 0.6, "V_pps": 0.8, "V_ppp": -0.2, "V_sfs": 0.15, "V_pfs": 0.2,
                 "V_pfp": -0.05, "V_ffs": -0.1, "V_ffp": 0.02, "V_ffd": -0.01, "V_fff": 0.005,
                 "V_sds": -0.3, "V_pds": -0.4, "V_pdp": 0.15, "V_dds": -0.5, "V_ddp": 0.3, "V_ddd": -0.05,
                 "V_dfs": 0.0, "V_dfp": 0.0, "V_dfd": 0.0},
    }


def ce_spdf_1atom():
    """T - north-star target: one Ce-like atom, s+p+d+f with SOC, simple cubic, 6 bonds. N=32."""
    a = 5.0
    return dict(name="T_ce_spdf_1atom", lattice=np.eye(3).tolist(), a=a,
                atoms=[("Ce", [0, 0, 0], list(SPDF))], periodicity=[True, True, True],
                bond_cut={"CeCe": {"NN": a * 1.1}}, params=_ce_params(), soc=True)


def ce_partial(n_orb=14):
    """T with only the first n_orb orbitals of the canonical list (N = 2 n_orb = 26, 28, 30): exercises the
    zero-padded rows/columns of the 32-wide register kernels.  Not a reference example."""
    spec = ce_spdf_1atom()
    spec["name"] = f"ce_partial_{n_orb}"
    spec["atoms"] = [("Ce", [0, 0, 0], list(SPDF[:n_orb]))]
    return spec


def ce_spdf_2atom():
    """C4 - two Ce-like atoms (bcc-like), 16 bonds. N=64."""
    a = 5.0
    return dict(name="C4_ce_spdf_2atom", lattice=np.eye(3).tolist(), a=a,
                atoms=[("Ce", [0, 0, 0], list(SPDF)), ("Ce", [0.5, 0.5, 0.5], list(SPDF))],
                periodicity=[True, True, True], bond_cut={"CeCe": {"NN": a * 0.9}},
                params=_ce_params(), soc=True)


def ce_spf_example():
    """The reference's own f-orbital example (s+p+f, 11 orbitals; f_orbital_example.py:25-115). N=22."""
    a = 5.0
    p = _ce_params()
    keep = ("V_sss", "V_sps", "V_pps", "V_ppp", "V_sfs", "V_pfs", "V_pfp", "V_ffs", "V_ffp", "V_ffd", "V_fff")
    p["CeCe"] = {k: p["CeCe"][k] for k in keep}
    p["Ce"] = {k: v for k, v in p["Ce"].items() if k != "e_d"}
    return dict(name="ce_spf_example", lattice=np.eye(3).tolist(), a=a,
                atoms=[("Ce", [0, 0, 0], ["s", "px", "py", "pz"] + F_ORBS)], periodicity=[True, True, True],
                bond_cut={"CeCe": {"NN": a * 1.1}}, params=p, soc=True)


def zigzag_ribbon(n_chains=256):
    """C5 - graphene zigzag ribbon (docs/source/examples/edge_states_example.py:50-135). N=2*n_chains."""
    a = 2.46
    d_cc = a / np.sqrt(3)
    vacuum = 10.0
    height = n_chains * a * np.sqrt(3) / 2 + vacuum
    A0, B0 = np.array([0.0, 0.0]), np.array([a / 2, a * np.sqrt(3) / 6])
    shift = np.array([a / 2, a * np.sqrt(3) / 2])
    atoms, edge = [], []
    for i in range(n_chains):
        for base in (A0, B0):
            c = base + i * shift
            c[0] = c[0] % a
            atoms.append(("C", [c[0] / a, c[1] / height, 0], ["pz"]))
        if i in (0, n_chains - 1):
            edge += [len(atoms) - 2, len(atoms) - 1]
    return dict(name=f"C5_zigzag_{n_chains}", lattice=[[1, 0, 0], [0, height / a, 0], [0, 0, vacuum / a]], a=a,
                atoms=atoms, periodicity=[True, False, False], bond_cut={"CC": {"NN": d_cc * 1.1}},
                params={"C": {"e_p": 0.0}, "CC": {"V_ppp": -2.7}}, soc=False, edge_atoms=edge)


def sp_chain():
    """1-D sp chain of the reference notebook (examples.ipynb cell 2). N=4 with SOC doubling."""
    return dict(name="sp_chain", lattice=[[1, 0, 0], [0, 10, 0], [0, 0, 10]], a=1,
                atoms=[("a", [0, 0, 0], ["s", "px"])], periodicity=None, bond_cut={"aa": {"NN": 1.2}},
                params={"a": {"e_p": 0, "e_s": 0}, "aa": {"V_sss": -0.2, "V_sps": -0.05, "V_pps": 0.2}},
                soc=True)


def custom_law_func(d):
    """User function of the `Custom` law in `lowsym_spd(laws="misc")` (module level: the reference's, this package's and
    the oracle's law objects all wrap this same callable)."""
    return -0.45 * (2.4 / d) ** 2.5 * np.exp(-0.2 * (d - 2.4))


def lowsym_spd(seed=3, laws=True):
    """Low-symmetry triclinic 3-atom s/p/d model with two elements, distance laws and p-SOC.
    Not in the reference's examples: it exercises generic direction cosines, the ScalingLaw path
    (system.py:205-213) and unequal orbital counts per atom.  Hermitian (no f, no lambda_d).
    laws="misc": the remaining law classes of scaling.py:312-447 - Polynomial, Tabulated (cubic spline), Custom."""
    if laws == "misc":
        spec = lowsym_spd(seed, laws=False)
        AA = dict(spec["params"]["AA"])
        AA["V_sss"] = law("Polynomial", coeffs=[-0.6, 0.35, -0.08, 0.01], d0=2.4, cutoff=4.4)
        AA["V_pps"] = law("Tabulated", distances=[1.8, 2.2, 2.6, 3.0, 3.4, 3.8], values=[1.55, 1.08, 0.77, 0.55, 0.4, 0.29])
        AA["V_pds"] = law("Custom", func=custom_law_func, d0=2.4, cutoff=3.3, smooth_width=0.7)
        AA["V_dds"] = law("Polynomial", coeffs=[-0.5], d0=2.4)
        spec["params"] = dict(spec["params"], AA=AA)
        spec["name"] = "lowsym_spd_misc"
        return spec
    r = np.random.default_rng(seed)
    lat = (np.eye(3) + 0.15 * r.standard_normal((3, 3))).tolist()
    sd = ["s", "px", "py", "pz", "dxy", "dyz", "dxz", "dx2-y2", "dz2"]
    atoms = [("A", [0.05, 0.1, 0.0], sd), ("B", [0.52, 0.47, 0.55], ["s", "px", "py", "pz"]),
             ("A", [0.3, 0.8, 0.25], sd)]
    def v(x, d0=2.4):
        return law("Harrison", V0=x, d0=d0, cutoff=4.4) if laws else x
    AA = {"V_sss": v(-0.6), "V_sps": v(0.7), "V_pps": law("Exponential", V0=0.9, d0=2.4, alpha=1.3) if laws else 0.9,
          "V_ppp": -0.25, "V_sds": v(-0.35), "V_pds": law("PowerLaw", V0=-0.45, d0=2.4, eta=3.5) if laws else -0.45,
          "V_pdp": 0.2, "V_dds": law("GSP", V0=-0.5, d0=2.4, n=2.0, nc=6.5, dc=3.6) if laws else -0.5,
          "V_ddp": 0.28, "V_ddd": -0.06}
    AB = {"V_sss": -0.5, "V_sps": 0.55, "V_pps": 0.75, "V_ppp": -0.2, "V_sds": -0.3, "V_pds": -0.4, "V_pdp": 0.18}
    BB = {"V_sss": -0.4, "V_sps": 0.5, "V_pps": 0.7, "V_ppp": -0.15}
    return dict(name="lowsym_spd" + ("_laws" if laws else ""), lattice=lat, a=3.1, atoms=atoms, periodicity=None,
                bond_cut={"AA": {"NN": 3.4}, "AB": {"NN": 2.9}, "BB": {"NN": 3.3}},
                params={"A": {"e_s": -1.0, "e_p": [0.4, 0.5, 0.6], "e_d": 1.2, "lambda": 0.08},
                        "B": {"e_s": 0.3, "e_p": 1.1, "lambda_p": 0.12}, "AA": AA, "AB": AB, "BB": BB},
                soc=True)


def d_soc_nonhermitian():
    """lambda_d != 0: the reference's d-SOC block is not Hermitian and `_sol_ham` raises."""
    return dict(name="d_soc_nonherm", lattice=np.eye(3).tolist(), a=3.0,
                atoms=[("M", [0, 0, 0], ["dxy", "dyz", "dxz", "dx2-y2", "dz2"])], periodicity=None,
                bond_cut={"MM": {"NN": 3.1}},
                params={"M": {"e_d": 0.0, "lambda_d": 1.0}, "MM": {"V_dds": -0.4, "V_ddp": 0.2, "V_ddd": -0.03}},
                soc=True)


def random_spec(seed=0):
    """Seeded random model for parity fuzzing (not a reference example): triclinic cell, 1-3 atoms of two elements,
    random orbital subsets (whole or partial shells, optionally the excited S orbital), random on-site terms, SOC
    strengths, hoppings (some through distance laws), periodicity and cut-offs.  f shells come without d-f hoppings
    and no model has lambda_d, so the reference's Hermiticity gate passes."""
    r = np.random.default_rng(1000 + seed)
    a = float(r.uniform(2.2, 3.6))
    lat = (np.eye(3) + 0.12 * r.standard_normal((3, 3))).tolist()
    P, D, F = ["px", "py", "pz"], ["dxy", "dyz", "dxz", "dx2-y2", "dz2"], list(SPDF[9:16])

    def orbitals():
        o = []
        if r.random() < 0.7: o += ["s"]
        if r.random() < 0.75: o += P if r.random() < 0.7 else [P[i] for i in sorted(r.choice(3, size=int(r.integers(1, 3)), replace=False))]
        if r.random() < 0.45: o += D if r.random() < 0.6 else [D[i] for i in sorted(r.choice(5, size=int(r.integers(1, 5)), replace=False))]
        if r.random() < 0.25: o += F if r.random() < 0.6 else [F[i] for i in sorted(r.choice(7, size=int(r.integers(1, 7)), replace=False))]
        if r.random() < 0.2: o += ["S"]
        return o or ["s"]
    orbs = {"A": orbitals(), "B": orbitals()}
    n_atoms = int(r.integers(1, 4))
    els = ["A"] + [("A", "B")[int(r.integers(0, 2))] for _ in range(n_atoms - 1)]
    atoms = [(el, r.random(3).round(3).tolist(), orbs[el]) for el in els]
    periodicity = [None, [True, True, True], [True, True, False], [True, False, False], [False, True, True]][int(r.integers(0, 5))]

    def onsite():
        t = {"e_s": float(r.normal()), "e_p": ([float(x) for x in r.normal(size=3)] if r.random() < 0.4 else float(r.normal())),
             "e_d": float(r.normal()), "e_f": float(r.normal()), "e_S": float(r.normal() + 3)}
        if r.random() < 0.7: t["lambda" if r.random() < 0.5 else "lambda_p"] = float(r.uniform(0.02, 0.3))
        if r.random() < 0.5: t["lambda_f"] = float(r.uniform(0.02, 0.3))
        return t

    def hop(d0):
        t = {}
        for n in ("V_sss", "V_sps", "V_pps", "V_ppp", "V_sds", "V_pds", "V_pdp", "V_dds", "V_ddp", "V_ddd", "V_sfs", "V_pfs",
                  "V_pfp", "V_ffs", "V_ffp", "V_ffd", "V_fff", "V_SSs", "V_sSs", "V_Sps", "V_Sds", "V_Sfs"):
            x = float(r.normal() * 0.5)
            u = r.random()
            if u < 0.15: t[n] = law("Harrison", V0=x, d0=d0, cutoff=3.0 * d0)
            elif u < 0.25: t[n] = law("Exponential", V0=x, d0=d0, alpha=float(r.uniform(0.8, 2.0)))
            elif u < 0.32: t[n] = law("PowerLaw", V0=x, d0=d0, eta=float(r.uniform(1.5, 4.0)))
            else: t[n] = x
        for n in ("V_dfs", "V_dfp", "V_dfd"): t[n] = 0.0
        return t
    nn = float(a * r.uniform(0.75, 1.15))
    params = {"A": onsite(), "B": onsite(), "AA": hop(0.8 * a), "AB": hop(0.7 * a), "BB": hop(0.8 * a)}
    cut = {k: {"NN": nn * float(r.uniform(0.9, 1.1))} for k in ("AA", "AB", "BB")}
    return dict(name=f"random_{seed}", lattice=lat, a=a, atoms=atoms, periodicity=periodicity, bond_cut=cut, params=params,
                soc=bool(r.random() < 0.7))


ALL = {f.__name__: f for f in (cubic_sp3, cubic_s, graphene_pz, sb_buckled, ce_spdf_1atom, ce_partial, ce_spdf_2atom, ce_spf_example,
                               zigzag_ribbon, sp_chain, lowsym_spd, d_soc_nonhermitian, random_spec)}


def resolve(params, scaling_ns):
    """Replace ('law', kind, kw) placeholders by instances of `scaling_ns.<kind>`."""
    out = {}
    for key, table in params.items():
        out[key] = {k: (getattr(scaling_ns, v[1])(**v[2]) if isinstance(v, tuple) and v and v[0] == "law" else v)
                    for k, v in table.items()}
    return out


def instantiate(spec, ns, scaling_ns=None, **ham_kw):
    """Build (structure, hamiltonian) from a spec with classes taken from namespace `ns`
    (this package, or the reference's modules)."""
    lat = ns.Lattice(spec["lattice"], spec["a"])
    atoms = [ns.Atom(el, list(pos), list(orbs)) for el, pos, orbs in spec["atoms"]]
    st = ns.Structure(lat, atoms, periodicity=spec["periodicity"], bond_cut=spec["bond_cut"])
    params = resolve(spec["params"], scaling_ns) if scaling_ns is not None else spec["params"]
    return st, ns.Hamiltonian(st, params, **ham_kw)
