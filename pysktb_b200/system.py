"""Model bookkeeping: orbital enumeration, hopping parameters per bond, on-site terms, SOC blocks, k-paths.

Host-side (k-independent) setup, restating the parts of the reference `System` that feed the k-point
hot path (pysktb/system.py):

* `get_kpts / get_kpt_path / get_kpt_len`            system.py:51-103   (bit-exact k-paths)
* `get_all_iter`, `get_param_key`, `_get_pair`        system.py:117-134, 158-165
* `get_hop_params` (floats or ScalingLaw at |d|)      system.py:173-213
* `get_onsite_term`, simple branch only               system.py:466-521  (the crystal-field branch
  :523-594 is unreachable: `scale_params` is always None-valued, system.py:35-42)
* `_get_soc_mat_i / get_soc_mat`                      system.py:596-854

The SOC blocks are stored as token grids (units of lambda/2) and expanded at run time; they are the
reference's matrices verbatim in value, including the d block that is not Hermitian.
"""
import itertools

import numpy as np
from scipy import linalg

from .atom import Atom
from .scaling import ScalingLaw

_P_ORBS = Atom.ORBITAL_FAMILIES["p"]
_D_ORBS = Atom.ORBITAL_FAMILIES["d"]
_F_ORBS = Atom.ORBITAL_FAMILIES["f"]

# token -> complex value; 'r' = sqrt(3), 'u' = sqrt(6), 'w' = sqrt(10), leading 'i' = imaginary unit
_R3, _R6, _R10 = np.sqrt(3), np.sqrt(6), np.sqrt(10)


def _tok(t):
    sign = -1.0 if t.startswith("-") else 1.0
    t = t.lstrip("-")
    unit = 1j if t.startswith("i") else 1.0
    t = t[1:] if t.startswith("i") else t
    mag = {"": 1.0, "r": _R3, "u": _R6, "w": _R10}[t] if not t.isdigit() else float(t)
    return sign * unit * mag


def _grid(rows):
    return np.array([[_tok(t) for t in row.split()] for row in rows], dtype=complex)


# basis [x^, y^, z^, xv, yv, zv]  (^ = up, v = down);  system.py:651-658
_SOC_P = _grid([
    "0 -i 0 0 0 1",
    "i 0 0 0 0 i",
    "0 0 0 -1 -i 0",
    "0 0 -1 0 i 0",
    "0 0 i -i 0 0",
    "1 -i 0 0 0 0",
])

# basis [dxy dyz dxz dx2-y2 dz2]^ then the same v;  system.py:690-702
_SOC_D = _grid([
    "0 i -1 0 0 0 0 -i -2 0",
    "-i 0 0 -i ir 0 0 0 -1 -r",
    "1 0 0 1 r i 0 0 -i -ir",
    "0 i -1 0 0 2 1 i 0 0",
    "0 -ir -r 0 0 0 r ir 0 0",
    "0 0 -i 2 0 0 -i 1 0 0",
    "0 0 0 1 r i 0 0 i -ir",
    "i 0 0 -i -ir -1 0 0 1 -r",
    "-2 -1 i 0 0 0 -i 1 0 0",
    "0 -r ir 0 0 0 ir -r 0 0",
])

# sparse (row, col, token) list, 14x14, basis [7 f orbitals]^ then v;  system.py:753-824
_SOC_F_ENTRIES = """
1 2 -i|2 1 i|3 4 -i2|4 3 i2|5 6 -i3|6 5 i3
8 9 i|9 8 -i|10 11 i2|11 10 -i2|12 13 i3|13 12 -i3
0 8 u|0 9 iu|8 0 u|9 0 -iu
1 10 w|1 11 iw|2 10 -iw|2 11 w|10 1 w|11 1 -iw|10 2 iw|11 2 w
3 12 u|3 13 iu|4 12 -iu|4 13 u|12 3 u|13 3 -iu|12 4 iu|13 4 u
7 1 u|7 2 -iu|1 7 u|2 7 iu
8 3 w|8 4 -iw|9 3 iw|9 4 w|3 8 w|4 8 iw|3 9 -iw|4 9 w
10 5 u|10 6 -iu|11 5 iu|11 6 u|5 10 u|6 10 iu|5 11 -iu|6 11 u
"""


def _soc_f_matrix():
    m = np.zeros((14, 14), dtype=complex)
    for item in _SOC_F_ENTRIES.replace("\n", "|").split("|"):
        if item.strip():
            r, c, t = item.split()
            m[int(r), int(c)] = _tok(t)
    return m


_SOC_F = _soc_f_matrix()


class System(object):
    def __init__(self, structure, orbitals, parameters, scale_params=None):
        self.structure = structure
        self.orbitals = orbitals
        self.set_orbitals()
        self.all_orbitals = self.get_all_orbitals()
        self.all_iter = self.get_all_iter()
        self.params = parameters
        # the reference discards the argument and installs a None-valued table (system.py:35-42)
        self.scale_params = {
            "".join(k): None for k in itertools.product(list(self.orbitals.keys()), repeat=2)
        }
        need = self.get_param_key()
        assert set(need).issubset(set(self.params.keys())), (
            "wrong parameter set\n" + f"given: {list(self.params.keys())}\n"
        ) + f"required: {need}"

    # -- k paths (host; bit-exact restatement) ----------------------------------------------
    def get_kpts(self, sp_kpts, kpt_den):
        sp_kpts = [sp_kpts]
        kpt_path = self.get_kpt_path(sp_kpts, kpt_den)
        kpts_len = self.get_kpt_len(kpt_path, self.structure.lattice.get_matrix())
        k_all = [k for seg in kpt_path for k in seg]
        flat = np.array(k_all).reshape(-1, 3)
        hits = [kpts_len[np.all(flat == sp, axis=1)] for sp in sp_kpts[0]]
        return k_all, kpts_len, np.unique(np.concatenate(hits).ravel())

    def get_kpt_path(self, sp_kpts, kpt_den=30):
        out = []
        for chain in sp_kpts:
            pts = [chain[0]]  # the first element stays the caller's list (system.py:71)
            for a, b in zip(chain[:-1], chain[1:]):
                k_i, k_f = np.array(a), np.array(b)
                for s in range(kpt_den):
                    frac = (s + 1.0) / float(kpt_den)
                    pts.append(k_f * frac + k_i * (1.0 - frac))
            out.append(pts)
        return out

    def get_kpt_len(self, kpts_path, lat_mat):
        rec = np.linalg.inv(lat_mat).T
        lens = []
        for chain in kpts_path:
            cart = [np.dot(rec, k) for k in chain]  # NB rec @ k here, k @ rec in calc_g
            seg = [np.linalg.norm(cart[i] - cart[i - 1]) for i in range(len(cart))]
            seg[0] = 0
            lens.extend(seg)
        return np.cumsum(lens)

    # -- orbital enumeration -------------------------------------------------------------------
    def set_orbitals(self):
        for atom in self.structure.atoms:
            atom.set_orbitals(self.orbitals[atom.element])

    def get_all_orbitals(self):
        return [(a.element, o) for a in self.structure.atoms for o in a.orbitals]

    def get_all_iter(self):
        return [
            (ai, oi, a.element, o)
            for ai, a in enumerate(self.structure.atoms)
            for oi, o in enumerate(a.orbitals)
        ]

    def get_param_key(self):
        elements = self.structure.get_elements()
        return list(elements) + [
            "".join(p) for p in itertools.combinations_with_replacement(elements, r=2)
        ]

    def _get_pair(self, ele_1, ele_2):
        keys = self.get_param_key()
        for cand in (f"{ele_1}{ele_2}", f"{ele_2}{ele_1}"):
            if cand in keys:
                return cand
        return None

    # -- hoppings --------------------------------------------------------------------------------
    def get_hop_params(self, atom_1_i, atom_2_i, image_i):
        atoms = self.structure.atoms
        pair = self._get_pair(atoms[atom_1_i].element, atoms[atom_2_i].element)
        hop = {k: v for k, v in self.params[pair].items() if k.startswith("V_")}
        if not any(isinstance(v, ScalingLaw) for v in hop.values()):
            return hop
        d = self.structure.dist_mat[image_i, atom_1_i, atom_2_i]
        return {k: (v(d) if isinstance(v, ScalingLaw) else v) for k, v in hop.items()}

    # -- on-site ---------------------------------------------------------------------------------
    def get_onsite_term(self, atom_i):
        atom = self.structure.atoms[atom_i]
        par = self.params[atom.element]
        have = set(atom.orbitals)
        level = {}
        if "s" in have:
            level["s"] = par["e_s"]
        if have & set(_P_ORBS):
            e_p = par["e_p"] if isinstance(par["e_p"], list) else [par["e_p"]] * 3
            level.update(zip(_P_ORBS, e_p))
        if have & set(_D_ORBS):
            level.update({o: par.get("e_d", 0) for o in _D_ORBS})
        if have & set(_F_ORBS):
            level.update({o: par.get("e_f", 0) for o in _F_ORBS})
        if "S" in have:
            level["S"] = par.get("e_S", 0)
        # canonical order, NOT the atom's own orbital order (system.py:483-521)
        return np.diag([level[o] for o in Atom.ORBITALS_ALL if o in have])

    # -- spin-orbit --------------------------------------------------------------------------------
    def _get_soc_mat_i(self, atom_i):
        atom = self.structure.atoms[atom_i]
        par = self.params[atom.element]
        orbs = atom.orbitals
        n = len(orbs)
        h = np.zeros((2 * n, 2 * n), dtype=complex)

        def place(shell, block):
            if not any(o in orbs for o in shell) or not all(o in orbs for o in shell):
                return
            up = [orbs.index(o) for o in shell]
            sel = up + [u + n for u in up]  # per-atom [all up | all down] layout
            h[np.ix_(sel, sel)] = block

        if "lambda" in par or "lambda_p" in par:
            lam = par.get("lambda_p", par.get("lambda", 0))
            if lam != 0:
                place(_P_ORBS, lam * _SOC_P / 2.0)
        if "lambda_d" in par and par["lambda_d"] != 0:
            place(_D_ORBS, par["lambda_d"] * _SOC_D / 2.0)
        if "lambda_f" in par and par["lambda_f"] != 0:
            place(_F_ORBS, _SOC_F * (par["lambda_f"] / 2.0))
        return h

    def get_soc_mat(self):
        return linalg.block_diag(*[self._get_soc_mat_i(i) for i in range(len(self.structure.atoms))])
