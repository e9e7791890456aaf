"""ORACLE (test infrastructure only).

Plain numpy/CPython restatement of pysktb's k-point path, used ONLY by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs as the checker.  The product (pysktb_b200) never
imports this package.

Every function cites the reference lines it restates.  Loops are kept literal (one np.dot per pair,
one exp per bonded pair, LAPACK zheevd through numpy, dense inverse for Green's functions) so that the
floating-point sequence is the reference's; only the redundant per-orbital recomputation of identical
phases is hoisted to per-atom-pair (same value, same operation, computed once).

Pinning: tests/golden/*.npz hold outputs of the real reference (made by oracle/make_golden.py in the
build container, where /root/reference exists); tests/test_oracle_golden.py checks this module against
them.  The eigen-decomposition and the inverse are third-party LAPACK (numpy / scipy-openblas, not part
of the reference tree; reference pins only numpy>=1.18.3, scipy>=1.2.1) - the oracle calls the same
numpy/scipy entry points the reference calls (hamiltonian.py:119,123; greens.py:70).
"""
import itertools

import numpy as np
from scipy import linalg as sla

from .sk_oracle import sk_table, V_NAMES

ORBITALS_ALL = [
    "s", "px", "py", "pz", "dxy", "dyz", "dxz", "dx2-y2", "dz2",
    "fz3", "fxz2", "fyz2", "fz(x2-y2)", "fxyz", "fx(x2-3y2)", "fy(3x2-y2)", "S",
]
P_ORBS, D_ORBS, F_ORBS = ORBITALS_ALL[1:4], ORBITALS_ALL[4:9], ORBITALS_ALL[9:16]


# ------------------------------------------------------------------------------------------------
# SOC blocks (system.py:651-658, 690-702, 741-827), units of lambda/2
# ------------------------------------------------------------------------------------------------
def _soc_p():
    j = 1j
    return np.array(
        [[0, -j, 0, 0, 0, 1], [j, 0, 0, 0, 0, j], [0, 0, 0, -1, -j, 0],
         [0, 0, -1, 0, j, 0], [0, 0, j, -j, 0, 0], [1, -j, 0, 0, 0, 0]], dtype=complex)


def _soc_d():
    j, s = 1j, np.sqrt(3)
    return np.array(
        [[0, j, -1, 0, 0, 0, 0, -j, -2, 0],
         [-j, 0, 0, -j, j * s, 0, 0, 0, -1, -s],
         [1, 0, 0, 1, s, j, 0, 0, -j, -j * s],
         [0, j, -1, 0, 0, 2, 1, j, 0, 0],
         [0, -j * s, -s, 0, 0, 0, s, j * s, 0, 0],
         [0, 0, -j, 2, 0, 0, -j, 1, 0, 0],
         [0, 0, 0, 1, s, j, 0, 0, j, -j * s],
         [j, 0, 0, -j, -j * s, -1, 0, 0, 1, -s],
         [-2, -1, j, 0, 0, 0, -j, 1, 0, 0],
         [0, -s, j * s, 0, 0, 0, j * s, -s, 0, 0]], dtype=complex)


def _soc_f():
    j = 1j
    s6, s10 = np.sqrt(6), np.sqrt(10)
    M = np.zeros((14, 14), dtype=complex)
    for k, (p, q) in enumerate([(1, 2), (3, 4), (5, 6)], start=1):     # Lz Sz, up block
        M[p, q], M[q, p] = -k * j, k * j
        M[p + 7, q + 7], M[q + 7, p + 7] = k * j, -k * j               # down block
    # ladder terms between spin blocks: (row pair, col pair, magnitude)
    for (r0, r1), (c0, c1), mag in [((1, 2), (10, 11), s10), ((3, 4), (12, 13), s6),
                                    ((8, 9), (3, 4), s10), ((10, 11), (5, 6), s6)]:
        M[r0, c0], M[r0, c1] = mag, (j if r0 < 7 else -j) * mag
        M[r1, c0], M[r1, c1] = (-j if r0 < 7 else j) * mag, mag
        M[c0, r0], M[c1, r0] = mag, (-j if r0 < 7 else j) * mag
        M[c0, r1], M[c1, r1] = (j if r0 < 7 else -j) * mag, mag
    M[0, 8], M[0, 9], M[8, 0], M[9, 0] = s6, j * s6, s6, -j * s6
    M[7, 1], M[7, 2], M[1, 7], M[2, 7] = s6, -j * s6, s6, j * s6
    return M


class OracleModel:
    """Structure + System + Hamiltonian of the reference, literal loops.

    atoms: list of (element, frac_pos, orbitals); orbitals are per ELEMENT - the last atom of an element
    wins and is then applied to every atom of that element (hamiltonian.py:68, system.py:105-108).
    """

    def __init__(self, lattice_matrix, lat_const, atoms, bond_cut, params, periodicity=None):
        self.lat = np.array(lattice_matrix) * lat_const                      # lattice.py:13
        self.rec = np.linalg.inv(self.lat).T                                 # lattice.py:76-81
        per_elem = {el: orbs for el, _, orbs in atoms}
        self.elements = [el for el, _, _ in atoms]
        self.pos = [np.array(p) for _, p, _ in atoms]
        self.orbitals = [per_elem[el] for el in self.elements]
        for orbs in self.orbitals:
            assert set(orbs).issubset(set(ORBITALS_ALL)), "wrong orbitals"   # atom.py:74
        self.periodicity = periodicity or [True, True, True]                 # structure.py:19
        self.max_image = 3 ** int(np.sum(self.periodicity))
        self.bond_cut = bond_cut
        self.params = params
        self.n_atom = len(atoms)
        self._geometry()
        self._check_params()
        # system.py:117-125
        self.all_iter = [(a, oi, self.elements[a], o) for a in range(self.n_atom)
                         for oi, o in enumerate(self.orbitals[a])]
        self.n_orbitals = len(self.all_iter)
        self._hoppings()
        self.soc_mat = sla.block_diag(*[self._soc_atom(a) for a in range(self.n_atom)])  # system.py:846-854

    # -- structure.py:58-143 ------------------------------------------------------------------
    def _geometry(self):
        axes = [np.arange(3) - 1 if p else [0] for p in self.periodicity]
        self.images = list(itertools.product(*axes))
        n, mi = self.n_atom, self.max_image
        lat_T = self.lat.T
        self.dist_mat_vec = np.zeros((mi, n, n, 3))
        for ii, image in enumerate(self.images):
            for i in range(n):
                for j in range(n):
                    diff = np.array(self.pos[i] + image) - np.array(self.pos[j])
                    if np.linalg.norm(diff) == 0:
                        continue
                    self.dist_mat_vec[ii, i, j, :] = np.dot(lat_T, diff)
        self.dist_mat = np.linalg.norm(self.dist_mat_vec, axis=-1)
        keys = list(self.bond_cut.keys())
        within = np.zeros((mi, n, n), dtype=bool)
        for i in range(n):
            for j in range(n):
                e1, e2 = self.elements[i], self.elements[j]
                pair = f"{e1}{e2}" if f"{e1}{e2}" in keys else (f"{e2}{e1}" if f"{e2}{e1}" in keys else None)
                cutoff = (self.bond_cut[pair] if pair is not None else None)["NN"]   # TypeError if missing
                within[:, i, j] = self.dist_mat[:, i, j] < cutoff
        self.bond_mat = within * (self.dist_mat > 0)

    def _uniq_elements(self):
        seen = []
        for e in self.elements:
            if e not in seen:
                seen.append(e)
        return seen

    def _check_params(self):                                                 # system.py:44-46, 127-134
        el = self._uniq_elements()
        self.param_keys = el + ["".join(p) for p in itertools.combinations_with_replacement(el, r=2)]
        assert set(self.param_keys).issubset(set(self.params.keys())), "wrong parameter set"

    def _pair_key(self, e1, e2):                                             # system.py:158-165
        if f"{e1}{e2}" in self.param_keys:
            return f"{e1}{e2}"
        if f"{e2}{e1}" in self.param_keys:
            return f"{e2}{e1}"
        return None

    # -- hamiltonian.py:425-488 ---------------------------------------------------------------
    def _hoppings(self):
        n, mi = self.n_orbitals, self.max_image
        self.H_wo_g = np.zeros((mi, n, n), dtype=complex)
        start = np.cumsum([0] + [len(o) for o in self.orbitals])
        self.bonds = []
        for a1 in range(self.n_atom):
            for a2 in range(self.n_atom):
                for im in range(mi):
                    if im == mi / 2 and a1 == a2:
                        continue
                    if not self.bond_mat[im, a1, a2]:
                        continue
                    self.bonds.append((im, a1, a2))
                    raw = self.params[self._pair_key(self.elements[a1], self.elements[a2])]
                    d = self.dist_mat[im, a1, a2]
                    V = {k: (v(d) if callable(v) else v) for k, v in raw.items() if k.startswith("V_")}
                    vec = self.dist_mat_vec[im, a1, a2, :]
                    l, m, nn = vec / np.linalg.norm(vec)                     # structure.py:168-174
                    table = sk_table(V, l, m, nn)
                    for o1i, o1 in enumerate(self.orbitals[a1]):
                        for o2i, o2 in enumerate(self.orbitals[a2]):
                            self.H_wo_g[im, start[a1] + o1i, start[a2] + o2i] = \
                                table[ORBITALS_ALL.index(o1)][ORBITALS_ALL.index(o2)]
        for a in range(self.n_atom):
            blk = slice(start[a], start[a + 1])
            self.H_wo_g[int(mi / 2), blk, blk] = self._onsite(a)

    def _onsite(self, a):                                                    # system.py:466-521
        par, have = self.params[self.elements[a]], self.orbitals[a]
        e = {}
        if "s" in have:
            e["s"] = par["e_s"]
        if set(P_ORBS) & set(have):
            ep = par["e_p"] if isinstance(par["e_p"], list) else [par["e_p"]] * 3
            e.update(dict(zip(P_ORBS, ep)))
        if set(D_ORBS) & set(have):
            e.update({o: par.get("e_d", 0) for o in D_ORBS})
        if set(F_ORBS) & set(have):
            e.update({o: par.get("e_f", 0) for o in F_ORBS})
        if "S" in have:
            e["S"] = par.get("e_S", 0)
        return np.diag([e[o] for o in ORBITALS_ALL if o in have])

    def _soc_atom(self, a):                                                  # system.py:596-844
        par, orbs = self.params[self.elements[a]], self.orbitals[a]
        n = len(orbs)
        h = np.zeros((2 * n, 2 * n), dtype=complex)

        def put(shell, blk):
            if all(o in orbs for o in shell):
                idx = [orbs.index(o) for o in shell]
                idx = idx + [i + n for i in idx]
                for r, ri in enumerate(idx):
                    for c, ci in enumerate(idx):
                        h[ri, ci] = blk[r, c]

        if "lambda" in par or "lambda_p" in par:
            lam = par.get("lambda_p", par.get("lambda", 0))
            if lam != 0:
                put(P_ORBS, lam * _soc_p() / 2.0)
        if "lambda_d" in par and par["lambda_d"] != 0:
            put(D_ORBS, par["lambda_d"] * _soc_d() / 2.0)
        if "lambda_f" in par and par["lambda_f"] != 0:
            put(F_ORBS, _soc_f() * (par["lambda_f"] / 2.0))
        return h

    # -- hot path: hamiltonian.py:275-317 (calc_g), :98-106 (get_ham), :108-126 (_sol_ham) -----------
    def calc_g(self, kpt):
        kc = np.dot(kpt, self.rec)
        n, mi = self.n_orbitals, self.max_image
        start = np.cumsum([0] + [len(o) for o in self.orbitals])
        g = np.zeros((mi, n, n), dtype=complex)
        for im, a1, a2 in zip(*np.nonzero(self.bond_mat)):
            phase = np.exp(2.0 * np.pi * 1j * np.dot(kc, self.dist_mat_vec[im, a1, a2, :]))
            g[im, start[a1]:start[a1 + 1], start[a2]:start[a2 + 1]] = phase
        g[int(mi / 2), :, :] += np.eye(n, dtype=complex)
        return g

    def get_ham(self, kpt, l_soc=True):
        h = np.sum(self.H_wo_g * self.calc_g(kpt), axis=0)
        if l_soc == True:  # noqa: E712  (the reference's comparison)
            h = np.kron(h, np.eye(2)) + self.soc_mat
        return h

    @staticmethod
    def sol_ham(ham, eig_vectors=False):
        if np.max(ham - ham.T.conj()) > 1.0e-9:                              # lexicographic complex max (Q5)
            raise Exception("\n\nHamiltonian matrix is not hermitian?!")
        if not eig_vectors:
            w = np.linalg.eigvalsh(ham)
            return np.sort(np.array(w.real, dtype=float))
        w, v = np.linalg.eigh(ham)
        order = np.array(w.real, dtype=float).argsort()
        return np.array(w.real, dtype=float)[order], v.T[order]

    def solve_kpath(self, k_list, eig_vectors=False, soc=True):
        """Serial loop of hamiltonian.py:191-209; returns [band,k] (and [band,k,comp])."""
        N = self.n_orbitals * (2 if soc else 1)
        vals = np.zeros((N, len(k_list)))
        vecs = np.zeros((N, len(k_list), N), dtype=complex) if eig_vectors else None
        for i, k in enumerate(k_list):
            out = self.sol_ham(self.get_ham(k, l_soc=soc), eig_vectors)
            if eig_vectors:
                vals[:, i], vecs[:, i, :] = out
            else:
                vals[:, i] = out
        return (vals, vecs) if eig_vectors else vals

    # -- system.py:51-103 ---------------------------------------------------------------------------
    def get_kpts(self, path, nk):
        pts = [path[0]]
        for a, b in zip(path[:-1], path[1:]):
            ki, kf = np.array(a), np.array(b)
            for s in range(nk):
                frac = (s + 1.0) / float(nk)
                pts.append(kf * frac + ki * (1.0 - frac))
        cart = [np.dot(self.rec, k) for k in pts]
        seg = [np.linalg.norm(cart[i] - cart[i - 1]) for i in range(len(cart))]
        seg[0] = 0
        lens = np.cumsum(seg)
        flat = np.array(pts).reshape(-1, 3)
        spl = [lens[np.all(flat == sp, axis=1)] for sp in path]
        return pts, lens, np.unique(np.concatenate(spl).ravel())

    # -- greens.py:46-167, 169-259, 431-523 ------------------------------------------------------------
    @staticmethod
    def kmesh(nk):
        return np.array([[i / max(nk[0], 1), j / max(nk[1], 1), l / max(nk[2], 1)]
                         for i in range(nk[0]) for j in range(nk[1]) for l in range(nk[2])])

    @staticmethod
    def kmesh_point(nk, idx):
        """Point `idx` of kmesh(nk) (same loop order and divisions as greens.py:104-109)."""
        l = idx % nk[2]
        j = (idx // nk[2]) % nk[1]
        i = idx // (nk[1] * nk[2])
        return np.array([i / max(nk[0], 1), j / max(nk[1], 1), l / max(nk[2], 1)])

    def spectral(self, E, k, eta=0.01, soc=True):
        H = self.get_ham(k, l_soc=soc)
        G = sla.inv((E + 1j * eta) * np.eye(H.shape[0]) - H)
        return -1.0 / np.pi * G.imag

    def dos(self, energies, nk, eta=0.01, soc=True, k_mesh=None):
        k_mesh = self.kmesh(nk) if k_mesh is None else k_mesh
        out = np.zeros(len(energies))
        for i, E in enumerate(energies):
            for k in k_mesh:
                out[i] += np.trace(self.spectral(E, k, eta, soc)).real
            out[i] /= len(k_mesh)
        return out

    def ldos(self, energies, atom_indices=None, orbital_indices=None, nk=(20, 20, 20), eta=0.01, soc=True):
        k_mesh = self.kmesh(nk)
        if atom_indices is None:
            atom_indices = list(range(self.n_atom))
        start = np.cumsum([0] + [len(o) for o in self.orbitals])
        out = np.zeros((len(atom_indices), len(energies)))
        for ei, E in enumerate(energies):
            for k in k_mesh:
                A = self.spectral(E, k, eta, soc)
                for i, a in enumerate(atom_indices):
                    idx = list(range(start[a], start[a + 1]))
                    if orbital_indices is not None:
                        idx = [idx[j] for j in orbital_indices if j < len(idx)]
                    for o in idx:
                        out[i, ei] += A[o, o].real
                        if soc:
                            out[i, ei] += A[o + self.n_orbitals, o + self.n_orbitals].real   # Q1 convention
            out[:, ei] /= len(k_mesh)
        return out

    def edge_dos(self, energies, surface_atoms, k_points=None, nk=50, eta=0.01, soc=True):
        if k_points is None:
            k_points = np.array([[k, 0, 0] for k in np.linspace(0, 1, nk)])   # endpoint included (Q10)
        start = np.cumsum([0] + [len(o) for o in self.orbitals])
        sel = [o for a in range(self.n_atom) if a in surface_atoms for o in range(start[a], start[a + 1])]
        out = np.zeros(len(energies))
        for ei, E in enumerate(energies):
            for k in k_points:
                A = self.spectral(E, k, eta, soc)
                n = A.shape[0]
                for o in sel:
                    out[ei] += A[o, o]
                    if soc and o + self.n_orbitals < n:
                        out[ei] += A[o + self.n_orbitals, o + self.n_orbitals]
            out[ei] /= len(k_points)
        return out


# ------------------------------------------------------------------------------------------------
# eigen-based closed forms (what the CUDA path computes) - for oracle-side cross-checks of the
# identity  -Im Tr G / pi = sum_n (eta/pi) / ((E - e_n)^2 + eta^2)  on Hermitian H  
# ------------------------------------------------------------------------------------------------
def lorentz_dos(evals, energies, eta):
    """evals [N, nk] -> dos [nE] (already divided by nk)."""
    E = np.asarray(energies)[:, None, None]
    return ((eta / np.pi) / ((E - evals[None]) ** 2 + eta * eta)).sum(axis=(1, 2)) / evals.shape[1]


# ---------------------------------------------------------------------------------------------------------
# Hamiltonian.get_dos (Gaussian) and the band energy of Forces.get_total_energy
# ---------------------------------------------------------------------------------------------------------
def gauss_dos(evals, energy, w):
    """hamiltonian.py:249-252: D(E) = sum_i exp(-(E - e_i)^2 / (2 w^2)) / (pi w sqrt(2))."""
    e = np.asarray(evals, dtype=float).reshape(-1)
    E = np.asarray(energy, dtype=float)
    return np.exp(-((E[..., None] - e) ** 2) / (2 * w ** 2)).sum(axis=-1) / (np.pi * w * np.sqrt(2))


def get_dos(model, energy, w=1e-2, nk=(20, 20, 20)):
    """hamiltonian.py:233-253 with eig=None: endpoint-included linspace mesh, kx outer / kz inner, soc=True."""
    kx, ky, kz = (np.linspace(0, 1, int(n)) for n in nk)
    ks = [[i, j, k] for i in kx for j in ky for k in kz]
    return gauss_dos(model.solve_kpath(ks, soc=True), energy, w)


def total_energy_band(model, n_electrons, nk=(10, 10, 10), soc=True):
    """forces.py:76-81, 381-395: mesh linspace(endpoint=False), sum of the lowest n_occ eigenvalues, weight 1/n_k,
    factor 2 without soc."""
    kx, ky, kz = (np.linspace(0, 1, int(n), endpoint=False) for n in nk)
    ks = [[x, y, z] for x in kx for y in ky for z in kz]
    ev = model.solve_kpath(ks, soc=soc)
    n_occ = n_electrons if soc else n_electrons // 2
    n_occ = min(n_occ, ev.shape[0])
    return (1.0 if soc else 2.0) * ev[:n_occ].sum() / len(ks)


def total_energy_mean(model, filled_band=0, nk=10, dim=3, soc=True):
    """hamiltonian.py:319-323 -> utils/energitics.py:26-63: mean of the lowest `neig` eigenvalues (neig = filled_band or
    int(N/2)) over the endpoint-included mesh linspace(0,1,nk)^dim.  (The reference obtains them with ARPACK eigsh
    'SR' on the full matrix; for Hermitian H(k) these are the lowest eigenvalues.)"""
    x = np.linspace(0, 1, int(nk))
    if dim == 3:
        ks = [[a, b, c] for a in x for b in x for c in x]
    elif dim == 2:
        ks = [[a, b, 0.0] for a in x for b in x]
    else:
        ks = [[a, 0.0, 0.0] for a in x]
    ev = model.solve_kpath(ks, soc=soc)
    neig = int(filled_band) if filled_band else int(ev.shape[0] / 2)
    return float(ev[:neig].mean())

