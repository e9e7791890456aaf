"""Crystal structure: periodic images, displacement vectors, bond matrix (host-side setup).

Behavioural restatement of the reference `Structure` (pysktb/structure.py:10-25, 58-143, 168-182)
with the O(max_image * n_atom^2) interpreted loops replaced by array operations.  What is kept exactly:

* image order = itertools.product over axes of [-1,0,1] (periodic) or [0]      (structure.py:78-85)
* displacement  d[I,i,j] = lat.T @ ((pos_i + I) - pos_j), 0 when the fractional
  difference is exactly zero                                                      (structure.py:112-123,138-142)
* bond[I,i,j] = (|d| < cutoff["NN"]) & (|d| > 0), cutoff looked up as "AB" then "BA";
  a missing pair raises TypeError like `None["NN"]` does                          (structure.py:61-71,86-94)

Bond *indexing* must be bit-identical to the reference: any pair whose distance is
within a few ulp of the cutoff (or of zero) is re-evaluated with the reference's own per-pair
`np.dot(lat.T, diff)` so the strict comparisons decide on the same float.
"""
import itertools
from collections import OrderedDict

import numpy as np

from .atom import Atom
from .lattice import Lattice


class Structure(object):
    def __init__(self, lattice, atoms, periodicity=None, name=None, bond_cut=None, numba=True):
        assert isinstance(lattice, Lattice), "not Lattice object"
        assert isinstance(atoms, list), "atoms is not list"
        assert isinstance(atoms[0], Atom), "atom is not Atom object"
        self.numba = numba
        self.name = name or "system"
        self.lattice = lattice
        self.atoms = atoms
        self.bond_cut = bond_cut
        self.periodicity = periodicity or [True, True, True]
        self.max_image = 3 ** np.sum(self.periodicity)

        self.images = np.array(
            list(itertools.product(*[(-1, 0, 1) if p else (0,) for p in self.periodicity])), dtype=np.int64
        )
        self.dist_mat_vec = self.get_dist_matrix_vec()
        self.dist_mat = np.linalg.norm(self.dist_mat_vec, axis=-1)
        self.bond_mat = self.get_bond_mat()
        self.dir_cos = self.get_dir_cos_all()

    # -- geometry ---------------------------------------------------------------------------
    def _frac_diff(self):
        pos = np.array([np.asarray(a.pos, dtype=float) for a in self.atoms])
        img = self.images.astype(float)
        # (pos_i + I) first, then - pos_j : same rounding sequence as structure.py:141 + :116
        shifted = pos[None, :, :] + img[:, None, :]
        return shifted[:, :, None, :] - pos[None, None, :, :]

    def _pair_vec_exact(self, image_i, i, j):
        """Reference arithmetic for one (image, i, j): np.dot(lat.T, diff) (structure.py:116-123)."""
        diff = np.array(self.atoms[i].pos + self.images[image_i]) - np.array(self.atoms[j].pos)
        if np.linalg.norm(diff) == 0:
            return np.zeros(3)
        return np.dot(self.lattice.get_matrix().T, diff)

    def get_dist_matrix_vec(self):
        diff = self._frac_diff()
        lat = self.lattice.get_matrix()
        d = diff @ lat  # row form of lat.T @ diff
        d[np.all(diff == 0.0, axis=-1)] = 0.0
        return d

    def get_dist_matrix(self):
        return np.linalg.norm(self.get_dist_matrix_vec(), axis=-1)

    def _cutoff_matrix(self):
        keys = list(self.bond_cut.keys())
        elems = [a.element for a in self.atoms]
        uniq = list(OrderedDict.fromkeys(elems))
        table = {}
        for e1 in uniq:
            for e2 in uniq:
                if f"{e1}{e2}" in keys:
                    entry = self.bond_cut[f"{e1}{e2}"]
                elif f"{e2}{e1}" in keys:
                    entry = self.bond_cut[f"{e2}{e1}"]
                else:
                    entry = None
                # reference: get_cutoff(...)["NN"] with get_cutoff -> None  => TypeError
                table[(e1, e2)] = entry["NN"]
        idx = np.array([uniq.index(e) for e in elems])
        grid = np.array([[table[(e1, e2)] for e2 in uniq] for e1 in uniq], dtype=float)
        return grid[idx[:, None], idx[None, :]]

    def get_bond_mat(self):
        cut = self._cutoff_matrix()[None, :, :]
        dist = self.dist_mat
        bond = (dist < cut) & (dist > 0)
        # re-decide near-ties with the reference's per-pair arithmetic
        scale = np.maximum(np.abs(cut), 1.0)
        near = (np.abs(dist - cut) <= 64 * np.finfo(float).eps * scale) | ((dist > 0) & (dist < 1e-13))
        for image_i, i, j in zip(*np.nonzero(near)):
            v = self._pair_vec_exact(image_i, i, j)
            dd = np.linalg.norm(v)
            self.dist_mat_vec[image_i, i, j] = v
            self.dist_mat[image_i, i, j] = dd
            bond[image_i, i, j] = (dd < cut[0, i, j]) and (dd > 0)
        return bond

    def get_dir_cos(self, image_i, atoms_i, atom_j):
        v = self.dist_mat_vec[image_i, atoms_i, atom_j, :]
        nrm = np.linalg.norm(v)
        if nrm == 0:
            return 0, 0, 0
        return v / nrm

    def get_dir_cos_all(self):
        nrm = np.linalg.norm(self.dist_mat_vec, axis=-1)
        nrm[nrm == 0] = 1e-10
        return self.dist_mat_vec / nrm[..., None]

    # -- bookkeeping ------------------------------------------------------------------------
    def get_lattice(self):
        return self.lattice.get_matrix()

    def get_pos(self):
        return np.concatenate([a.pos for a in self.atoms]).ravel()

    def get_elements(self):
        return list(OrderedDict.fromkeys(a.element for a in self.atoms))

    def bond_list(self):
        """Directed bonds (image, i, j) in the order `calc_ham_wo_k` visits them
        (atom_1 outer, atom_2, image inner; hamiltonian.py:447-454)."""
        img, i, j = np.nonzero(self.bond_mat)
        order = np.lexsort((img, j, i))
        return np.stack([img[order], i[order], j[order]], axis=1).astype(np.int32)

    def get_supercell(self, sc, vac=(0, 0, 0)):
        raise NotImplementedError("get_supercell needs pymatgen; outside the k-point hot path (SURVEY.md §2 row 5)")

    @staticmethod
    def read_poscar(file_name="./POSCAR", kwargs={}):
        raise NotImplementedError  # same as the reference (structure.py:152-155)
