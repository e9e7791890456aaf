"""Lattice (host-side setup).  Reference: pysktb/lattice.py:4-87.

Rows of `matrix` are the lattice vectors already multiplied by the lattice constant.
`get_rec_lattice()` is inv(matrix).T (no 2*pi), the convention `calc_g` uses as
`k_frac @ rec` (hamiltonian.py:279-280).
"""
import numpy as np


class Lattice:
    def __init__(self, *args):
        matrix, lat_const = args
        self.matrix = np.array(matrix) * lat_const
        # cell parameters are taken from the UNSCALED vectors, as the reference does (lattice.py:19-32)
        v = [np.asarray(matrix[i], dtype=float) for i in range(3)]
        self.a, self.b, self.c = (float(np.linalg.norm(x)) for x in v)

        def _angle(p, q):
            return np.arctan2(np.linalg.norm(np.cross(p, q)), np.dot(p, q))

        self.alpha, self.beta, self.gamma = _angle(v[1], v[2]), _angle(v[2], v[0]), _angle(v[0], v[1])

    def get_matrix(self):
        return self.matrix

    def get_rec_lattice(self):
        return np.linalg.inv(self.matrix).T

    def __repr__(self):
        return " ".join(str(x) for x in (self.a, self.b, self.c, self.alpha, self.beta, self.gamma))
