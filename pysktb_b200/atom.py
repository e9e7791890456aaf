"""Atom record for the tight-binding model (host-side setup).

Mirrors the public surface of the reference `Atom` (pysktb/atom.py:4-78):
`Atom(element, pos, orbitals=None)`, `ORBITALS_ALL` (the canonical 17-orbital
order that indexes the Slater-Koster table), `set_orbitals`.
"""
import numpy as np

_S = ("s",)
_P = ("px", "py", "pz")
_D = ("dxy", "dyz", "dxz", "dx2-y2", "dz2")
_F = ("fz3", "fxz2", "fyz2", "fz(x2-y2)", "fxyz", "fx(x2-3y2)", "fy(3x2-y2)")
_S2 = ("S",)


class Atom:
    # canonical order: s | p | d | f | S  (reference atom.py:21-39)
    ORBITALS_ALL = list(_S + _P + _D + _F + _S2)
    ORBITAL_FAMILIES = {"s": list(_S), "p": list(_P), "d": list(_D), "f": list(_F), "S": list(_S2)}
    ORBITAL_L = {**{o: 0 for o in _S + _S2}, **{o: 1 for o in _P}, **{o: 2 for o in _D}, **{o: 3 for o in _F}}

    def __init__(self, element, pos, orbitals=None):
        self.element = element
        self.pos = np.array(pos)
        self.orbitals = None
        if orbitals is not None:
            self.set_orbitals(orbitals)

    def set_orbitals(self, orbitals=None):
        # same failure mode as the reference (atom.py:74): AssertionError("wrong orbitals")
        assert set(orbitals).issubset(set(Atom.ORBITALS_ALL)), "wrong orbitals"
        self.orbitals = orbitals

    def __repr__(self):
        return f"{self.element} {self.pos}"


ORBITAL_INDEX = {name: i for i, name in enumerate(Atom.ORBITALS_ALL)}
