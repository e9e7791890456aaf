"""Load the REAL reference (read-only tree at /root/reference) - build container only.

`import pysktb` fails there because pysktb/__init__.py:44 pulls matplotlib (absent); we register an empty
package shell and import the submodules of the hot path directly.  The reference must be
used with `numba=0`: its `@nb.jit(nopython=False)` phase loop no longer compiles with numba 0.65, and
hamiltonian.py:299-316 is the identical plain-Python loop.

Nothing on the GPU box may call this (the tree does not exist there): it is used by oracle/make_golden.py
to write tests/golden/*.npz, and by tests that skip when the tree is absent.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SKTB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "pysktb"))


def load():
    """Returns a namespace with Lattice, Atom, Structure, System, Hamiltonian, GreensFunction,
    SurfaceGreensFunction, scaling (module), get_hop_int of the real reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "pysktb" not in sys.modules or not getattr(sys.modules["pysktb"], "_sktb_oracle_shell", False):
        pkg = types.ModuleType("pysktb")
        pkg.__path__ = [os.path.join(REF_ROOT, "pysktb")]
        pkg._sktb_oracle_shell = True
        sys.modules["pysktb"] = pkg
    from pysktb.hamiltonian import Hamiltonian
    from pysktb.greens import GreensFunction, SurfaceGreensFunction
    from pysktb.structure import Structure
    from pysktb.system import System
    from pysktb.atom import Atom
    from pysktb.lattice import Lattice
    from pysktb import scaling
    from pysktb._params import get_hop_int

    class _RefHam(Hamiltonian):
        def __init__(self, structure, inter, numba=0):
            super().__init__(structure, inter, numba=numba)

    return types.SimpleNamespace(Lattice=Lattice, Atom=Atom, Structure=Structure, System=System,
                                 Hamiltonian=_RefHam, GreensFunction=GreensFunction,
                                 SurfaceGreensFunction=SurfaceGreensFunction, scaling=scaling,
                                 get_hop_int=get_hop_int)
