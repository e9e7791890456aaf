"""sktb-b200: B200-native k-point hot path of pysktb behind pysktb's own Python surface.

    from pysktb_b200 import Structure, Atom, Lattice, Hamiltonian, GreensFunction, SurfaceGreensFunction

Same names and call signatures as `pysktb` (reference pysktb/__init__.py:1-53) for the hot path:
Hamiltonian(structure, params) -> get_kpts / solve_kpath / solve_k / get_ham, GreensFunction.dos / ldos /
ldos_by_orbital / spectral_kpath, SurfaceGreensFunction.edge_dos / edge_spectral_kpath.  All numerics run
in hand-written sm_100a kernels (pysktb_b200/libsktb.so, C ABI in include/sktb.h); importing the package
is cheap and does not touch CUDA, but constructing a Hamiltonian without the built library or without a
GPU raises (no CPU fallback).
"""
from .atom import Atom
from .lattice import Lattice
from .structure import Structure
from .system import System
from . import scaling
from .scaling import (ScalingLaw, Constant, Harrison, PowerLaw, Exponential, GSP, Polynomial, Tabulated,
                      Custom, ensure_scaling)
from .hamiltonian import Hamiltonian
from .greens import GreensFunction, SurfaceGreensFunction
from . import dist
from .forces import Forces

__version__ = "0.1.0"
__all__ = ["Structure", "Atom", "Lattice", "System", "Hamiltonian", "GreensFunction", "SurfaceGreensFunction", "Forces",
           "ScalingLaw", "Constant", "Harrison", "PowerLaw", "Exponential", "GSP", "Polynomial", "Tabulated",
           "Custom", "ensure_scaling", "dist"]
