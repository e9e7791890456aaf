"""sktb-b200: B200-native k-point hot path of pysktb behind pysktb's own Python surface."""
from .atom import Atom
from .lattice import Lattice
from .structure import Structure
from .system import System
from . import scaling
from .scaling import (ScalingLaw, Constant, Harrison, PowerLaw, Exponential, GSP, Polynomial, Tabulated,
                      Custom, ensure_scaling)
