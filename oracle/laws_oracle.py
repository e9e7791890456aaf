"""ORACLE (test infrastructure only): distance laws V(d), scalar restatement.

Reference: ScalingLaw.__call__ (pysktb/scaling.py:106-113) = raw(d) * cosine cutoff (:60-75);
raw values: Harrison :193, PowerLaw :218, Exponential :248, GSP :286-290, Polynomial :327-329,
Tabulated :380 (scipy CubicSpline, not-a-knot, extrapolating - third-party, the same entry point the reference calls at
:373), Custom :425 (the user's function), Constant :170.  Scalars only (the path evaluates one bond length at a time, system.py:206-210).
"""
import numpy as np


def _cut(d, cutoff, width):
    if cutoff is None:
        return 1.0
    rs = cutoff - width
    if d > cutoff:
        return 0.0
    if d >= rs:
        return 0.5 + 0.5 * np.cos(np.pi * (d - rs) / (cutoff - rs))
    return 1.0


def make(kind, cutoff=None, smooth_width=0.5, **p):
    if kind == "Tabulated":
        from scipy.interpolate import CubicSpline
        spline = CubicSpline(np.asarray(p["distances"]), np.asarray(p["values"]), extrapolate=True)
        return lambda d: float(spline(np.float64(d)) * _cut(np.float64(d), cutoff, smooth_width))
    if kind == "Custom":
        return lambda d: float(p["func"](np.float64(d)) * _cut(np.float64(d), cutoff, smooth_width))
    raw = {
        "Harrison": lambda d: p["V0"] * (p["d0"] / d) ** 2,
        "PowerLaw": lambda d: p["V0"] * (p["d0"] / d) ** p.get("eta", 2.0),
        "Exponential": lambda d: p["V0"] * np.exp(-p.get("alpha", 1.0) * (d - p["d0"])),
        "GSP": lambda d: p["V0"] * (p["d0"] / d) ** p.get("n", 2.0) * np.exp(
            p.get("n", 2.0) * (-((d / p.get("dc", 3.5)) ** p.get("nc", 4.0)) + (p["d0"] / p.get("dc", 3.5)) ** p.get("nc", 4.0))),
        "Polynomial": lambda d: np.polyval(np.asarray(p["coeffs"], float)[::-1], d - p["d0"]),
        "Constant": lambda d: float(p["value"]),
    }[kind]
    return lambda d: float(raw(np.float64(d)) * _cut(np.float64(d), cutoff, smooth_width))
