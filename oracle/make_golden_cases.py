"""ORACLE (test infrastructure only): the law cases of tests/golden/laws_misc.npz, shared by the generator
(oracle/make_golden.py, build container) and the tests (which run where /root/reference does not exist)."""
import numpy as np

from pysktb_b200 import workloads as W

LAW_CASES = {
    "poly": ("Polynomial", dict(coeffs=[-2.7, 1.5, -0.3, 0.07], d0=1.42)),
    "poly_cut": ("Polynomial", dict(coeffs=[-0.6, 0.35, -0.08, 0.01], d0=2.4, cutoff=4.4)),
    "poly_const": ("Polynomial", dict(coeffs=[-0.5], d0=2.4)),
    "tab": ("Tabulated", dict(distances=[1.2, 1.4, 1.6, 1.8, 2.0], values=[-3.2, -2.7, -2.3, -1.9, -1.6])),
    "tab_cut": ("Tabulated", dict(distances=[1.8, 2.2, 2.6, 3.0, 3.4, 3.8], values=[1.55, 1.08, 0.77, 0.55, 0.4, 0.29], cutoff=3.6, smooth_width=0.4)),
    "custom": ("Custom", dict(func=W.custom_law_func, d0=2.4)),
    "custom_cut": ("Custom", dict(func=W.custom_law_func, d0=2.4, cutoff=3.3, smooth_width=0.7)),
    "const": ("Constant", dict(value=-0.37)),
}
LAW_DISTANCES = np.concatenate([np.linspace(1.0, 5.0, 81), [2.4, 3.3, 3.6, 4.4, 2.6, 2.9 + 1e-12]])
