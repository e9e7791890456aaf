import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def oracle_model(spec):
    """Instantiate the numpy oracle from a workloads spec (scaling-law placeholders -> plain callables)."""
    from oracle.sktb_oracle import OracleModel
    from oracle import laws_oracle

    params = {k: {kk: (laws_oracle.make(v[1], **v[2]) if isinstance(v, tuple) and v and v[0] == "law" else v)
                  for kk, v in t.items()} for k, t in spec["params"].items()}
    return OracleModel(spec["lattice"], spec["a"], spec["atoms"], spec["bond_cut"], params, spec["periodicity"])


GOLDEN_SPECS = {
    "C1_cubic_sp3": ("cubic_sp3", {}), "C2_graphene_pz": ("graphene_pz", {}), "C3_sb_buckled": ("sb_buckled", {}),
    "T_ce_spdf_1atom": ("ce_spdf_1atom", {}), "C4_ce_spdf_2atom": ("ce_spdf_2atom", {}),
    "ce_spf_example": ("ce_spf_example", {}), "C5_zigzag_8": ("zigzag_ribbon", {"n_chains": 8}),
    "C5_zigzag_32": ("zigzag_ribbon", {"n_chains": 32}), "sp_chain": ("sp_chain", {}),
    "lowsym_spd_laws": ("lowsym_spd", {"laws": True}), "lowsym_spd": ("lowsym_spd", {"laws": False}),
    "d_soc_nonherm": ("d_soc_nonhermitian", {}),
    "lowsym_spd_misc": ("lowsym_spd", {"laws": "misc"}),          # Polynomial / Tabulated / Custom laws (scaling.py:312-447)
}

# oracle-only cases (no fixture file): padded / non-SOC variants of the 32-wide register kernels
ORACLE_ONLY_SPECS = {
    "ce_partial_13": ("ce_partial", {"n_orb": 13}), "ce_partial_14": ("ce_partial", {"n_orb": 14}),
    "ce_partial_15": ("ce_partial", {"n_orb": 15}), "zigzag_13": ("zigzag_ribbon", {"n_chains": 13}),
    "zigzag_14": ("zigzag_ribbon", {"n_chains": 14}), "zigzag_16": ("zigzag_ribbon", {"n_chains": 16}),
    # 32 < N <= 64: CTA-resident kernel, padded and full
    "zigzag_17": ("zigzag_ribbon", {"n_chains": 17}), "zigzag_20": ("zigzag_ribbon", {"n_chains": 20}),
    "zigzag_31": ("zigzag_ribbon", {"n_chains": 31}),
    "cubic_s": ("cubic_s", {}),
}


def spec_for(golden_name):
    from pysktb_b200 import workloads as W
    fn, kw = GOLDEN_SPECS[golden_name] if golden_name in GOLDEN_SPECS else ORACLE_ONLY_SPECS[golden_name]
    return W.ALL[fn](**kw)
