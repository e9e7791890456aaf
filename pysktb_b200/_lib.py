"""ctypes binding of libsktb.so (include/sktb.h).

The library is built in-tree (pysktb_b200/csrc/Makefile -> pysktb_b200/libsktb.so).  There is no CPU
fallback: if the library is missing, or no CUDA device is present, every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKTB_LIB_PATH") or os.path.join(_HERE, "libsktb.so")     # override: A/B builds of the same ABI (tools/ab_*.sh)

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class ModelDesc(C.Structure):
    """Mirror of `sktb_model_desc` (include/sktb.h)."""
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_orb", C.c_int32),
        ("atom_orb_start", c_i32p), ("orb_kind", c_i32p), ("onsite", c_f64p),
        ("n_bonds", C.c_int32),
        ("bond_image", c_i32p), ("bond_i", c_i32p), ("bond_j", c_i32p),
        ("bond_vec", c_f64p), ("bond_lmn", c_f64p), ("bond_V", c_f64p), ("rec_lattice", c_f64p),
        ("soc_nnz", C.c_int32), ("soc_row", c_i32p), ("soc_col", c_i32p), ("soc_val", c_f64p),
    ]


class SktbError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "sktb_last_error": (C.c_char_p, []),
    "sktb_version": (C.c_char_p, []),
    "sktb_v_names": (C.c_char_p, []),
    "sktb_last_kernel_info": (C.c_char_p, []),
    "sktb_launch_count": (C.c_longlong, []),
    "sktb_plan_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "sktb_plan_destroy": (None, [C.c_void_p]),
    "sktb_plan_hoppings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "sktb_plan_asym_bound": (C.c_int, [C.c_void_p, c_f64p]),
    "sktb_sk_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "sktb_sk_table_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "sktb_kmesh": (C.c_int, [c_i32p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "sktb_hk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sktb_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                             C.c_void_p, C.c_void_p]),
    "sktb_solve_mesh": (C.c_int, [C.c_void_p, c_i32p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_void_p]),
    "sktb_heev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sktb_heev_twostage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sktb_solve_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, c_i32p]),
    "sktb_dos": (C.c_int, [C.c_void_p, C.c_void_p, c_i32p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_double,
                           C.c_int32, c_i32p, c_i32p, C.c_void_p, c_i32p, C.c_void_p]),
    "sktb_spectral": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_int32,
                                c_i32p, C.c_void_p, c_i32p, C.c_void_p]),
    "sktb_plan_herm_defect": (C.c_int, [C.c_void_p, C.c_int32, c_f64p]),
    "sktb_gf_dos": (C.c_int, [C.c_void_p, C.c_void_p, c_i32p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_double,
                              C.c_int32, c_i32p, c_i32p, C.c_int32, C.c_void_p, C.c_void_p]),
    "sktb_gf_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "sktb_gf_retarded": (C.c_int, [C.c_void_p, c_f64p, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "sktb_gauss_dos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "sktb_band_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "sktb_fp64_peak": (C.c_int, [C.c_int, c_f64p, c_f64p]),
    "sktb_last_solve_time": (C.c_int, [C.c_void_p, c_f64p, c_i32p]),
    "sktb_ql_failures": (C.c_int, [C.c_void_p, c_i32p]),
}

EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load libsktb.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SktbError(
                f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). pysktb_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise SktbError(lib().sktb_last_error().decode())


def v_names():
    return lib().sktb_v_names().decode().split(",")


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_i32(a):
    return a.ctypes.data_as(c_i32p)


def ptr_f64(a):
    return a.ctypes.data_as(c_f64p)
