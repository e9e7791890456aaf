"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the reference's Slater-Koster two-centre table `get_hop_int`
(pysktb/_params.py:36-513): 25 `V_*` parameters + direction cosines (l, m, n) -> 17x17 real table in
the canonical orbital order  s | px py pz | dxy dyz dxz dx2-y2 dz2 | fz3 fxz2 fyz2 fz(x2-y2) fxyz
fx(x2-3y2) fy(3x2-y2) | S.

It restates what the reference *computes*, including its deviations from textbook SK tables
(SURVEY.md Q2-Q4):
  * px-f is assigned three times; only the last set survives            (_params.py:249-257)
  * transposition signs: p-f "+", d-f "-", f-f symmetric                 (_params.py:279-282, 380-383, 471-474)
  * fxz2/fyz2 diagonal carries a term with no V factor                   (_params.py:396-403)
  * the f-f diagonals 12, 14, 15 collapse through conditional-expression
    precedence (`a + b if c else 0 + d`)                                 (_params.py:406-423)
Pinned against golden vectors produced by running the reference itself
(tests/golden/sk_table.npz, made by oracle/make_golden.py).
"""
import numpy as np

V_NAMES = (
    "V_sss", "V_sps", "V_pps", "V_ppp", "V_sds", "V_pds", "V_pdp", "V_dds", "V_ddp", "V_ddd",
    "V_sfs", "V_pfs", "V_pfp", "V_dfs", "V_dfp", "V_dfd", "V_ffs", "V_ffp", "V_ffd", "V_fff",
    "V_SSs", "V_sSs", "V_Sps", "V_Sds", "V_Sfs",
)
EPS_AXIS = 1e-10  # the reference's guard on l^2+m^2


def _s_like_row(V_p, V_d, V_f, l, m, n):
    """Row of an s-type orbital against p, d, f  (_params.py:120-163 and :486-509)."""
    l2, m2, n2 = l * l, m * m, n * n
    r3, r15 = np.sqrt(3.0), np.sqrt(15.0)
    p = [l * V_p, m * V_p, n * V_p]
    d = [
        r3 * (l * m) * V_d,
        r3 * (m * n) * V_d,
        r3 * (n * l) * V_d,
        r3 / 2.0 * (l2 - m2) * V_d,
        (n2 - 0.5 * (l2 + m2)) * V_d,
    ]
    f = [
        n * (5 * n2 - 3) / 2.0 * V_f,
        np.sqrt(3.0 / 8.0) * l * (5 * n2 - 1) * V_f,
        np.sqrt(3.0 / 8.0) * m * (5 * n2 - 1) * V_f,
        r15 / 2.0 * n * (l2 - m2) * V_f,
        r15 * (l * m * n) * V_f,
        np.sqrt(5.0 / 8.0) * l * (l2 - 3 * m2) * V_f,
        np.sqrt(5.0 / 8.0) * m * (3 * l2 - m2) * V_f,
    ]
    return p, d, f


def sk_table(V, l, m, n):
    """V: dict of V_* (missing -> 0; unknown V_* key -> TypeError like the reference's kwargs)."""
    for key in V:
        if key not in V_NAMES:
            raise TypeError(f"get_hop_int() got an unexpected keyword argument '{key}'")
    g = {name[2:]: V.get(name, 0) for name in V_NAMES}
    T = np.zeros((17, 17))
    with np.errstate(all="ignore"):
        _fill(T, g, np.float64(l), np.float64(m), np.float64(n))
    return T


def _fill(T, g, l, m, n):
    l2, m2, n2 = l * l, m * m, n * n
    lm, mn, nl = l * m, m * n, n * l
    lmn = l * m * n
    dm = l2 - m2            # l^2 - m^2
    sp = l2 + m2            # l^2 + m^2
    off_axis = sp > EPS_AXIS
    r3 = np.sqrt(3.0)
    r2, r6, r10, r30 = np.sqrt(2.0), np.sqrt(6.0), np.sqrt(10.0), np.sqrt(30.0)
    a = 5 * n2 - 1          # recurring cubic-harmonic factors
    b = 5 * n2 - 3

    # ---- s row/col ------------------------------------------------------------------------
    T[0, 0] = g["sss"]
    p, d, f = _s_like_row(g["sps"], g["sds"], g["sfs"], l, m, n)
    T[0, 1:4], T[0, 4:9], T[0, 9:16] = p, d, f
    T[1:4, 0] = -T[0, 1:4]          # odd
    T[4:9, 0] = T[0, 4:9]           # even
    T[9:16, 0] = -T[0, 9:16]        # odd

    # ---- p-p --------------------------------------------------------------------------------
    pps, ppp = g["pps"], g["ppp"]
    T[1, 1] = l2 * pps + (1.0 - l2) * ppp
    T[2, 2] = m2 * pps + (1.0 - m2) * ppp
    T[3, 3] = n2 * pps + (1.0 - n2) * ppp
    T[1, 2] = T[2, 1] = lm * (pps - ppp)
    T[1, 3] = T[3, 1] = nl * (pps - ppp)
    T[2, 3] = T[3, 2] = mn * (pps - ppp)

    # ---- p-d --------------------------------------------------------------------------------
    pds, pdp = g["pds"], g["pdp"]
    cross = lmn * (r3 * pds - 2 * pdp)
    z2 = n2 - 0.5 * sp
    T[1, 4:9] = [
        r3 * l2 * m * pds + m * (1.0 - 2 * l2) * pdp,
        cross,
        r3 * l2 * n * pds + n * (1.0 - 2 * l2) * pdp,
        0.5 * r3 * l * dm * pds + l * (1.0 - l2 + m2) * pdp,
        l * z2 * pds - r3 * l * n2 * pdp,
    ]
    T[2, 4:9] = [
        r3 * m2 * l * pds + l * (1.0 - 2 * m2) * pdp,
        r3 * m2 * n * pds + n * (1.0 - 2 * m2) * pdp,
        cross,
        0.5 * r3 * m * dm * pds - m * (1.0 + l2 - m2) * pdp,
        m * z2 * pds - r3 * m * n2 * pdp,
    ]
    T[3, 4:9] = [
        cross,
        r3 * n2 * m * pds + m * (1.0 - 2 * n2) * pdp,
        r3 * n2 * l * pds + l * (1.0 - 2 * n2) * pdp,
        0.5 * r3 * n * dm * pds - n * dm * pdp,
        n * z2 * pds + r3 * n * sp * pdp,
    ]
    T[4:9, 1:4] = -T[1:4, 4:9].T

    # ---- p-f (surviving assignments only) ---------------------------------------------------
    pfs, pfp = g["pfs"], g["pfp"]
    q = sp if off_axis else 1
    x3 = l2 - 3 * m2        # from fx(x2-3y2)
    y3 = 3 * l2 - m2        # from fy(3x2-y2)
    T[1, 9:16] = [
        r6 / 4.0 * l * a * pfs - r6 / 2.0 * l * n2 * pfp,
        (3.0 / 4.0 * l2 * a - a / 2.0) * pfs + (l2 * (1 - 5 * n2 / 2.0) + n2) * pfp,
        3.0 / 4.0 * lm * a * pfs - 5.0 / 2.0 * lm * n2 * pfp,
        r10 / 4.0 * nl * dm * pfs - r10 / 2.0 * nl * (1 + dm / 2.0) * pfp,
        r10 / 2.0 * mn * l2 * pfs - r10 / 2.0 * mn * (1 - l2) * pfp,
        (r30 / 8.0 * l * x3 * pfs - r30 / 4.0 * l * x3 * (1 - l2 + m2) / q * pfp) if off_axis else 0,
        (r10 / 8.0 * m * y3 * pfs + r10 / 4.0 * m * (1 - y3 * (1 + l2 - m2) / q) * pfp) if off_axis else 0,
    ]
    T[2, 9:16] = [
        r6 / 4.0 * m * a * pfs - r6 / 2.0 * m * n2 * pfp,
        3.0 / 4.0 * lm * a * pfs - 5.0 / 2.0 * lm * n2 * pfp,
        (3.0 / 4.0 * m2 * a - a / 2.0) * pfs + (m2 * (1 - 5 * n2 / 2.0) + n2) * pfp,
        r10 / 4.0 * mn * dm * pfs + r10 / 2.0 * mn * (1 - dm / 2.0) * pfp,
        r10 / 2.0 * nl * m2 * pfs - r10 / 2.0 * nl * (1 - m2) * pfp,
        (r30 / 8.0 * m * x3 * pfs + r30 / 4.0 * m * x3 * (1 + l2 - m2) / q * pfp) if off_axis else 0,
        (r10 / 8.0 * l * y3 * pfs - r10 / 4.0 * l * (1 + y3 * (1 - l2 + m2) / q) * pfp) if off_axis else 0,
    ]
    T[3, 9:16] = [
        (n2 * b + 3.0 / 2.0 * (1 - n2)) * pfs + r6 * n2 * (1 - n2) * pfp,
        r6 / 4.0 * nl * b * pfs + r6 / 2.0 * nl * (1 - 5 * n2 + 4 * n2) * pfp,
        r6 / 4.0 * mn * b * pfs + r6 / 2.0 * mn * (1 - 5 * n2 + 4 * n2) * pfp,
        r10 / 4.0 * n2 * dm * pfs + r10 / 4.0 * dm * (1 - 2 * n2) * pfp,
        r10 / 2.0 * lm * n2 * pfs - r10 / 2.0 * lm * (1 - n2) * pfp,
        r30 / 8.0 * nl * x3 * pfs + r30 / 4.0 * nl * (1 - x3 / 2.0) * pfp,
        r10 / 8.0 * mn * y3 * pfs + r10 / 4.0 * mn * (1 - y3 / 2.0) * pfp,
    ]
    T[9:16, 1:4] = T[1:4, 9:16].T          # same sign (reference's choice)

    # ---- d-d --------------------------------------------------------------------------------
    dds, ddp, ddd = g["dds"], g["ddp"], g["ddd"]
    mix = 3 * dds - 4 * ddp + ddd
    pd_ = ddp - ddd
    T[4, 4] = l2 * m2 * mix + (l2 + m2) * ddp + n2 * ddd
    T[5, 5] = m2 * n2 * mix + (m2 + n2) * ddp + l2 * ddd
    T[6, 6] = n2 * l2 * mix + (n2 + l2) * ddp + m2 * ddd
    T[4, 5] = T[5, 4] = l * m2 * n * mix + nl * pd_
    T[4, 6] = T[6, 4] = n * l2 * m * mix + mn * pd_
    T[5, 6] = T[6, 5] = m * n2 * l * mix + lm * pd_
    T[4, 7] = T[7, 4] = 0.5 * lm * dm * mix
    T[5, 7] = T[7, 5] = 0.5 * mn * (dm * mix - 2 * pd_)
    T[6, 7] = T[7, 6] = 0.5 * nl * (dm * mix + 2 * pd_)
    T[4, 8] = T[8, 4] = r3 * (lm * z2 * dds - 2 * lm * n2 * ddp + 0.5 * lm * (1.0 + n2) * ddd)
    T[5, 8] = T[8, 5] = r3 * (mn * z2 * dds + mn * (sp - n2) * ddp - 0.5 * mn * sp * ddd)
    T[6, 8] = T[8, 6] = r3 * (nl * z2 * dds + nl * (sp - n2) * ddp - 0.5 * nl * sp * ddd)
    T[7, 7] = 0.25 * dm ** 2 * mix + sp * ddp + n2 * ddd
    T[8, 8] = 0.75 * sp ** 2 * ddd + 3 * sp * n2 * ddp + 0.25 * (sp - 2 * n2) ** 2 * dds
    T[7, 8] = T[8, 7] = r3 * 0.25 * dm * (n2 * (2 * dds - 4 * ddp + ddd) + ddd - sp * dds)

    # ---- d-f --------------------------------------------------------------------------------
    dfs, dfp = g["dfs"], g["dfp"]          # V_dfd is accepted but never used by the reference
    T[4, 9:16] = [
        r10 / 4.0 * lm * a * dfs - r10 / 2.0 * lm * n2 * dfp,
        r30 / 8.0 * m * (l2 * a - a / 3.0) * dfs + r30 / 4.0 * m * (n2 - l2 * a / 3.0) * dfp,
        r30 / 8.0 * l * (m2 * a - a / 3.0) * dfs + r30 / 4.0 * l * (n2 - m2 * a / 3.0) * dfp,
        r2 / 2.0 * lm * n * dm * dfs + r2 * lm * n * (1 - dm / 2.0) * dfp,
        r2 * n * l2 * m2 * dfs - r2 * n * (l2 + m2 - 2 * l2 * m2) * dfp,
        (r6 / 8.0 * m * (l2 * x3 / sp) * dfs - r6 / 4.0 * m * l2 * dfp) if off_axis else 0,
        (r6 / 8.0 * l * (m2 * y3 / sp) * dfs - r6 / 4.0 * l * m2 * dfp) if off_axis else 0,
    ]
    T[5, 9:16] = [
        r10 / 4.0 * mn * b * dfs + r10 / 2.0 * mn * (1 - n2) * dfp,
        r30 / 8.0 * lmn * a * dfs - r30 / 4.0 * lmn * (1 + n2) * dfp,
        r30 / 8.0 * (m2 * n * a - n * a / 3.0) * dfs + r30 / 4.0 * n * (m2 - a * m2 / 3.0) * dfp,
        r2 / 2.0 * mn * n * dm * dfs - r2 * mn * (l2 - m2 / 2.0) * dfp,
        r2 * l * m2 * n2 * dfs - r2 * l * (m2 + n2 - 2 * m2 * n2) * dfp,
        r6 / 8.0 * mn * x3 * dfs - r6 / 4.0 * mn * dm * dfp,
        r6 / 8.0 * nl * y3 * dfs - r6 / 4.0 * nl * dm * dfp,
    ]
    T[6, 9:16] = [
        r10 / 4.0 * nl * b * dfs + r10 / 2.0 * nl * (1 - n2) * dfp,
        r30 / 8.0 * (l2 * n * a - n * a / 3.0) * dfs + r30 / 4.0 * n * (l2 - a * l2 / 3.0) * dfp,
        r30 / 8.0 * lmn * a * dfs - r30 / 4.0 * lmn * (1 + n2) * dfp,
        r2 / 2.0 * nl * n * dm * dfs + r2 * nl * (m2 - l2 / 2.0) * dfp,
        r2 * m * l2 * n2 * dfs - r2 * m * (l2 + n2 - 2 * l2 * n2) * dfp,
        r6 / 8.0 * nl * x3 * dfs + r6 / 4.0 * nl * dm * dfp,
        r6 / 8.0 * mn * y3 * dfs + r6 / 4.0 * mn * dm * dfp,
    ]
    q2 = sp ** 2 if off_axis else 1
    T[7, 9:16] = [
        r10 / 8.0 * n * dm * a * dfs - r10 / 4.0 * n * dm * n2 * dfp,
        r30 / 16.0 * l * dm * a * dfs - r30 / 8.0 * l * (n2 * dm - l2 + m2) * dfp,
        r30 / 16.0 * m * dm * a * dfs + r30 / 8.0 * m * (n2 * dm + l2 - m2) * dfp,
        (r2 / 4.0 * n * dm ** 2 * dfs + r2 / 2.0 * n * sp * (1 - dm ** 2 / q2) * dfp) if off_axis else 0,
        r2 / 2.0 * lm * n * dm * dfs - r2 * lm * n * dfp,
        (r6 / 16.0 * l * dm * x3 / q * dfs) if off_axis else 0,
        (r6 / 16.0 * m * dm * y3 / q * dfs) if off_axis else 0,
    ]
    w = a * (3 * n2 - 1) - 6 * n2
    T[8, 9:16] = [
        (n2 * b / 2.0 + 3.0 / 4.0 * (1 - n2) * a) * dfs + r6 / 2.0 * n2 * (1 - n2) * dfp,
        r2 / 8.0 * l * w * dfs + r3 / 2.0 * l * n2 * (1 - n2) * dfp,
        r2 / 8.0 * m * w * dfs + r3 / 2.0 * m * n2 * (1 - n2) * dfp,
        r30 / 8.0 * n * dm * z2 * dfs + r30 / 4.0 * n * dm * (1 - n2) * dfp,
        r30 / 4.0 * lm * n * z2 * dfs - r30 / 2.0 * lm * n * (1 - n2) * dfp,
        r10 / 16.0 * l * x3 * (3 * n2 - 1) * dfs - r10 / 8.0 * l * x3 * n2 * dfp,
        r10 / 16.0 * m * y3 * (3 * n2 - 1) * dfs - r10 / 8.0 * m * y3 * n2 * dfp,
    ]
    T[9:16, 4:9] = -T[4:9, 9:16].T

    # ---- f-f --------------------------------------------------------------------------------
    ffs, ffp, ffd = g["ffs"], g["ffp"], g["ffd"]   # V_fff is accepted but never used by the reference
    T[9, 9] = n2 * b ** 2 / 4.0 * ffs + 3.0 / 2.0 * n2 * (1 - n2) * a * ffp + 15.0 / 4.0 * (1 - n2) ** 2 * ffd
    # NB the middle-left term has no V factor in the reference (_params.py:397, :402)
    T[10, 10] = 3.0 / 8.0 * l2 * a ** 2 * ffs + l2 * (1 - 5 * n2 / 2.0) ** 2 + n2 * (1 - l2) * ffp + (1 - l2) * (1 - n2) * ffd
    T[11, 11] = 3.0 / 8.0 * m2 * a ** 2 * ffs + m2 * (1 - 5 * n2 / 2.0) ** 2 + n2 * (1 - m2) * ffp + (1 - m2) * (1 - n2) * ffd
    # conditional-expression precedence: (A + B) if c else (0 + C)
    T[12, 12] = (15.0 / 4.0 * n2 * dm ** 2 * ffs + (sp - n2 * dm) ** 2 / sp * ffp) if off_axis else (0 + l2 * m2 * ffd)
    s3 = l2 * m2 + m2 * n2 + n2 * l2
    T[13, 13] = 15.0 * l2 * m2 * n2 * ffs + (s3 - 4 * l2 * m2 * n2) * ffp + (s3 - l2 * m2 * n2) * ffd
    # chains of `x if c else 0 + y if c else 0 + z if c else 0` keep only the first term
    T[14, 14] = (5.0 / 8.0 * l2 * x3 ** 2 / sp ** 2 * ffs) if off_axis else 0
    T[15, 15] = (5.0 / 8.0 * m2 * y3 ** 2 / sp ** 2 * ffs) if off_axis else 0

    T[9, 10] = r6 / 8.0 * nl * a * b * ffs + r6 / 4.0 * nl * a * (1 - n2) * ffp
    T[9, 11] = r6 / 8.0 * mn * a * b * ffs + r6 / 4.0 * mn * a * (1 - n2) * ffp
    T[9, 12] = r10 / 8.0 * n2 * dm * b * ffs + r10 / 4.0 * dm * (1 - n2) * a * ffp
    T[9, 13] = r10 / 4.0 * lm * n2 * b * ffs + r10 / 2.0 * lm * (1 - n2) * a * ffp
    T[9, 14] = r30 / 16.0 * nl * x3 * b * ffs
    T[9, 15] = r30 / 16.0 * mn * y3 * b * ffs
    T[10, 11] = 3.0 / 8.0 * lm * a ** 2 * ffs + lm * (1 - 5 * n2 / 2.0) ** 2 * ffp
    T[10, 12] = r10 / 16.0 * l * dm * a * ffs
    T[10, 13] = r10 / 8.0 * m * n * l * a * ffs
    T[11, 12] = r10 / 16.0 * m * dm * a * ffs
    T[11, 13] = r10 / 8.0 * l * n * m * a * ffs
    T[12, 13] = r6 / 4.0 * lm * n * dm * ffs
    T[13, 14] = r6 / 8.0 * m * n * x3 * ffs
    T[13, 15] = r6 / 8.0 * l * n * y3 * ffs
    T[14, 15] = (5.0 / 8.0 * lm * x3 * y3 / sp ** 2 * ffs) if off_axis else 0
    for i in range(9, 16):
        for j in range(i + 1, 16):
            T[j, i] = T[i, j]

    # ---- excited s (S) --------------------------------------------------------------------------
    T[16, 16] = g["SSs"]
    T[0, 16] = T[16, 0] = g["sSs"]
    p, d, f = _s_like_row(g["Sps"], g["Sds"], g["Sfs"], l, m, n)
    T[16, 1:4], T[16, 4:9], T[16, 9:16] = p, d, f
    T[1:4, 16] = -T[16, 1:4]
    T[4:9, 16] = T[16, 4:9]
    T[9:16, 16] = -T[16, 9:16]
