// K4 `sk_table`: Slater-Koster two-centre table on the device (setup time, one thread per bond).
//
// Computes what the reference's get_hop_int computes (pysktb/_params.py:36-513): 25 V_* values and the
// direction cosines (l,m,n) of a bond -> 17x17 real table in canonical orbital order
//   0 s | 1-3 px py pz | 4-8 dxy dyz dxz dx2-y2 dz2 | 9-15 fz3 fxz2 fyz2 fz(x2-y2) fxyz fx(x2-3y2) fy(3x2-y2) | 16 S
// including the reference's departures from textbook tables: surviving px-f assignment
// (_params.py:249-257), transposition signs (p-f +, d-f -, f-f symmetric; :279-282,:380-383,:471-474), the
// V-less term on the fxz2/fyz2 diagonals (:397,:402) and the collapsed f-f diagonals 12/14/15 (:406-423).
// V_dfd and V_fff are accepted and unused, as in the reference.
#pragma once

namespace sktb {

enum VIdx {
    V_sss = 0, V_sps, V_pps, V_ppp, V_sds, V_pds, V_pdp, V_dds, V_ddp, V_ddd, V_sfs, V_pfs, V_pfp, V_dfs, V_dfp,
    V_dfd, V_ffs, V_ffp, V_ffd, V_fff, V_SSs, V_sSs, V_Sps, V_Sds, V_Sfs, V_COUNT
};

// Forward-mode derivative carrier for the Hellmann-Feynman forces: value + gradient with respect to the three Cartesian
// components of the bond vector.  sk_fill<SkDual> differentiates the table exactly as written (the reference's
// departures from the textbook tables included), where the reference itself falls back to a forward difference
// with eps = 1e-6 over the direction cosines (forces.py:283-297).
struct SkDual {
    double v, g[3];
    __host__ __device__ SkDual() : v(0.0), g{0.0, 0.0, 0.0} {}
    __host__ __device__ SkDual(double x) : v(x), g{0.0, 0.0, 0.0} {}
    __host__ __device__ SkDual(double x, double a, double b, double c) : v(x), g{a, b, c} {}
};
__host__ __device__ inline SkDual operator+(SkDual a, SkDual b) { return SkDual(a.v + b.v, a.g[0] + b.g[0], a.g[1] + b.g[1], a.g[2] + b.g[2]); }
__host__ __device__ inline SkDual operator-(SkDual a, SkDual b) { return SkDual(a.v - b.v, a.g[0] - b.g[0], a.g[1] - b.g[1], a.g[2] - b.g[2]); }
__host__ __device__ inline SkDual operator-(SkDual a) { return SkDual(-a.v, -a.g[0], -a.g[1], -a.g[2]); }
__host__ __device__ inline SkDual operator*(SkDual a, SkDual b) {
    return SkDual(a.v * b.v, a.g[0] * b.v + a.v * b.g[0], a.g[1] * b.v + a.v * b.g[1], a.g[2] * b.v + a.v * b.g[2]);
}
__host__ __device__ inline SkDual operator/(SkDual a, SkDual b) {
    const double q = a.v / b.v, ib = 1.0 / b.v;
    return SkDual(q, (a.g[0] - q * b.g[0]) * ib, (a.g[1] - q * b.g[1]) * ib, (a.g[2] - q * b.g[2]) * ib);
}
__host__ __device__ inline SkDual operator+(double a, SkDual b) { return SkDual(a) + b; }
__host__ __device__ inline SkDual operator+(SkDual a, double b) { return a + SkDual(b); }
__host__ __device__ inline SkDual operator-(double a, SkDual b) { return SkDual(a) - b; }
__host__ __device__ inline SkDual operator-(SkDual a, double b) { return a - SkDual(b); }
__host__ __device__ inline SkDual operator*(double a, SkDual b) { return SkDual(a * b.v, a * b.g[0], a * b.g[1], a * b.g[2]); }
__host__ __device__ inline SkDual operator*(SkDual a, double b) { return b * a; }
__host__ __device__ inline SkDual operator/(SkDual a, double b) { return a * (1.0 / b); }
__host__ __device__ inline SkDual operator/(double a, SkDual b) { return SkDual(a) / b; }
__host__ __device__ inline bool operator>(SkDual a, double b) { return a.v > b; }

// row of an s-like orbital (s or S) against p,d,f; `row` selects which orbital index it is
template <class R>
__device__ __forceinline__ void sk_s_row(R* T, int row, R Vp, R Vd, R Vf, R l, R m, R n) {
#define TT(i, j) T[(i) * 17 + (j)]
    const R l2 = l * l, m2 = m * m, n2 = n * n;
    const R r3 = sqrt(3.0), r15 = sqrt(15.0);
    TT(row, 1) = l * Vp;
    TT(row, 2) = m * Vp;
    TT(row, 3) = n * Vp;
    TT(row, 4) = r3 * (l * m) * Vd;
    TT(row, 5) = r3 * (m * n) * Vd;
    TT(row, 6) = r3 * (n * l) * Vd;
    TT(row, 7) = r3 / 2.0 * (l2 - m2) * Vd;
    TT(row, 8) = (n2 - 0.5 * (l2 + m2)) * Vd;
    TT(row, 9) = n * (5 * n2 - 3) / 2.0 * Vf;
    TT(row, 10) = sqrt(3.0 / 8.0) * l * (5 * n2 - 1) * Vf;
    TT(row, 11) = sqrt(3.0 / 8.0) * m * (5 * n2 - 1) * Vf;
    TT(row, 12) = r15 / 2.0 * n * (l2 - m2) * Vf;
    TT(row, 13) = r15 * (l * m * n) * Vf;
    TT(row, 14) = sqrt(5.0 / 8.0) * l * (l2 - 3 * m2) * Vf;
    TT(row, 15) = sqrt(5.0 / 8.0) * m * (3 * l2 - m2) * Vf;
    for (int j = 1; j < 4; ++j) TT(j, row) = -TT(row, j);    // p odd
    for (int j = 4; j < 9; ++j) TT(j, row) = TT(row, j);     // d even
    for (int j = 9; j < 16; ++j) TT(j, row) = -TT(row, j);   // f odd
#undef TT
}

template <class R>
__device__ inline void sk_fill(const R* __restrict__ V, R l, R m, R n, R* __restrict__ T) {
#define TT(i, j) T[(i) * 17 + (j)]
    for (int i = 0; i < 289; ++i) T[i] = R(0.0);
    const R l2 = l * l, m2 = m * m, n2 = n * n;
    const R lm = l * m, mn = m * n, nl = n * l, lmn = l * m * n;
    const R dm = l2 - m2, sp = l2 + m2;
    const bool off = sp > 1e-10;                 // the reference's on-axis guard
    const R q = off ? sp : 1.0;
    const R r2 = sqrt(2.0), r3 = sqrt(3.0), r6 = sqrt(6.0), r10 = sqrt(10.0), r30 = sqrt(30.0);
    const R a = 5 * n2 - 1, b = 5 * n2 - 3;
    const R x3 = l2 - 3 * m2, y3 = 3 * l2 - m2;
    const R z2 = n2 - 0.5 * sp;

    // s and S rows/cols
    TT(0, 0) = V[V_sss];
    sk_s_row(T, 0, V[V_sps], V[V_sds], V[V_sfs], l, m, n);
    TT(16, 16) = V[V_SSs];
    sk_s_row(T, 16, V[V_Sps], V[V_Sds], V[V_Sfs], l, m, n);
    TT(0, 16) = V[V_sSs];
    TT(16, 0) = V[V_sSs];

    {   // p-p
        const R s = V[V_pps], p = V[V_ppp];
        TT(1, 1) = l2 * s + (1.0 - l2) * p;
        TT(2, 2) = m2 * s + (1.0 - m2) * p;
        TT(3, 3) = n2 * s + (1.0 - n2) * p;
        TT(1, 2) = TT(2, 1) = lm * (s - p);
        TT(1, 3) = TT(3, 1) = nl * (s - p);
        TT(2, 3) = TT(3, 2) = mn * (s - p);
    }
    {   // p-d
        const R s = V[V_pds], p = V[V_pdp];
        const R cross = lmn * (r3 * s - 2 * p);
        TT(1, 4) = r3 * l2 * m * s + m * (1.0 - 2 * l2) * p;
        TT(1, 5) = cross;
        TT(1, 6) = r3 * l2 * n * s + n * (1.0 - 2 * l2) * p;
        TT(1, 7) = 0.5 * r3 * l * dm * s + l * (1.0 - l2 + m2) * p;
        TT(1, 8) = l * z2 * s - r3 * l * n2 * p;
        TT(2, 4) = r3 * m2 * l * s + l * (1.0 - 2 * m2) * p;
        TT(2, 5) = r3 * m2 * n * s + n * (1.0 - 2 * m2) * p;
        TT(2, 6) = cross;
        TT(2, 7) = 0.5 * r3 * m * dm * s - m * (1.0 + l2 - m2) * p;
        TT(2, 8) = m * z2 * s - r3 * m * n2 * p;
        TT(3, 4) = cross;
        TT(3, 5) = r3 * n2 * m * s + m * (1.0 - 2 * n2) * p;
        TT(3, 6) = r3 * n2 * l * s + l * (1.0 - 2 * n2) * p;
        TT(3, 7) = 0.5 * r3 * n * dm * s - n * dm * p;
        TT(3, 8) = n * z2 * s + r3 * n * sp * p;
        for (int i = 1; i < 4; ++i)
            for (int j = 4; j < 9; ++j) TT(j, i) = -TT(i, j);
    }
    {   // p-f : the assignments that survive in the reference
        const R s = V[V_pfs], p = V[V_pfp];
        TT(1, 9) = r6 / 4.0 * l * a * s - r6 / 2.0 * l * n2 * p;
        TT(1, 10) = (3.0 / 4.0 * l2 * a - a / 2.0) * s + (l2 * (1 - 5 * n2 / 2.0) + n2) * p;
        TT(1, 11) = 3.0 / 4.0 * lm * a * s - 5.0 / 2.0 * lm * n2 * p;
        TT(1, 12) = r10 / 4.0 * nl * dm * s - r10 / 2.0 * nl * (1 + dm / 2.0) * p;
        TT(1, 13) = r10 / 2.0 * mn * l2 * s - r10 / 2.0 * mn * (1 - l2) * p;
        TT(1, 14) = off ? (r30 / 8.0 * l * x3 * s - r30 / 4.0 * l * x3 * (1 - l2 + m2) / q * p) : 0.0;
        TT(1, 15) = off ? (r10 / 8.0 * m * y3 * s + r10 / 4.0 * m * (1 - y3 * (1 + l2 - m2) / q) * p) : 0.0;
        TT(2, 9) = r6 / 4.0 * m * a * s - r6 / 2.0 * m * n2 * p;
        TT(2, 10) = 3.0 / 4.0 * lm * a * s - 5.0 / 2.0 * lm * n2 * p;
        TT(2, 11) = (3.0 / 4.0 * m2 * a - a / 2.0) * s + (m2 * (1 - 5 * n2 / 2.0) + n2) * p;
        TT(2, 12) = r10 / 4.0 * mn * dm * s + r10 / 2.0 * mn * (1 - dm / 2.0) * p;
        TT(2, 13) = r10 / 2.0 * nl * m2 * s - r10 / 2.0 * nl * (1 - m2) * p;
        TT(2, 14) = off ? (r30 / 8.0 * m * x3 * s + r30 / 4.0 * m * x3 * (1 + l2 - m2) / q * p) : 0.0;
        TT(2, 15) = off ? (r10 / 8.0 * l * y3 * s - r10 / 4.0 * l * (1 + y3 * (1 - l2 + m2) / q) * p) : 0.0;
        TT(3, 9) = (n2 * b + 3.0 / 2.0 * (1 - n2)) * s + r6 * n2 * (1 - n2) * p;
        TT(3, 10) = r6 / 4.0 * nl * b * s + r6 / 2.0 * nl * (1 - 5 * n2 + 4 * n2) * p;
        TT(3, 11) = r6 / 4.0 * mn * b * s + r6 / 2.0 * mn * (1 - 5 * n2 + 4 * n2) * p;
        TT(3, 12) = r10 / 4.0 * n2 * dm * s + r10 / 4.0 * dm * (1 - 2 * n2) * p;
        TT(3, 13) = r10 / 2.0 * lm * n2 * s - r10 / 2.0 * lm * (1 - n2) * p;
        TT(3, 14) = r30 / 8.0 * nl * x3 * s + r30 / 4.0 * nl * (1 - x3 / 2.0) * p;
        TT(3, 15) = r10 / 8.0 * mn * y3 * s + r10 / 4.0 * mn * (1 - y3 / 2.0) * p;
        for (int i = 1; i < 4; ++i)
            for (int j = 9; j < 16; ++j) TT(j, i) = TT(i, j);      // same sign, as the reference does
    }
    {   // d-d
        const R s = V[V_dds], p = V[V_ddp], d = V[V_ddd];
        const R mix = 3 * s - 4 * p + d, pd = p - d;
        TT(4, 4) = l2 * m2 * mix + (l2 + m2) * p + n2 * d;
        TT(5, 5) = m2 * n2 * mix + (m2 + n2) * p + l2 * d;
        TT(6, 6) = n2 * l2 * mix + (n2 + l2) * p + m2 * d;
        TT(4, 5) = TT(5, 4) = l * m2 * n * mix + nl * pd;
        TT(4, 6) = TT(6, 4) = n * l2 * m * mix + mn * pd;
        TT(5, 6) = TT(6, 5) = m * n2 * l * mix + lm * pd;
        TT(4, 7) = TT(7, 4) = 0.5 * lm * dm * mix;
        TT(5, 7) = TT(7, 5) = 0.5 * mn * (dm * mix - 2 * pd);
        TT(6, 7) = TT(7, 6) = 0.5 * nl * (dm * mix + 2 * pd);
        TT(4, 8) = TT(8, 4) = r3 * (lm * z2 * s - 2 * lm * n2 * p + 0.5 * lm * (1.0 + n2) * d);
        TT(5, 8) = TT(8, 5) = r3 * (mn * z2 * s + mn * (sp - n2) * p - 0.5 * mn * sp * d);
        TT(6, 8) = TT(8, 6) = r3 * (nl * z2 * s + nl * (sp - n2) * p - 0.5 * nl * sp * d);
        TT(7, 7) = 0.25 * (dm * dm) * mix + sp * p + n2 * d;
        TT(8, 8) = 0.75 * (sp * sp) * d + 3 * sp * n2 * p + 0.25 * ((sp - 2 * n2) * (sp - 2 * n2)) * s;
        TT(7, 8) = TT(8, 7) = r3 * 0.25 * dm * (n2 * (2 * s - 4 * p + d) + d - sp * s);
    }
    {   // d-f
        const R s = V[V_dfs], p = V[V_dfp];
        const R q2 = off ? sp * sp : 1.0;
        const R w = a * (3 * n2 - 1) - 6 * n2;
        TT(4, 9) = r10 / 4.0 * lm * a * s - r10 / 2.0 * lm * n2 * p;
        TT(4, 10) = r30 / 8.0 * m * (l2 * a - a / 3.0) * s + r30 / 4.0 * m * (n2 - l2 * a / 3.0) * p;
        TT(4, 11) = r30 / 8.0 * l * (m2 * a - a / 3.0) * s + r30 / 4.0 * l * (n2 - m2 * a / 3.0) * p;
        TT(4, 12) = r2 / 2.0 * lm * n * dm * s + r2 * lm * n * (1 - dm / 2.0) * p;
        TT(4, 13) = r2 * n * l2 * m2 * s - r2 * n * (l2 + m2 - 2 * l2 * m2) * p;
        TT(4, 14) = off ? (r6 / 8.0 * m * (l2 * x3 / sp) * s - r6 / 4.0 * m * l2 * p) : 0.0;
        TT(4, 15) = off ? (r6 / 8.0 * l * (m2 * y3 / sp) * s - r6 / 4.0 * l * m2 * p) : 0.0;
        TT(5, 9) = r10 / 4.0 * mn * b * s + r10 / 2.0 * mn * (1 - n2) * p;
        TT(5, 10) = r30 / 8.0 * lmn * a * s - r30 / 4.0 * lmn * (1 + n2) * p;
        TT(5, 11) = r30 / 8.0 * (m2 * n * a - n * a / 3.0) * s + r30 / 4.0 * n * (m2 - a * m2 / 3.0) * p;
        TT(5, 12) = r2 / 2.0 * mn * n * dm * s - r2 * mn * (l2 - m2 / 2.0) * p;
        TT(5, 13) = r2 * l * m2 * n2 * s - r2 * l * (m2 + n2 - 2 * m2 * n2) * p;
        TT(5, 14) = r6 / 8.0 * mn * x3 * s - r6 / 4.0 * mn * dm * p;
        TT(5, 15) = r6 / 8.0 * nl * y3 * s - r6 / 4.0 * nl * dm * p;
        TT(6, 9) = r10 / 4.0 * nl * b * s + r10 / 2.0 * nl * (1 - n2) * p;
        TT(6, 10) = r30 / 8.0 * (l2 * n * a - n * a / 3.0) * s + r30 / 4.0 * n * (l2 - a * l2 / 3.0) * p;
        TT(6, 11) = r30 / 8.0 * lmn * a * s - r30 / 4.0 * lmn * (1 + n2) * p;
        TT(6, 12) = r2 / 2.0 * nl * n * dm * s + r2 * nl * (m2 - l2 / 2.0) * p;
        TT(6, 13) = r2 * m * l2 * n2 * s - r2 * m * (l2 + n2 - 2 * l2 * n2) * p;
        TT(6, 14) = r6 / 8.0 * nl * x3 * s + r6 / 4.0 * nl * dm * p;
        TT(6, 15) = r6 / 8.0 * mn * y3 * s + r6 / 4.0 * mn * dm * p;
        TT(7, 9) = r10 / 8.0 * n * dm * a * s - r10 / 4.0 * n * dm * n2 * p;
        TT(7, 10) = r30 / 16.0 * l * dm * a * s - r30 / 8.0 * l * (n2 * dm - l2 + m2) * p;
        TT(7, 11) = r30 / 16.0 * m * dm * a * s + r30 / 8.0 * m * (n2 * dm + l2 - m2) * p;
        TT(7, 12) = off ? (r2 / 4.0 * n * (dm * dm) * s + r2 / 2.0 * n * sp * (1 - (dm * dm) / q2) * p) : 0.0;
        TT(7, 13) = r2 / 2.0 * lm * n * dm * s - r2 * lm * n * p;
        TT(7, 14) = off ? (r6 / 16.0 * l * dm * x3 / q * s) : 0.0;
        TT(7, 15) = off ? (r6 / 16.0 * m * dm * y3 / q * s) : 0.0;
        TT(8, 9) = (n2 * b / 2.0 + 3.0 / 4.0 * (1 - n2) * a) * s + r6 / 2.0 * n2 * (1 - n2) * p;
        TT(8, 10) = r2 / 8.0 * l * w * s + r3 / 2.0 * l * n2 * (1 - n2) * p;
        TT(8, 11) = r2 / 8.0 * m * w * s + r3 / 2.0 * m * n2 * (1 - n2) * p;
        TT(8, 12) = r30 / 8.0 * n * dm * z2 * s + r30 / 4.0 * n * dm * (1 - n2) * p;
        TT(8, 13) = r30 / 4.0 * lm * n * z2 * s - r30 / 2.0 * lm * n * (1 - n2) * p;
        TT(8, 14) = r10 / 16.0 * l * x3 * (3 * n2 - 1) * s - r10 / 8.0 * l * x3 * n2 * p;
        TT(8, 15) = r10 / 16.0 * m * y3 * (3 * n2 - 1) * s - r10 / 8.0 * m * y3 * n2 * p;
        for (int i = 4; i < 9; ++i)
            for (int j = 9; j < 16; ++j) TT(j, i) = -TT(i, j);
    }
    {   // f-f
        const R s = V[V_ffs], p = V[V_ffp], d = V[V_ffd];
        const R h = 1 - 5 * n2 / 2.0;
        const R s3 = l2 * m2 + m2 * n2 + n2 * l2;
        TT(9, 9) = n2 * (b * b) / 4.0 * s + 3.0 / 2.0 * n2 * (1 - n2) * a * p + 15.0 / 4.0 * ((1 - n2) * (1 - n2)) * d;
        // the l2*h*h / m2*h*h term carries no V factor in the reference (_params.py:397,:402)
        TT(10, 10) = 3.0 / 8.0 * l2 * (a * a) * s + l2 * (h * h) + n2 * (1 - l2) * p + (1 - l2) * (1 - n2) * d;
        TT(11, 11) = 3.0 / 8.0 * m2 * (a * a) * s + m2 * (h * h) + n2 * (1 - m2) * p + (1 - m2) * (1 - n2) * d;
        // `A + B if c else 0 + C` parses as (A+B) if c else (0+C)  
        TT(12, 12) = off ? (15.0 / 4.0 * n2 * (dm * dm) * s + ((sp - n2 * dm) * (sp - n2 * dm)) / sp * p) : (0 + l2 * m2 * d);
        TT(13, 13) = 15.0 * l2 * m2 * n2 * s + (s3 - 4 * l2 * m2 * n2) * p + (s3 - l2 * m2 * n2) * d;
        TT(14, 14) = off ? (5.0 / 8.0 * l2 * (x3 * x3) / (sp * sp) * s) : 0.0;
        TT(15, 15) = off ? (5.0 / 8.0 * m2 * (y3 * y3) / (sp * sp) * s) : 0.0;
        TT(9, 10) = r6 / 8.0 * nl * a * b * s + r6 / 4.0 * nl * a * (1 - n2) * p;
        TT(9, 11) = r6 / 8.0 * mn * a * b * s + r6 / 4.0 * mn * a * (1 - n2) * p;
        TT(9, 12) = r10 / 8.0 * n2 * dm * b * s + r10 / 4.0 * dm * (1 - n2) * a * p;
        TT(9, 13) = r10 / 4.0 * lm * n2 * b * s + r10 / 2.0 * lm * (1 - n2) * a * p;
        TT(9, 14) = r30 / 16.0 * nl * x3 * b * s;
        TT(9, 15) = r30 / 16.0 * mn * y3 * b * s;
        TT(10, 11) = 3.0 / 8.0 * lm * (a * a) * s + lm * (h * h) * p;
        TT(10, 12) = r10 / 16.0 * l * dm * a * s;
        TT(10, 13) = r10 / 8.0 * m * n * l * a * s;
        TT(11, 12) = r10 / 16.0 * m * dm * a * s;
        TT(11, 13) = r10 / 8.0 * l * n * m * a * s;
        TT(12, 13) = r6 / 4.0 * lm * n * dm * s;
        TT(13, 14) = r6 / 8.0 * m * n * x3 * s;
        TT(13, 15) = r6 / 8.0 * l * n * y3 * s;
        TT(14, 15) = off ? (5.0 / 8.0 * lm * x3 * y3 / (sp * sp) * s) : 0.0;
        for (int i = 9; i < 16; ++i)
            for (int j = i + 1; j < 16; ++j) TT(j, i) = TT(i, j);
    }
#undef TT
}

// one thread per bond; V [nb][25], lmn [nb][3], out [nb][17*17]
__global__ void sk_table_kernel(const double* __restrict__ V, const double* __restrict__ lmn, double* __restrict__ out,
                                int nb) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    sk_fill<double>(V + (size_t)b * V_COUNT, lmn[3 * b], lmn[3 * b + 1], lmn[3 * b + 2], out + (size_t)b * 289);
}

// Table and its Cartesian gradient for the forces: V [nb][25] values, dV [nb][25] derivatives dV/dd at the bond length,
// vec [nb][3] Cartesian bond vectors; out [nb][4][289]: T | dT/dx | dT/dy | dT/dz.  work [nb][289] SkDual scratch.
__global__ void sk_table_grad_kernel(const double* __restrict__ V, const double* __restrict__ dV, const double* __restrict__ vec,
                                     double* __restrict__ out, SkDual* __restrict__ work, int nb) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const double x = vec[3 * b], y = vec[3 * b + 1], z = vec[3 * b + 2];
    const double d = sqrt(x * x + y * y + z * z), id = 1.0 / d;
    const double c[3] = {x * id, y * id, z * id};
    SkDual lmn[3];
    for (int a = 0; a < 3; ++a) lmn[a] = SkDual(c[a], ((a == 0) - c[a] * c[0]) * id, ((a == 1) - c[a] * c[1]) * id, ((a == 2) - c[a] * c[2]) * id);
    SkDual Vd[V_COUNT];
    for (int q = 0; q < V_COUNT; ++q) {
        const double dv = dV[(size_t)b * V_COUNT + q];
        Vd[q] = SkDual(V[(size_t)b * V_COUNT + q], dv * c[0], dv * c[1], dv * c[2]);
    }
    SkDual* T = work + (size_t)b * 289;
    sk_fill<SkDual>(Vd, lmn[0], lmn[1], lmn[2], T);
    double* o = out + (size_t)b * 4 * 289;
    for (int e = 0; e < 289; ++e) { o[e] = T[e].v; o[289 + e] = T[e].g[0]; o[578 + e] = T[e].g[1]; o[867 + e] = T[e].g[2]; }
}

}  // namespace sktb
