// Shared device helpers for libsktb (sm_100a).  FP64 arithmetic throughout.  The batched Hermitian eigenproblem of tight binding is not
// GEMM-shaped for N <= ~200 (register-resident DFMA kernels); from N = 208 the two-stage reduction (heev_twostage.cuh) puts its BLAS-3
// part on the FP64 tensor instruction (DMMA), which on B200 shares the DFMA pipe but reaches its full rate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sktb {

typedef double2 cplx;

#define SKTB_HD __host__ __device__ __forceinline__
#define SKTB_D __device__ __forceinline__

SKTB_HD cplx cmake(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
SKTB_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
SKTB_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
SKTB_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
SKTB_HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SKTB_HD cplx cmulc(cplx a, cplx b) { return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }  // conj(a)*b
SKTB_HD cplx cscale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
#ifdef __CUDACC__
// 1/sqrt(a): hardware seed (~2^-22) and one cubic correction step (full double precision), no slow path
SKTB_D double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a * y, y, 1.0);
    const double c = fma(e, 0.375, 0.5);
    return fma(c, y * e, y);
}
// 1/a: hardware seed (rcp.approx.ftz.f64 looks at the upper 32 bits of a: relative error e <= ~2^-19) and ONE cubic step
// r (1 + e + e^2): error e^3 < 2^-56, i.e. below the rounding of the last fma; 3 FP64 operations instead of fast_rcp's 5
SKTB_D double fast_rcp3(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}
// 1/a: hardware seed and two Newton steps (full double precision for normal a), no slow path
SKTB_D double fast_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
#endif
// acc += a*b
SKTB_HD void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
SKTB_HD void cfmac(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= a*b
SKTB_HD void cfms(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= conj(a)*b
SKTB_HD void cfmsc(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// ---------------------------------------------------------------------------------------------------------
// Householder generator, LAPACK zlarfg semantics without the safmin rescaling loop (eV-scale inputs).
// In : alpha = x[0], xnorm2 = sum_{i>=1} |x_i|^2.   Out: beta (real, new x[0]), tau, scale (x[1:] *= scale).
// Returns false when H = I (tau = 0): then beta = Re(alpha).
// ---------------------------------------------------------------------------------------------------------
SKTB_HD bool householder(cplx alpha, double xnorm2, double& beta, cplx& tau, cplx& scale) {
    if (xnorm2 == 0.0 && alpha.y == 0.0) {
        beta = alpha.x; tau = cmake(0.0, 0.0); scale = cmake(0.0, 0.0);
        return false;
    }
    double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
    beta = (alpha.x >= 0.0) ? -nrm : nrm;   // -sign(nrm, Re alpha)
    double ib = 1.0 / beta;
    tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
    double dx = alpha.x - beta, dy = alpha.y;
    double id = 1.0 / (dx * dx + dy * dy);
    scale = cmake(dx * id, -dy * id);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Implicit-shift QL on a real symmetric tridiagonal (d[0..n), e[0..n-1) couples i,i+1), eigenvalues only,
// strided storage (element i at base[i*stride]) so that a warp can keep one matrix per lane with
// conflict-free shared memory.  e must have room for n entries; destroyed.  Returns iterations > limit ? 1 : 0.
// ---------------------------------------------------------------------------------------------------------
SKTB_HD int tql_values(double* d, double* e, int n, int stride) {
    const double eps = 2.220446049250313e-16;
#define D_(i) d[(i) * stride]
#define E_(i) e[(i) * stride]
    if (n > 0) E_(n - 1) = 0.0;
    int fail = 0;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < n - 1; ++m) {
                double dd = fabs(D_(m)) + fabs(D_(m + 1));
                if (fabs(E_(m)) <= eps * dd) break;
            }
            if (m == l) break;
            if (++iter > 80) { fail = 1; break; }
            double el = E_(l);
            double g = (D_(l + 1) - D_(l)) / (2.0 * el);
            double r = sqrt(g * g + 1.0);
            g = D_(m) - D_(l) + el / (g + (g >= 0.0 ? r : -r));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            bool under = false;
            for (; i >= l; --i) {
                double ei = E_(i);
                double f = s * ei, b = c * ei;
                r = sqrt(f * f + g * g);
                E_(i + 1) = r;
                if (r == 0.0) { D_(i + 1) -= p; E_(m) = 0.0; under = true; break; }
                double ir = 1.0 / r;
                s = f * ir; c = g * ir;
                g = D_(i + 1) - p;
                r = (D_(i) - g) * s + 2.0 * c * b;
                p = s * r;
                D_(i + 1) = g + p;
                g = c * r - b;
            }
            if (under) continue;
            D_(l) -= p; E_(l) = g; E_(m) = 0.0;
        }
    }
#undef D_
#undef E_
    return fail;
}

// block-wide sums (all threads get the result). scratch: >= 64 doubles of shared memory.
SKTB_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SKTB_D double block_sum(double v, double* scratch) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect scratch from the previous use
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double t = (lane < nw) ? scratch[lane] : 0.0;
    return warp_sum(t);
}
SKTB_D cplx block_sum(cplx v, double* scratch) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v.x = warp_sum(v.x); v.y = warp_sum(v.y);
    __syncthreads();
    if (lane == 0) { scratch[wid] = v.x; scratch[32 + wid] = v.y; }
    __syncthreads();
    double tx = (lane < nw) ? scratch[lane] : 0.0, ty = (lane < nw) ? scratch[32 + lane] : 0.0;
    return cmake(warp_sum(tx), warp_sum(ty));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------
// Sturm count (number of eigenvalues < x) of the symmetric tridiagonal block rows i0..i1 in PRODUCT form:
// p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}, 3 FP64 operations per row (the quotient form of dstebz costs a division).
// de2[i] = (d_i, e_{i-1}^2), scaled so that |d|, |e| <= 1 (then |p| grows by < 2^2.2 per row): the pair (p_{i-1}, p_i) is renormalised
// by an exact power of two every 8 rows when its exponent leaves +-300; an exact zero becomes -2^-100 p_{i-1} (a sign change, dstebz's
// pivmin rule) - blocks of 8 rows run without the zero test on the FP64 pipe and are redone carefully when an integer test saw a zero.
// ---------------------------------------------------------------------------------------------------------
SKTB_D int sturm_count_prod(const double2* __restrict__ de2, int i0, int i1, double x) {
    const double tiny = -7.888609052210118e-31;                   // -2^-100
    double p0 = 1.0, p1 = de2[i0].x - x;
    if (p1 == 0.0) p1 = tiny;
    int cnt = p1 < 0.0;
    int r = i0 + 1;
#define SKTB_STURM_STEP(RR)                                                                         \
    {                                                                                               \
        const double2 v = de2[RR];                                                                  \
        double pn = fma(v.x - x, p1, -(v.y * p0));                                                  \
        if (pn == 0.0) pn = tiny * p1;                                                              \
        cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p1)) >> 31;                           \
        p0 = p1; p1 = pn;                                                                           \
    }
#define SKTB_STURM_FAST(RR)                                                                         \
    {                                                                                               \
        const double2 v = de2[RR];                                                                  \
        const double pn = fma(v.x - x, p1, -(v.y * p0));                                            \
        const int hn = __double2hiint(pn);                                                          \
        cnt += (unsigned)(hn ^ __double2hiint(p1)) >> 31;                                           \
        nz = min(nz, (unsigned)(hn << 1) | (unsigned)__double2loint(pn));                           \
        p0 = p1; p1 = pn;                                                                           \
    }
    for (; r + 8 <= i1 + 1; r += 8) {
        const double s0 = p0, s1 = p1; const int sc0 = cnt;
        unsigned nz = 0xffffffffu;
        SKTB_STURM_FAST(r) SKTB_STURM_FAST(r + 1) SKTB_STURM_FAST(r + 2) SKTB_STURM_FAST(r + 3)
        SKTB_STURM_FAST(r + 4) SKTB_STURM_FAST(r + 5) SKTB_STURM_FAST(r + 6) SKTB_STURM_FAST(r + 7)
        if (nz == 0u) {
            p0 = s0; p1 = s1; cnt = sc0;
            for (int q = 0; q < 8; ++q) SKTB_STURM_STEP(r + q)
        }
        const int exq = ((__double2hiint(p1) >> 20) & 0x7ff) - 1023;
        if (exq > 300 || exq < -300) {
            const double rs = __hiloint2double((1023 - max(exq, -1000)) << 20, 0);
            p0 *= rs; p1 *= rs;
        }
    }
    for (; r <= i1; ++r) SKTB_STURM_STEP(r)
#undef SKTB_STURM_STEP
#undef SKTB_STURM_FAST
    return cnt;
}
#endif

// k-mesh point of flat index idx (greens.py:104-109): true IEEE divisions, no reciprocal multiply.
SKTB_HD void kmesh_point(long long idx, int n0, int n1, int n2, double& kx, double& ky, double& kz) {
    long long l, j, i;
    if ((unsigned long long)idx <= 0xffffffffULL) {          // 32-bit index arithmetic whenever the flat index fits (64-bit
        const unsigned u = (unsigned)idx;                     // integer division costs ~100 instructions per k-point)
        const unsigned t32 = u / (unsigned)n2;
        l = u - t32 * (unsigned)n2;
        const unsigned i32 = t32 / (unsigned)n1;
        j = t32 - i32 * (unsigned)n1;
        i = i32;
    } else {
        l = idx % n2;
        const long long t = idx / n2;
        j = t % n1; i = t / n1;
    }
    double d0 = (double)(n0 > 1 ? n0 : 1), d1 = (double)(n1 > 1 ? n1 : 1), d2 = (double)(n2 > 1 ? n2 : 1);
#ifdef __CUDA_ARCH__
    kx = __ddiv_rn((double)i, d0); ky = __ddiv_rn((double)j, d1); kz = __ddiv_rn((double)l, d2);
#else
    kx = (double)i / d0; ky = (double)j / d1; kz = (double)l / d2;
#endif
}

// Device view of the k-independent model tables (see sktb_model_desc in include/sktb.h).
struct PlanView {
    int n_atoms, n_orb, n_bonds, n_pairs, soc_nnz;
    const int* atom_start;    // [n_atoms+1]
    const int* orb_atom;      // [n_orb]
    const int* pair_of;       // [n_atoms*n_atoms] -> pair id or -1
    const int* pair_begin;    // [n_pairs+1] range into the pair-sorted bond arrays
    const double* bond_vec;   // [n_bonds][3]  (pair-sorted, image ascending inside a pair)
    const long long* bond_off;  // [n_bonds] offset of the bond's [n_i][n_j] block in T
    const double* T;          // packed hopping blocks
    const double* onsite;     // [n_orb]
    const int* soc_row; const int* soc_col; const cplx* soc_val;
    double rec[9];            // inv(lattice).T row-major
};

#ifdef __CUDACC__
// phase of one bond at fractional k: exp(2 pi i (k @ rec) . d), same operation order as calc_g
// (hamiltonian.py:280, :311): kc = k@rec ; arg = (2*pi) * dot(kc, d).
SKTB_D cplx bond_phase(const double* rec, double k0, double k1, double k2, const double* d) {
    double c0 = __dadd_rn(__dadd_rn(__dmul_rn(k0, rec[0]), __dmul_rn(k1, rec[3])), __dmul_rn(k2, rec[6]));
    double c1 = __dadd_rn(__dadd_rn(__dmul_rn(k0, rec[1]), __dmul_rn(k1, rec[4])), __dmul_rn(k2, rec[7]));
    double c2 = __dadd_rn(__dadd_rn(__dmul_rn(k0, rec[2]), __dmul_rn(k1, rec[5])), __dmul_rn(k2, rec[8]));
    double dot = __dadd_rn(__dadd_rn(__dmul_rn(c0, d[0]), __dmul_rn(c1, d[1])), __dmul_rn(c2, d[2]));
    double arg = __dmul_rn(6.283185307179586, dot);
    cplx ph;
    sincos(arg, &ph.y, &ph.x);
    return ph;
}
#endif

}  // namespace sktb
