// K0+K1+K2 for N <= 32 (configs C1-C3 and the north-star 32-orbital model), register-resident.
// H(k) is assembled straight into registers and never touches HBM; per k-point the HBM traffic is 24 B of k in
// (0 B for meshes: K0 is fused), a 16*NC-byte (d,e) scratch round trip through L2, and 8*N B of eigenvalues out.
//
// Two kernels per chunk of k-points:
//
//  tridiag_small_kernel  (FP64-pipe bound)   L = 4/8/16/32 lanes own one matrix (32/L matrices per warp side by
//    side); lane i holds row i of the mirrored-lower Hermitian matrix
//        A_eff[i][k] = H[i][k] (k<i), conj(H[k][i]) (k>i), Re H[i][i]
//    i.e. LAPACK zheevd(UPLO='L') semantics, which is what the reference's numpy call reads (hamiltonian.py:119).
//    The mirrored tables are prepared on the host (build_tables) so no cross-lane transposition is needed:
//        Lt[b][kappa][iota] = T_b[iota][kappa] (kappa <= iota) | T_b[kappa][iota] (kappa > iota, conj phase).
//    Householder tridiagonalisation keeps register indexing static WITHOUT unrolling the column loop: after
//    every step the register window is rotated by one column (column j always sits in register 0), and the
//    loop is split into phases of S steps whose bodies are compiled for a fixed window width W = NC, NC-S, ...
//    (a fully unrolled N=32 body is ~600 KB of SASS and made the kernel instruction-fetch bound at 3 % of FP64
//    peak; the rolled phases are ~40 KB and stay in the instruction cache).  v and w are exchanged through a
//    2 KB shared-memory buffer per warp (broadcast LDS.128), norms and dots by xor-shuffles inside the group.
//
//  tql_small_kernel      (latency bound, high occupancy)   one tridiagonal matrix PER LANE: (d,e) are transposed
//    through shared memory (odd stride, conflict-free), every lane runs the implicit-shift QL of its own matrix
//    (no idle lanes, rsqrt-based rotations), sorts its eigenvalues with a register bitonic network and the warp
//    stores them coalesced along k (the reference's [band, k] layout).
//
// FP64 CUDA cores only: the work is batched rank-2 updates on <=32x32 Hermitian matrices, not a GEMM.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/sktb.h"
#include "common.cuh"

namespace sktb {
namespace small {

struct SmallTables {
    int n_orb = 0, n_bonds = 0;
    const double* Lt = nullptr;     // [nb][n][n]   mirrored-lower hoppings, lane index fastest
    const double* Dt = nullptr;     // [nb][n][n]   T_b[iota][kappa] - T_b[kappa][iota]  (Hermiticity gate)
    const double* bvec = nullptr;   // [nb][3]
    const double* onsite = nullptr; // [n]
    const cplx* St = nullptr;       // [2n][2n]     mirrored-lower soc_mat, St[k*2n + i]
    const double* SAt = nullptr;    // [2n][2n]     Re S[i][k] - Re S[k][i], SAt[k*2n + i]
    const int* thread_order = nullptr;   // [2][nb] slot order + conj flags of the one-matrix-per-thread kernel (heev_thread.cuh)
    double rec[9];
    double asym_bound = 0.0;
    int soc_any = 0;
};

struct SmallArgs {
    SmallTables t;
    const double* k;            // [nk][3] or NULL -> mesh
    int mesh0, mesh1, mesh2;
    long long first, nk;        // this launch: mesh points first.. / k rows 0..nk
    int N, soc;
    double* scratch;            // [nk][2][NC]  (d | e) of every matrix
    int* flags; long long flag_off;
    const cplx* Hsrc = nullptr; // pair kernels in LOAD mode (sktb_heev): [nk][N][N] row-major matrices instead of K0+K1
    int Hld = 0;                // LOAD mode: leading dimension / matrix stride of Hsrc (0 -> N); tridiag_cta64's 32 x 32 blocks use 32
    int out_stride = 0, d_off = 0, e_off = 0;   // pair kernels: (d,e) record layout, 0 -> [2][32] per k; the cta64 split writes into [2][64] at 32
    int v_per_k = 0, v_stride = 0, v_off = 0, v_tau = 0;   // reflector store layout (eigenvector path), 0 -> [32][32] + tau at 1024;
                                                            // the cta64 split stores rows / reflectors 32.. of a [64][64] + tau at 4096 record
};

struct TqlArgs {
    const double* scratch; long long nk; int N;
    double* evals; long long ld, col0;
    int* fail;
};

inline bool supports(int N) { return N >= 1 && N <= 32; }

constexpr int SMALL_WARPS = 4;
#ifndef SKTB_RCP3
#define SKTB_RCP3 1            // 1: reciprocals of the PWK sweep / of zlarfg by one cubic step on the hardware seed (3 FP64 ops instead of 5;
#endif                         //    measured +1.6 % on the T step, eigenvalues unchanged at the 1e-14 level); 0: two Newton steps
#ifndef SKTB_SMALL_SYNC
#define SKTB_SMALL_SYNC 1
#endif
#ifndef SKTB_FOLD_S
#define SKTB_FOLD_S 4
#endif
constexpr int DSTRIDE = 33;   // odd stride: conflict-free both for the transposing stores and the per-lane QL

// ---------------------------------------------------------------------------------------------------------
template <int P2>
SKTB_D void bitonic_sort(double (&a)[P2]) {
#pragma unroll
    for (int k = 2; k <= P2; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < P2; ++i) {
                int l = i ^ j;
                if (l > i) {
                    bool up = (i & k) == 0;
                    double lo = fmin(a[i], a[l]), hi = fmax(a[i], a[l]);
                    a[i] = up ? lo : hi;
                    a[l] = up ? hi : lo;
                }
            }
        }
    }
}

// implicit-shift QL, one matrix per lane, strided shared memory (stride DSTRIDE doubles), rsqrt rotations.
// Flat state machine: every trip of the single loop is one QL sweep of the lane's current unreduced block
// [l..m] (or a cheap bookkeeping step), so a warp's trip count is the MAX over lanes of the total number of
// sweeps instead of the sum over eigenvalues of the per-eigenvalue maximum (lanes converge at different rates).
// The deflation tests of the next sweep are folded into the rotation loop; rescans happen only at block ends.
SKTB_D int tql_scan(const double* d, const double* e, int l, int n) {
    const double eps = 2.220446049250313e-16;
    int m = l;
    double dm = fabs(d[m * DSTRIDE]);
    for (; m < n - 1; ++m) {
        double dn = fabs(d[(m + 1) * DSTRIDE]);
        if (fabs(e[m * DSTRIDE]) <= eps * (dm + dn)) break;
        dm = dn;
    }
    return m;
}

template <bool FAST = true>
SKTB_D int tql_lane(double* d, double* e, int n) {
    const double eps = 2.220446049250313e-16;
#define D_(i) d[(i) * DSTRIDE]
#define E_(i) e[(i) * DSTRIDE]
    E_(n - 1) = 0.0;
    int fail = 0, l = 0, iter = 0;
    int m = tql_scan(d, e, 0, n);
    while (true) {
        if (m == l || iter > 80) {                    // d[l] converged (or given up): next eigenvalue
            if (m != l) fail = 1;
            ++l; iter = 0;
            if (l >= n) break;
            m = tql_scan(d, e, l, n);
            continue;
        }
        ++iter;
        double el = E_(l), dl = D_(l);
        double g, r;
        if (FAST) {
            // Wilkinson shift with Newton-refined hardware reciprocals (no IEEE slow paths); a non-finite intermediate
            // (|e_l| tiny against the diagonal gap) means the shift is d[m] - d[l] to working precision
            g = (D_(l + 1) - dl) * fast_rcp(2.0 * el);
            const double h = fma(g, g, 1.0);
            r = h * fast_rsqrt(h);
            const double q = el * fast_rcp(g + (g >= 0.0 ? r : -r));
            g = D_(m) - dl + ((fabs(q) <= 1.7e308) ? q : 0.0);
        } else {
            g = (D_(l + 1) - dl) / (2.0 * el);
            r = sqrt(fma(g, g, 1.0));
            g = D_(m) - dl + el / (g + (g >= 0.0 ? r : -r));
        }
        double s = 1.0, c = 1.0, p = 0.0;
        bool under = false;
        double di1 = D_(m);                           // original d[i+1]
        double dnext = 0.0;                           // |d[i+2]| after this sweep
        int mb = m;                                   // smallest index in (l, m) whose new e is negligible
        for (int i = m - 1; i >= l; --i) {
            double ei = E_(i), di = D_(i);
            double f = s * ei, b = c * ei;
            double h2 = fma(f, f, g * g);
            if (FAST ? h2 < 1e-290 : h2 == 0.0) { E_(i + 1) = 0.0; D_(i + 1) = di1 - p; E_(m) = 0.0; under = true; break; }
            double ir = FAST ? fast_rsqrt(h2) : rsqrt(h2);
            r = h2 * ir;
            s = f * ir; c = g * ir;
            g = di1 - p;
            double rr = fma(di - g, s, 2.0 * c * b);
            p = s * rr;
            double dnew = g + p;
            D_(i + 1) = dnew;
            double an = fabs(dnew);
            if (i + 1 < m) {                          // e[i+1] couples i+1 and i+2, both final now
                E_(i + 1) = r;
                if (r <= eps * (an + dnext)) mb = i + 1;
            }
            dnext = an;
            g = fma(c, rr, -b);
            di1 = di;
        }
        if (under) { m = tql_scan(d, e, l, n); continue; }
        double dl_new = di1 - p;
        D_(l) = dl_new; E_(l) = g; E_(m) = 0.0;
        if (fabs(g) <= eps * (fabs(dl_new) + dnext)) { ++l; iter = 0; }   // front deflated: block becomes [l+1..mb]
        m = mb;
    }
#undef D_
#undef E_
    return fail;
}

// Square-root-free implicit QL (Pal-Walker-Kahan), eigenvalues only, one matrix per lane: the recurrence of LAPACK
// dsterf's QL branch - the routine numpy's eigvalsh ends in (hamiltonian.py:119 -> zheevd(jobz='N') -> dsterf) - on
// (d, e^2), as the same flat state machine as tql_lane: one trip = one sweep of the lane's unreduced block [l..m], the
// deflation tests of the next sweep folded into the sweep.  Per inner step: 2 LDS + 2 STS, two independent hardware
// reciprocals (1/r and 1/p; 1/c = r/p) with Newton refinement, ~20 FP64 ops - about half of the rotation-based sweep
// (one rsqrt, but both the (c, s) rotation and the f/b/r recurrence).  Deflation: e^2 <= eps^2 |d_m d_{m+1}| (dsterf's
// test) + a norm-relative floor (1e-3 eps ||T||)^2 so that exactly-zero diagonals (zero modes) deflate as well.
// tools/proto_pwk.py is the scalar model of this function.
SKTB_D int tql_scan_pwk(const double* d, const double* e, int l, int n, const double eps2, const double floor2) {
    int m = l;
    for (; m < n - 1; ++m)
        if (e[m * DSTRIDE] <= fma(eps2, fabs(d[m * DSTRIDE] * d[(m + 1) * DSTRIDE]), floor2)) break;
    return m;
}

SKTB_D int tql_lane_pwk(double* d, double* e, int n) {
    const double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
#define D_(i) d[(i) * DSTRIDE]
#define E_(i) e[(i) * DSTRIDE]
    double anorm = 0.0;
    for (int i = 0; i < n; ++i) {
        const double ei = (i < n - 1) ? E_(i) : 0.0;
        anorm = fmax(anorm, fmax(fabs(D_(i)), fabs(ei)));
        E_(i) = ei * ei;
    }
    const double floor2 = eps2 * anorm * anorm * 1.0e-6;
    int fail = 0, l = 0, iter = 0;
    int m = tql_scan_pwk(d, e, 0, n, eps2, floor2);
    while (true) {
        if (m == l || iter > 80) {
            if (m != l) fail = 1;
            ++l; iter = 0;
            if (l >= n) break;
            m = tql_scan_pwk(d, e, l, n, eps2, floor2);
            continue;
        }
        ++iter;
        const double p = D_(l), e2l = E_(l);
        const double rte = e2l * fast_rsqrt(e2l);
        const double sg = (D_(l + 1) - p) * fast_rcp(2.0 * rte);
        const double h = fma(sg, sg, 1.0);
        const double r0 = h * fast_rsqrt(h);
        const double q = rte * fast_rcp(sg + (sg >= 0.0 ? r0 : -r0));
        const double sigma = p - ((fabs(q) <= 1.7e308) ? q : 0.0);
        double c = 1.0, s = 0.0;
        double gamma = D_(m) - sigma;
        double pp = gamma * gamma;
        double dnext = 0.0;
        int mb = m;
        for (int i = m - 1; i >= l; --i) {
            const double bb = E_(i), alpha = D_(i);
            const double r = pp + bb;
#if SKTB_RCP3
            const double ir = fast_rcp3(r), ipp = fast_rcp3(pp);   // independent: 1/c = r / pp
#else
            const double ir = fast_rcp(r), ipp = fast_rcp(pp);     // independent: 1/c = r / pp
#endif
            const double e2new = s * r;
            const double oldc = c, oldgam = gamma;
            c = pp * ir; s = bb * ir;
            gamma = fma(c, alpha - sigma, -s * oldgam);
            const double dnew = oldgam + (alpha - gamma);
            D_(i + 1) = dnew;
            const double g2 = gamma * gamma;
            pp = (c > 1.0e-290) ? g2 * (r * ipp) : oldc * bb;
            const double adn = fabs(dnew);
            if (i != m - 1) {
                E_(i + 1) = e2new;
                // (measured and dropped: a one-DSETP filter against eps^2 anorm^2 in front of this test - the branch costs more than the
                // two FP64 operations it saves: T 3.91e7 -> 3.85e7 k/s)
                // (also measured and dropped: the same test in integer arithmetic on the upper words of the bit patterns, a sufficient
                // condition within a factor 1.2 of the threshold - no gain, 3.89e7 against 3.90e7)
                if (e2new <= fma(eps2, adn * dnext, floor2)) mb = i + 1;
            }
            dnext = adn;
        }
        const double e2n = s * pp, dl = sigma + gamma;
        E_(l) = e2n; D_(l) = dl; E_(m) = 0.0;
        if (e2n <= fma(eps2, fabs(dl) * dnext, floor2)) { ++l; iter = 0; }
        m = mb;
    }
#undef D_
#undef E_
    return fail;
}

// MODE 0: IEEE sqrt / division rotations, 1: rsqrt rotations (tql_lane<true>), 2: square-root-free PWK sweep
template <int MODE>
SKTB_D int tql_lane_mode(double* d, double* e, int n) {
    if constexpr (MODE == 2) return tql_lane_pwk(d, e, n);
    else return tql_lane<MODE == 1>(d, e, n);
}
inline int tql_mode_env() {
    static const int mode = getenv("SKTB_TQL_MODE") ? atoi(getenv("SKTB_TQL_MODE")) : ((getenv("SKTB_TQL_FAST") && getenv("SKTB_TQL_FAST")[0] == '0') ? 0 : 2);
    return mode;
}

template <int L>
SKTB_D double group_sum(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// One Householder step on a register window of W columns: B[q] = A[i][j+q].  v/w buffers hold 2L entries per
// group, the upper L are zero, so window columns beyond the matrix read v = w = 0.
// ---------------------------------------------------------------------------------------------------------
template <int NC> struct PhaseSteps { static constexpr int value = NC >= 16 ? 8 : 4; };

// reduction inputs of a pivot column: xn = sum_{i > j+1} |A[i][j]|^2, alpha = A[j+1][j]
template <int L>
SKTB_D void pivot_reduce(const cplx x, const int j, const int i, double& xn, cplx& alpha) {
    double t = (i > j + 1) ? fma(x.x, x.x, x.y * x.y) : 0.0;
    xn = group_sum<L>(t);
    alpha.x = __shfl_sync(0xffffffffu, x.x, j + 1, L);
    alpha.y = __shfl_sync(0xffffffffu, x.y, j + 1, L);
}

// A[i][k] -= v_i conj(w_k) + w_i conj(v_k)
SKTB_D void rank2(cplx& t, const cplx v, const cplx w, const cplx vk, const cplx wk) {
    t.x = fma(-v.x, wk.x, t.x); t.x = fma(-v.y, wk.y, t.x);
    t.y = fma(-v.y, wk.x, t.y); t.y = fma(v.x, wk.y, t.y);
    t.x = fma(-w.x, vk.x, t.x); t.x = fma(-w.y, vk.y, t.x);
    t.y = fma(-w.y, vk.x, t.y); t.y = fma(w.x, vk.y, t.y);
}

// Software-pipelined step: (xn, alpha) of the current pivot column arrive from the previous step; the next pivot
// column (window register 1) is updated first and its shuffle reduction + the scalar Householder chain are issued
// before the remaining W-2 column updates, which are independent of it and hide its latency.
template <int W, int NC, int L>
SKTB_D void hh_step(cplx (&B)[NC], const int j, const int i, const int g, const int lane, cplx* vbuf, cplx* wbuf, double& dval,
                    double& eval, double& xn, cplx& alpha) {
    const cplx x = B[0];
    const bool below = i > j + 1;
    dval = (i == j) ? x.x : dval;
    // Householder generator (zlarfg): one rsqrt + one reciprocal
    double beta; cplx tau, scale;
    {
        const bool trivial = (xn == 0.0) && (alpha.y == 0.0);
        double n2 = fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xn));
        double inrm = trivial ? 0.0 : rsqrt(n2);
        double nrm = n2 * inrm;
        double sgn = alpha.x >= 0.0 ? -1.0 : 1.0;
        beta = trivial ? alpha.x : sgn * nrm;
        double ib = sgn * inrm;                                   // 1 / beta
        tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
        double dx = alpha.x - beta, dy = alpha.y;
        double den = fma(dx, dx, dy * dy);
        double id = trivial ? 0.0 : __drcp_rn(den);
        scale = cmake(dx * id, -dy * id);
    }
    eval = (i == j) ? beta : eval;
    if (j == NC - 2) return;                                      // the last reflector only fixes e[NC-2]
    cplx v = (i == j + 1) ? cmake(1.0, 0.0) : (below ? cmul(x, scale) : cmake(0.0, 0.0));
    vbuf[lane + g * L] = v;                                       // group g owns [2gL, 2gL+L), zero tail above
    __syncwarp();
    const cplx* vg = vbuf + 2 * g * L + j;
    const cplx* wg = wbuf + 2 * g * L + j;
    cplx acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = cmake(0.0, 0.0);
#pragma unroll
    for (int q = 1; q < W; ++q) cfma(acc[q & 3], B[q], vg[q]);
    cplx p = cmul(tau, cadd(cadd(acc[0], acc[1]), cadd(acc[2], acc[3])));
    if (i <= j) p = cmake(0.0, 0.0);
    cplx dp = cmulc(p, v);
    dp.x = group_sum<L>(dp.x); dp.y = group_sum<L>(dp.y);
    cplx a2 = cmul(cmake(-0.5 * tau.x, -0.5 * tau.y), dp);
    cplx w = p;
    cfma(w, a2, v);
    wbuf[lane + g * L] = w;
    __syncwarp();
    if (W > 1) {
        rank2(B[1], v, w, vg[1], wg[1]);
        pivot_reduce<L>(B[1], j + 1, i, xn, alpha);               // next step's reduction, overlapped with the rest
    }
#pragma unroll
    for (int q = 2; q < W; ++q) rank2(B[q], v, w, vg[q], wg[q]);
    __syncwarp();
}

// phases W, W-S, ... down to (but not including) WEND; WEND = 0 runs the whole tridiagonalisation
template <int W, int S, int NC, int L, int WEND = 0>
SKTB_D void hh_phases(cplx (&B)[NC], int& j, const int i, const int g, const int lane, cplx* vbuf, cplx* wbuf, double& dval,
                      double& eval, double& xn, cplx& alpha) {
    constexpr int steps = (W <= S) ? W - 1 : S;
#pragma unroll 1
    for (int s = 0; s < steps; ++s, ++j) {
        hh_step<W, NC, L>(B, j, i, g, lane, vbuf, wbuf, dval, eval, xn, alpha);
#pragma unroll
        for (int q = 0; q + 1 < W; ++q) B[q] = B[q + 1];          // rotate: column j+1 -> register 0
        B[W - 1] = cmake(0.0, 0.0);
    }
    if constexpr (W > S && W - S > WEND) hh_phases<W - S, S, NC, L, WEND>(B, j, i, g, lane, vbuf, wbuf, dval, eval, xn, alpha);
}

// ---------------------------------------------------------------------------------------------------------
// NC: compile-time column count (>= N, multiple of 4); L: lanes per matrix (power of two >= NC);
// SOC: kron(h, I2) + soc_mat assembly; GATE: evaluate the reference's Hermiticity gate per k.
// ---------------------------------------------------------------------------------------------------------
struct SmemView {
    double *Lt, *Dt, *On, *Bv, *SA;
    cplx *St, *warp_base;
};

template <bool SOC, bool GATE>
SKTB_D SmemView carve_and_load(unsigned char* smem_raw, const SmallArgs& a) {
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N, tid = threadIdx.x;
    SmemView s;
    s.Lt = reinterpret_cast<double*>(smem_raw);
    s.Dt = s.Lt + (size_t)nb * n * n;
    s.On = s.Dt + (GATE ? (size_t)nb * n * n : 0);
    s.Bv = s.On + n;
    s.SA = s.Bv + 3 * nb;
    size_t off = (size_t)((s.SA + ((GATE && SOC) ? N * N : 0)) - s.Lt);
    off = (off + 1) & ~(size_t)1;              // 16-byte align
    s.St = reinterpret_cast<cplx*>(s.Lt + off);
    s.warp_base = s.St + (SOC ? N * N : 0);
    for (int t = tid; t < nb * n * n; t += blockDim.x) { s.Lt[t] = a.t.Lt[t]; if (GATE) s.Dt[t] = a.t.Dt[t]; }
    for (int t = tid; t < n; t += blockDim.x) s.On[t] = a.t.onsite[t];
    for (int t = tid; t < 3 * nb; t += blockDim.x) s.Bv[t] = a.t.bvec[t];
    if (SOC) for (int t = tid; t < N * N; t += blockDim.x) { s.St[t] = a.t.St[t]; if (GATE) s.SA[t] = a.t.SAt[t]; }
    return s;
}

// bond phases exp(2 pi i (k@rec).d) of one k-point, one sincos per lane of the group (calc_g arithmetic,
// hamiltonian.py:280,311); lanes i, i+L, ... of the group fill phs[0..nb)
template <int L>
SKTB_D void bond_phases(const SmallArgs& a, const SmemView& s, long long kq, int i, cplx* phs) {
    double k0, k1, k2;
    if (a.k) { k0 = a.k[3 * kq]; k1 = a.k[3 * kq + 1]; k2 = a.k[3 * kq + 2]; }
    else kmesh_point(a.first + kq, a.mesh0, a.mesh1, a.mesh2, k0, k1, k2);
    double c0 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[0]), __dmul_rn(k1, a.t.rec[3])), __dmul_rn(k2, a.t.rec[6]));
    double c1 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[1]), __dmul_rn(k1, a.t.rec[4])), __dmul_rn(k2, a.t.rec[7]));
    double c2 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[2]), __dmul_rn(k1, a.t.rec[5])), __dmul_rn(k2, a.t.rec[8]));
    for (int b = i; b < a.t.n_bonds; b += L) {
        double dot = __dadd_rn(__dadd_rn(__dmul_rn(c0, s.Bv[3 * b]), __dmul_rn(c1, s.Bv[3 * b + 1])), __dmul_rn(c2, s.Bv[3 * b + 2]));
        cplx ph;
        sincos(__dmul_rn(6.283185307179586, dot), &ph.y, &ph.x);
        phs[b] = ph;
    }
}

// K1: row i of the mirrored-lower H(k), straight into registers (bond loop outside: NH independent FMA chains per
// bond).  Returns max_k Re(H - H^dagger)[i][k] when GATE (else -1).
template <int NC, bool SOC, bool GATE>
SKTB_D double build_row(cplx (&A)[NC], const SmallArgs& a, const SmemView& s, const cplx* phs, const int i) {
    constexpr int NH = SOC ? NC / 2 : NC;     // orbital columns built per row
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N;
    const bool row_ok = i < N;
    const int mu = SOC ? (i >> 1) : i;
    const int mu_c = (row_ok && mu < n) ? mu : 0;
    const int spin = i & 1;
    double gate_max = -1.0;
    double re[NH], im[NH], gs[GATE ? NH : 1];
#pragma unroll
    for (int nu = 0; nu < NH; ++nu) { re[nu] = 0.0; im[nu] = 0.0; if (GATE) gs[nu] = 0.0; }
#pragma unroll 2
    for (int b = 0; b < nb; ++b) {
        const cplx ph = phs[b];
        const double* Lb = s.Lt + (size_t)b * n * n + mu_c;
        const double* Db = s.Dt + (size_t)b * n * n + mu_c;
#pragma unroll
        for (int nu = 0; nu < NH; ++nu) {
            if (nu < n) {
                double t = Lb[nu * n];
                re[nu] = fma(t, ph.x, re[nu]); im[nu] = fma(t, ph.y, im[nu]);
                if (GATE) gs[nu] = fma(Db[nu * n], ph.x, gs[nu]);
            }
        }
    }
    const bool orb_ok = row_ok && mu < n;
#pragma unroll
    for (int nu = 0; nu < NH; ++nu) {
        double r_ = orb_ok ? re[nu] : 0.0, i_ = orb_ok ? im[nu] : 0.0;
        if (nu == mu) { r_ += orb_ok ? s.On[mu_c] : 0.0; i_ = 0.0; }
        else if (nu > mu) i_ = -i_;
        if (SOC) {
            cplx s0 = cmake(0.0, 0.0), s1 = s0;
            if (row_ok && 2 * nu < N) {
                s0 = s.St[(size_t)(2 * nu) * N + i]; s1 = s.St[(size_t)(2 * nu + 1) * N + i];
                if (GATE) {
                    double gq = (orb_ok && nu < n) ? gs[nu] : 0.0;
                    gate_max = fmax(gate_max, (spin == 0 ? gq : 0.0) + s.SA[(size_t)(2 * nu) * N + i]);
                    gate_max = fmax(gate_max, (spin == 1 ? gq : 0.0) + s.SA[(size_t)(2 * nu + 1) * N + i]);
                }
            }
            A[2 * nu] = cmake((spin == 0 ? r_ : 0.0) + s0.x, (spin == 0 ? i_ : 0.0) + s0.y);
            A[2 * nu + 1] = cmake((spin == 1 ? r_ : 0.0) + s1.x, (spin == 1 ? i_ : 0.0) + s1.y);
        } else {
            if (GATE && orb_ok && nu < n) gate_max = fmax(gate_max, gs[nu]);
            A[nu] = cmake(r_, i_);
        }
    }
    return gate_max;
}

// ---------------------------------------------------------------------------------------------------------
// NC: compile-time column count (>= N, multiple of 4); L: lanes per matrix (power of two >= NC);
// SOC: kron(h, I2) + soc_mat assembly; GATE: evaluate the reference's Hermiticity gate per k.
// ---------------------------------------------------------------------------------------------------------
template <int NC, int L, bool SOC, bool GATE>
__global__ void __launch_bounds__(SMALL_WARPS * 32, (NC >= 24 ? 3 : 4)) tridiag_small_kernel(SmallArgs a) {
    constexpr int G = 32 / L;                 // matrices side by side in a warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = a.t.n_bonds;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = lane & (L - 1), g = lane / L;
    const SmemView s = carve_and_load<SOC, GATE>(smem_raw, a);
    const size_t per_warp_cplx = (size_t)2 * 64 + (size_t)G * nb;   // v,w (zero-padded) | phases
    cplx* myw = s.warp_base + per_warp_cplx * warp;
    cplx* vbuf = myw;
    cplx* wbuf = myw + 64;
    cplx* phs = myw + 128 + (size_t)g * nb;
    for (int t = lane; t < 128; t += 32) myw[t] = cmake(0.0, 0.0);      // v/w buffers incl. their zero tails
    __syncthreads();

#pragma unroll 1
    for (long long kcta = (long long)blockIdx.x * SMALL_WARPS * G; kcta < a.nk; kcta += (long long)gridDim.x * SMALL_WARPS * G) {
#if SKTB_SMALL_SYNC
        __syncthreads();          // keep the CTA's warps on the same phase body: one instruction stream per CTA
#endif
        long long kq = kcta + warp * G + g;
        const bool k_ok = kq < a.nk;
        if (!k_ok) kq = a.nk - 1;
        bond_phases<L>(a, s, kq, i, phs);
        __syncwarp();
        cplx A[NC];
        const double gate_max = build_row<NC, SOC, GATE>(A, a, s, phs, i);
        if (a.flags) {
            int bad = 0;
            if (GATE) {
                unsigned m = __ballot_sync(0xffffffffu, gate_max > 1.0e-9);
                unsigned gm = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (g * L));
                bad = (m & gm) ? 1 : 0;
            }
            if (i == 0 && k_ok) a.flags[a.flag_off + kq] = bad;
        }

        // ---- K2a: Householder tridiagonalisation, rotating register window ----
        double dval = 0.0, eval = 0.0;
        {
            int j = 0;
            double xn; cplx alpha;
            pivot_reduce<L>(A[0], 0, i, xn, alpha);
            hh_phases<NC, PhaseSteps<NC>::value, NC, L>(A, j, i, g, lane, vbuf, wbuf, dval, eval, xn, alpha);
        }
        dval = (i == NC - 1) ? A[0].x : dval;        // after NC-1 rotations column NC-1 sits in register 0
        if (i < NC && k_ok) {
            double* out = a.scratch + (size_t)kq * (2 * NC);
            out[i] = dval; out[NC + i] = eval;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// N in (24, 32]: "folded" variant.  Rows retire as the reduction proceeds, so after 16 steps half of the lanes of
// a row-per-lane warp are idle.  Each warp therefore runs the wide phases (window 32, 24: steps 0..15) of TWO
// matrices one after the other, parks the two remaining 16x16 trailing blocks in shared memory (4 KB each) and
// finishes both side by side as two 16-lane groups (window 16, 8).  Same arithmetic, ~14 % fewer issue slots.
// ---------------------------------------------------------------------------------------------------------
template <bool SOC, bool GATE>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 3) tridiag_fold32_kernel(SmallArgs a) {
    constexpr int NC = 32, S = SKTB_FOLD_S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = a.t.n_bonds;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SmemView s = carve_and_load<SOC, GATE>(smem_raw, a);
    const size_t per_warp_cplx = (size_t)2 * 64 + (size_t)nb + (size_t)2 * 16 * 16;   // v,w | phases | parked blocks
    cplx* myw = s.warp_base + per_warp_cplx * warp;
    cplx* vbuf = myw;
    cplx* wbuf = myw + 64;
    cplx* phs = myw + 128;
    cplx* park = myw + 128 + nb;                                        // [2][16 cols][16 rows]
    for (int t = lane; t < 128; t += 32) myw[t] = cmake(0.0, 0.0);
    __syncthreads();

#pragma unroll 1
    for (long long kcta = (long long)blockIdx.x * SMALL_WARPS * 2; kcta < a.nk; kcta += (long long)gridDim.x * SMALL_WARPS * 2) {
#if SKTB_SMALL_SYNC
        __syncthreads();
#endif
        const long long kpair = kcta + warp * 2;
        // ---- wide phases of the two matrices, one after the other (32 lanes = 32 rows) ----
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            long long kq = kpair + h;
            const bool k_ok = kq < a.nk;
            if (!k_ok) kq = a.nk - 1;
            bond_phases<32>(a, s, kq, lane, phs);
            __syncwarp();
            cplx A[NC];
            const double gate_max = build_row<NC, SOC, GATE>(A, a, s, phs, lane);
            if (a.flags) {
                int bad = 0;
                if (GATE) bad = __ballot_sync(0xffffffffu, gate_max > 1.0e-9) ? 1 : 0;
                if (lane == 0 && k_ok) a.flags[a.flag_off + kq] = bad;
            }
            double dval = 0.0, eval = 0.0;
            int j = 0;
            double xn; cplx alpha;
            pivot_reduce<32>(A[0], 0, lane, xn, alpha);
            hh_phases<NC, S, NC, 32, 16>(A, j, lane, 0, lane, vbuf, wbuf, dval, eval, xn, alpha);     // steps 0..15
            if (lane < 16 && k_ok) {
                double* out = a.scratch + (size_t)kq * (2 * NC);
                out[lane] = dval; out[NC + lane] = eval;
            }
            if (lane >= 16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) park[(h * 16 + q) * 16 + (lane - 16)] = A[q];
            }
            __syncwarp();
        }
        // ---- narrow phases of both trailing 16x16 blocks side by side (two 16-lane groups) ----
        {
            const int g = lane >> 4, r = lane & 15;
            // v/w layout changes from one 32-lane group to two 16-lane groups: restore the zero tails
            for (int t = lane; t < 64; t += 32) { vbuf[t] = cmake(0.0, 0.0); wbuf[t] = cmake(0.0, 0.0); }
            __syncwarp();
            cplx B[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) B[q] = park[(g * 16 + q) * 16 + r];
            double dval = 0.0, eval = 0.0;
            int j = 0;
            double xn; cplx alpha;
            pivot_reduce<16>(B[0], 0, r, xn, alpha);
            hh_phases<16, S, 16, 16>(B, j, r, g, lane, vbuf, wbuf, dval, eval, xn, alpha);
            dval = (r == 15) ? B[0].x : dval;
            const long long kq = kpair + g;
            if (kq < a.nk) {
                double* out = a.scratch + (size_t)kq * (2 * NC);
                out[16 + r] = dval; out[NC + 16 + r] = eval;
            }
            __syncwarp();
            for (int t = lane; t < 64; t += 32) { vbuf[t] = cmake(0.0, 0.0); wbuf[t] = cmake(0.0, 0.0); }
            __syncwarp();
        }
    }
}

// one tridiagonal matrix per lane: transpose (d,e) through shared memory, QL, bitonic sort, coalesced store
template <int NC, int P2, int MODE = 2>
__global__ void __launch_bounds__(SMALL_WARPS * 32) tql_small_kernel(TqlArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* dbuf = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * NC * DSTRIDE;
    double* ebuf = dbuf + NC * DSTRIDE;
    const long long n_batches = (a.nk + 31) / 32;
    const long long gwarp = (long long)blockIdx.x * SMALL_WARPS + warp, nwarps = (long long)gridDim.x * SMALL_WARPS;
    const int N = a.N;
#pragma unroll 1
    for (long long batch = gwarp; batch < n_batches; batch += nwarps) {
        const long long kbase = batch * 32;
        const int cnt = (int)min((long long)32, a.nk - kbase);
        {   // every lane fetches ITS matrix: NC independent 16-byte loads in flight per lane, stored straight into the
            // lane's own (conflict-free, odd-stride) column - no cross-lane transposition, no load->store dependence chain
            const double2* src = reinterpret_cast<const double2*>(a.scratch + (size_t)(kbase + min(lane, cnt - 1)) * (2 * NC));
#pragma unroll
            for (int t = 0; t < NC / 2; ++t) {
                const double2 v = src[t];
                dbuf[(2 * t) * DSTRIDE + lane] = v.x; dbuf[(2 * t + 1) * DSTRIDE + lane] = v.y;
            }
#pragma unroll
            for (int t = 0; t < NC / 2; ++t) {
                const double2 v = src[NC / 2 + t];
                ebuf[(2 * t) * DSTRIDE + lane] = v.x; ebuf[(2 * t + 1) * DSTRIDE + lane] = v.y;
            }
        }
        __syncwarp();
        if (tql_lane_mode<MODE>(dbuf + lane, ebuf + lane, N)) atomicAdd(a.fail, 1);
        double ev[P2];
#pragma unroll
        for (int q = 0; q < P2; ++q) ev[q] = (q < N && q < NC) ? dbuf[q * DSTRIDE + lane] : __longlong_as_double(0x7ff0000000000000LL);
        bitonic_sort<P2>(ev);
        if (lane < cnt) {
#pragma unroll
            for (int q = 0; q < P2; ++q)
                if (q < N) a.evals[(long long)q * a.ld + a.col0 + kbase + lane] = ev[q];
        }
        __syncwarp();
    }
}

// Per-lane QL on the tridiagonal matrices of a peeled run (64 < N <= 160): (d, e) of the first P steps from the generic
// kernel (head[mat][2][P]) and of the trailing Nt x Nt block from the register-resident kernels (tail[mat][128]: d at
// 0.., e at 64..).  One matrix per lane like tql_small_kernel; the ascending order comes from a rank count in shared
// memory (a register sorting network of this size would not fit).  Dynamic shared memory: warps * 2 * N * DSTRIDE * 8.
template <int MODE>
__global__ void __launch_bounds__(SMALL_WARPS * 32) tql_merge_kernel(const double* __restrict__ head, int P, const double* __restrict__ tail, int Nt,
                                                                     long long nm, double* __restrict__ evals, long long ld, long long col0, int* fail) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int N = P + Nt;
    double* dbuf = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * N * DSTRIDE;
    double* ebuf = dbuf + (size_t)N * DSTRIDE;
    const long long n_batches = (nm + 31) / 32;
#pragma unroll 1
    for (long long batch = (long long)blockIdx.x * warps + warp; batch < n_batches; batch += (long long)gridDim.x * warps) {
        const long long kbase = batch * 32;
        const int cnt = (int)min((long long)32, nm - kbase);
        const long long mat = kbase + min(lane, cnt - 1);
        const double* hd = head + (size_t)mat * 2 * P;
        const double* tl = tail + (size_t)mat * 128;
#pragma unroll 4
        for (int t = 0; t < P; ++t) { dbuf[t * DSTRIDE + lane] = hd[t]; ebuf[t * DSTRIDE + lane] = hd[P + t]; }
#pragma unroll 4
        for (int t = 0; t < Nt; ++t) { dbuf[(P + t) * DSTRIDE + lane] = tl[t]; ebuf[(P + t) * DSTRIDE + lane] = tl[64 + t]; }
        __syncwarp();
        if (tql_lane_mode<MODE>(dbuf + lane, ebuf + lane, N)) atomicAdd(fail, 1);
        // ascending rank of every eigenvalue of this lane's matrix (ties by index), permuted into ebuf
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double di = dbuf[i * DSTRIDE + lane];
            int rank = 0;
#pragma unroll 4
            for (int q = 0; q < N; ++q) {
                const double dq = dbuf[q * DSTRIDE + lane];
                rank += (dq < di || (dq == di && q < i)) ? 1 : 0;
            }
            ebuf[rank * DSTRIDE + lane] = di;
        }
        __syncwarp();
        if (lane < cnt) {
#pragma unroll 4
            for (int q = 0; q < N; ++q) evals[(long long)q * ld + col0 + kbase + lane] = ebuf[q * DSTRIDE + lane];
        }
        __syncwarp();
    }
}

template <int NC> SKTB_D void thread_tridiag(cplx (&A)[NC * (NC + 1) / 2], double (&d)[NC], double (&e)[NC]);   // heev_thread.cuh
#include "heev_pair.cuh"
#include "heev_thread.cuh"
#include "heev_cta64.cuh"
#include "heev_vec32.cuh"
#include "heev_vec64.cuh"

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <class T>
static bool up(const std::vector<T>& h, const T** d, std::vector<void*>* owned) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    if (cudaMalloc(&p, bytes) != cudaSuccess) return false;
    if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    owned->push_back(p);
    *d = reinterpret_cast<const T*>(p);
    return true;
}

inline size_t table_doubles(int n, int nb, int N, bool soc, bool gate) {
    size_t t = (size_t)nb * n * n * (gate ? 2 : 1) + n + 3 * (size_t)nb + ((gate && soc) ? (size_t)N * N : 0);
    t = (t + 1) & ~(size_t)1;
    return t + (soc ? (size_t)2 * N * N : 0);
}

static bool build_tables(const sktb_model_desc* D, const std::vector<double>& T, const std::vector<long long>& off,
                         const std::vector<int>& ni, const std::vector<int>& nj, SmallTables* out, std::vector<void*>* owned,
                         std::string* why) {
    const int n = D->n_orb, nb = D->n_bonds;
    if (n > 32) { *why = ""; return false; }
    if (table_doubles(n, nb, 2 * n <= 32 ? 2 * n : n, 2 * n <= 32, true) * 8 > 64 * 1024) { *why = "dense bond tables exceed the shared-memory budget"; return false; }
    std::vector<double> Lt((size_t)nb * n * n, 0.0), Dt((size_t)nb * n * n, 0.0), dense((size_t)n * n);
    for (int b = 0; b < nb; ++b) {
        std::fill(dense.begin(), dense.end(), 0.0);
        int r0 = D->atom_orb_start[D->bond_i[b]], c0 = D->atom_orb_start[D->bond_j[b]];
        for (int r = 0; r < ni[b]; ++r)
            for (int c = 0; c < nj[b]; ++c) dense[(size_t)(r0 + r) * n + c0 + c] = T[off[b] + (long long)r * nj[b] + c];
        for (int io = 0; io < n; ++io)
            for (int ka = 0; ka < n; ++ka) {
                Lt[((size_t)b * n + ka) * n + io] = ka <= io ? dense[(size_t)io * n + ka] : dense[(size_t)ka * n + io];
                Dt[((size_t)b * n + ka) * n + io] = dense[(size_t)io * n + ka] - dense[(size_t)ka * n + io];
            }
    }
    const int N2 = 2 * n;
    std::vector<cplx> S((size_t)N2 * N2, cmake(0.0, 0.0)), St((size_t)N2 * N2);
    std::vector<double> SAt((size_t)N2 * N2);
    for (int t = 0; t < D->soc_nnz; ++t) S[(size_t)D->soc_row[t] * N2 + D->soc_col[t]] = cmake(D->soc_val[2 * t], D->soc_val[2 * t + 1]);
    for (int r = 0; r < N2; ++r)
        for (int c = 0; c < N2; ++c) {
            cplx e = c < r ? S[(size_t)r * N2 + c] : (c > r ? cconj(S[(size_t)c * N2 + r]) : cmake(S[(size_t)r * N2 + r].x, 0.0));
            St[(size_t)c * N2 + r] = e;
            SAt[(size_t)c * N2 + r] = S[(size_t)r * N2 + c].x - S[(size_t)c * N2 + r].x;
        }
    std::vector<double> bvec(D->bond_vec, D->bond_vec + (size_t)3 * nb), ons(D->onsite, D->onsite + n);
    out->n_orb = n; out->n_bonds = nb; out->soc_any = D->soc_nnz > 0;
    for (int c = 0; c < 9; ++c) out->rec[c] = D->rec_lattice[c];
    std::vector<int> order;
    thread_bond_order(bvec.data(), nb, &order);
    if (!up(Lt, &out->Lt, owned) || !up(Dt, &out->Dt, owned) || !up(bvec, &out->bvec, owned) || !up(ons, &out->onsite, owned) ||
        !up(St, &out->St, owned) || !up(SAt, &out->SAt, owned) || !up(order, &out->thread_order, owned)) { *why = "cudaMalloc failed"; return false; }
    return true;
}

struct LaunchCtx {
    SmallArgs a;               // a.nk / a.first / a.k describe the WHOLE request
    double* evals; long long ld, col0;
    int* fail;
    double* scratch; size_t scratch_bytes;
    cudaStream_t st;
    std::string* info;
};

// ---------------------------------------------------------------------------------------------------------
// pair-layout path (24 < N <= 32, no gate): wide kernel -> tail kernel -> QL, chunked and pipelined over two
// internal streams so that tail + QL of chunk c run beside the wide kernel of chunk c+1 on the same SMs.
//   SKTB_PAIR_MODE = 0: one kernel does all 31 steps (no tail kernel), caller's stream
//                    1: wide, tail, QL back to back on the caller's stream
//                    2: pipelined (default)
// ---------------------------------------------------------------------------------------------------------
struct PairStreams {
    cudaStream_t sA = nullptr, sB = nullptr;
    cudaEvent_t evStart = nullptr, evWide[2] = {nullptr, nullptr}, evDone[2] = {nullptr, nullptr}, evEndA = nullptr;
};
static PairStreams* pair_streams(int dev) {
    static thread_local PairStreams tab[64];      // per host thread: calls on one plan are not re-entrant, plans on other threads get their own
    PairStreams& p = tab[dev & 63];
    if (!p.sA) {
        if (cudaStreamCreateWithFlags(&p.sA, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&p.sB, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&p.evStart, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&p.evEndA, cudaEventDisableTiming);
        for (int b = 0; b < 2; ++b) {
            cudaEventCreateWithFlags(&p.evWide[b], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&p.evDone[b], cudaEventDisableTiming);
        }
    }
    return &p;
}

constexpr size_t PAIR_BYTES_PER_K = 2 * 32 * 8 + 256 * 16;      // (d,e) + trailing 16x16 block

template <bool SOC>
static int launch_pair(const LaunchCtx& c) {
    constexpr int NC = 32;
    static const int mode = getenv("SKTB_PAIR_MODE") ? atoi(getenv("SKTB_PAIR_MODE")) : 2;
    static const int pair_s = getenv("SKTB_PAIR_S") ? atoi(getenv("SKTB_PAIR_S")) : SKTB_PAIR_S;
    const int n = c.a.t.n_orb, nb = c.a.t.n_bonds, N = c.a.N;
    // (3-warp wide CTAs that leave registers for a co-resident tail/QL CTA, and 3 CTAs x 3 warps, were measured:
    //  2.6e7 / 2.8e7 k/s against 3.4e7 for 2 x 4 warps at 255 registers - not kept)
    static const int tail_s = getenv("SKTB_TAIL_S") ? atoi(getenv("SKTB_TAIL_S")) : 4;
    constexpr int WWr = PAIR_WARPS;
    const size_t smemW = table_doubles(n, nb, N, SOC, false) * 8 + (size_t)WWr * (pair_buf_cplx<32>() + nb) * 16;
    const size_t smemT = (size_t)PAIR_WARPS * 2 * pair_buf_cplx<16>() * 16;
    const size_t smemB = (size_t)SMALL_WARPS * 2 * NC * DSTRIDE * 8;
    void (*kW)(SmallArgs, cplx*, cplx*);
    constexpr int SRC = SOC ? PAIR_SRC_SOC : PAIR_SRC_NOSOC;
    if (mode == 0) kW = pair_s == 8 ? pair_wide32_kernel<SRC, 8, true> : pair_wide32_kernel<SRC, 4, true>;
    else kW = pair_s == 8 ? pair_wide32_kernel<SRC, 8, false> : pair_wide32_kernel<SRC, 4, false>;
    // SKTB_TAIL8=1: one-matrix-per-thread stage for the last 8 x 8 block (pair_tail16 stops after its two-slot phases).
    // Measured (same box, T 256^3): 3.69e7 k/s against 3.76e7 with all 15 tail steps in pair_tail16 - the 1 KB per-thread
    // block hand-off and the 196-register thread kernel cost more than the 7 lane-group steps they replace; kept as a knob.
    static const bool tail8 = getenv("SKTB_TAIL8") && getenv("SKTB_TAIL8")[0] == '1';
    void (*kT)(const cplx*, double*, long long, int, int, int, cplx*) =
        tail8 ? (tail_s == 8 ? pair_tail16_kernel<8, false, true> : pair_tail16_kernel<4, false, true>) : (tail_s == 8 ? pair_tail16_kernel<8> : pair_tail16_kernel<4>);
    const int tql_mode = tql_mode_env();
    void (*kB)(TqlArgs) = tql_mode == 2 ? tql_small_kernel<NC, 32, 2> : (tql_mode == 1 ? tql_small_kernel<NC, 32, 1> : tql_small_kernel<NC, 32, 0>);
    if (smemW > 48 * 1024 && cudaFuncSetAttribute(kW, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemW) != cudaSuccess) {
        *c.info = "cudaFuncSetAttribute(smem=" + std::to_string(smemW) + ") failed"; return -1;
    }
    if (smemB > 48 * 1024 && cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB) != cudaSuccess) {
        *c.info = "cudaFuncSetAttribute(smem=" + std::to_string(smemB) + ") failed"; return -1;
    }
    int occW = 0, occT = 0, occB = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occW, kW, WWr * 32, smemW);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occT, kT, PAIR_WARPS * 32, smemT);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, kB, SMALL_WARPS * 32, smemB);
    if (occW < 1 || occT < 1 || occB < 1) { *c.info = "pair kernels do not fit (smem " + std::to_string(smemW) + ")"; return -1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PairStreams* ps = mode == 2 ? pair_streams(dev) : nullptr;
    if (mode == 2 && !ps) { *c.info = "cannot create the pipeline streams"; return -1; }
    const int nbuf = mode == 2 ? 2 : 1;
    const long long chunk = std::max<long long>(32, std::min<long long>((c.a.nk + 31) / 32 * 32, (long long)(c.scratch_bytes / (nbuf * PAIR_BYTES_PER_K)) / 32 * 32));
    unsigned char* base = reinterpret_cast<unsigned char*>(c.scratch);
    cudaStream_t sW = mode == 2 ? ps->sA : c.st, sT = mode == 2 ? ps->sB : c.st;
    if (mode == 2) {
        cudaEventRecord(ps->evStart, c.st);
        cudaStreamWaitEvent(ps->sA, ps->evStart, 0);
        cudaStreamWaitEvent(ps->sB, ps->evStart, 0);
    }
    int launches = 0;
    unsigned gridW = 0, gridT = 0, gridB = 0;
    long long ci = 0;
    for (long long c0 = 0; c0 < c.a.nk; c0 += chunk, ++ci) {
        const int b = (int)(ci % nbuf);
        const long long cnt = std::min(chunk, c.a.nk - c0);
        double* de = reinterpret_cast<double*>(base + (size_t)b * chunk * PAIR_BYTES_PER_K);
        cplx* blocks = reinterpret_cast<cplx*>(base + (size_t)b * chunk * PAIR_BYTES_PER_K + (size_t)chunk * 2 * NC * 8);
        SmallArgs a = c.a;
        a.k = c.a.k ? c.a.k + 3 * c0 : nullptr;
        a.first = c.a.first + c0; a.nk = cnt; a.scratch = de;
        a.flag_off = c.a.flag_off + c0;
        if (mode == 2 && ci >= nbuf) cudaStreamWaitEvent(sW, ps->evDone[b], 0);       // buffer b is free again
        gridW = (unsigned)std::max<long long>(1, std::min<long long>((cnt + WWr - 1) / WWr, (long long)sms * occW));
        kW<<<gridW, WWr * 32, smemW, sW>>>(a, blocks, nullptr);
        ++launches;
        if (mode == 2) { cudaEventRecord(ps->evWide[b], sW); cudaStreamWaitEvent(sT, ps->evWide[b], 0); }
        if (mode != 0) {
            gridT = (unsigned)std::max<long long>(1, std::min<long long>(((cnt + 1) / 2 + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 8));
            kT<<<gridT, PAIR_WARPS * 32, smemT, sT>>>(blocks, de, cnt, 2 * NC, 0, NC, nullptr);
            ++launches;
            if (tail8) {
                tail8_thread_kernel<0><<<(unsigned)std::max<long long>(1, std::min<long long>((cnt + 127) / 128, (long long)sms * 8)), 128, 0, sT>>>(blocks, de, cnt, 2 * NC, 0, NC);
                ++launches;
            }
        }
        TqlArgs t{de, cnt, N, c.evals, c.ld, c.col0 + c0, c.fail};
        gridB = (unsigned)std::max<long long>(1, std::min<long long>(((cnt + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * occB));
        kB<<<gridB, SMALL_WARPS * 32, smemB, sT>>>(t);
        ++launches;
        if (mode == 2) cudaEventRecord(ps->evDone[b], sT);
        if (cudaGetLastError() != cudaSuccess) { *c.info = "launch failed"; return -1; }
    }
    if (mode == 2) {
        cudaEventRecord(ps->evEndA, ps->sA);
        cudaStreamWaitEvent(c.st, ps->evEndA, 0);
        cudaStreamWaitEvent(c.st, ps->evDone[(int)((ci - 1) % nbuf)], 0);
        if (ci >= 2) cudaStreamWaitEvent(c.st, ps->evDone[(int)(ci % nbuf)], 0);
    }
    char buf[400];
    snprintf(buf, sizeof buf, "pair_wide32<soc=%d,S=%d,mode=%d> grid=%u x %d thr smem=%zuB %d CTA/SM + pair_tail16%s grid=%u smem=%zuB + tql_small grid=%u smem=%zuB %d CTA/SM; N=%d, chunk=%lld k",
             (int)SOC, pair_s, mode, gridW, WWr * 32, smemW, occW, tail8 ? "<stop8> + tail8_thread" : "", gridT, smemT, gridB, smemB, occB, N, chunk);
    *c.info = buf;
    return launches;
}

template <int NC, int L, bool SOC, bool GATE>
static int launch_one(const LaunchCtx& c) {
    constexpr bool FOLD = (NC == 32);         // two matrices per warp in the trailing half (tridiag_fold32_kernel)
    if constexpr (NC == 32 && !GATE) {        // pair layout (heev_pair.cuh); SKTB_NO_PAIR=1 falls back to fold32 (A/B runs)
        static const bool no_pair = getenv("SKTB_NO_PAIR") != nullptr && getenv("SKTB_NO_PAIR")[0] == '1';
        if (!no_pair) return launch_pair<SOC>(c);
    }
    constexpr int G = FOLD ? 2 : 32 / L;      // k-points per warp per iteration
    constexpr int P2 = L;
    const int n = c.a.t.n_orb, nb = c.a.t.n_bonds, N = c.a.N;
    const size_t per_warp = FOLD ? ((size_t)128 + nb + 512) : ((size_t)128 + (size_t)(32 / L) * nb);
    const int warpsA = SMALL_WARPS;
    size_t smemA = table_doubles(n, nb, N, SOC, GATE) * 8 + (size_t)warpsA * per_warp * 16;
    size_t smemB = (size_t)SMALL_WARPS * 2 * NC * DSTRIDE * 8;
    void (*kA)(SmallArgs);
    if constexpr (FOLD) kA = tridiag_fold32_kernel<SOC, GATE>;
    else kA = tridiag_small_kernel<NC, L, SOC, GATE>;
    void (*kB)(TqlArgs) = tql_mode_env() == 2 ? tql_small_kernel<NC, P2, 2> : tql_small_kernel<NC, P2, 1>;
    if (smemA > 48 * 1024 && cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA) != cudaSuccess) {
        *c.info = "cudaFuncSetAttribute(smem=" + std::to_string(smemA) + ") failed"; return -1;
    }
    if (smemB > 48 * 1024 && cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB) != cudaSuccess) {
        *c.info = "cudaFuncSetAttribute(smem=" + std::to_string(smemB) + ") failed"; return -1;
    }
    int occA = 0, occB = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, kA, warpsA * 32, smemA);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, kB, SMALL_WARPS * 32, smemB);
    if (occA < 1 || occB < 1) { *c.info = "small kernels do not fit (smem " + std::to_string(smemA) + "/" + std::to_string(smemB) + ")"; return -1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long per_k = 2LL * NC * 8;
    const long long chunk = std::max<long long>(32, std::min<long long>(c.a.nk, (long long)(c.scratch_bytes / per_k) / 32 * 32));
    int launches = 0;
    unsigned gridA = 0, gridB = 0;
    for (long long c0 = 0; c0 < c.a.nk; c0 += chunk) {
        const long long cnt = std::min(chunk, c.a.nk - c0);
        SmallArgs a = c.a;
        a.k = c.a.k ? c.a.k + 3 * c0 : nullptr;
        a.first = c.a.first + c0; a.nk = cnt; a.scratch = c.scratch;
        a.flag_off = c.a.flag_off + c0;
        long long wantA = (cnt + (long long)warpsA * G - 1) / ((long long)warpsA * G);
        gridA = (unsigned)std::max<long long>(1, std::min<long long>(wantA, (long long)sms * occA));
        kA<<<gridA, warpsA * 32, smemA, c.st>>>(a);
        TqlArgs t{c.scratch, cnt, N, c.evals, c.ld, c.col0 + c0, c.fail};
        long long wantB = ((cnt + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS;
        gridB = (unsigned)std::max<long long>(1, std::min<long long>(wantB, (long long)sms * occB));
        kB<<<gridB, SMALL_WARPS * 32, smemB, c.st>>>(t);
        launches += 2;
        if (cudaGetLastError() != cudaSuccess) { *c.info = "launch failed"; return -1; }
    }
    char buf[320];
    snprintf(buf, sizeof buf, "%s<NC=%d,L=%d,soc=%d,gate=%d> grid=%u x %d thr smem=%zuB %d CTA/SM + tql_small grid=%u smem=%zuB %d CTA/SM; N=%d, chunk=%lld k",
             FOLD ? "tridiag_fold32" : "tridiag_small", NC, L, (int)SOC, (int)GATE, gridA, warpsA * 32, smemA, occA, gridB, smemB, occB, N, chunk);
    *c.info = buf;
    return launches;
}

template <int NC, int L>
static int launch_nc(const LaunchCtx& c, bool gate) {
    if (c.a.soc) return gate ? launch_one<NC, L, true, true>(c) : launch_one<NC, L, true, false>(c);
    return gate ? launch_one<NC, L, false, true>(c) : launch_one<NC, L, false, false>(c);
}

inline int padded_nc(int N) { return N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 12 ? 12 : (N <= 16 ? 16 : (N <= 24 ? 24 : 32)))); }

// 32 < N <= 64, eigenvalues only: H (K1 output, [nm][N][N]) -> tridiag_cta64_kernel -> tql_small_kernel<64>.
// scratch: >= nm * cta64_bytes_per_k() bytes.  Returns the number of kernels launched or -1 (text in *info).
size_t cta64_bytes_per_k();
struct C64Build;    // heev_cta64.cuh: non-NULL -> K0 + K1 fused into tridiag_cta64 (H is ignored and may be NULL)
int launch_cta64(const cplx* H, int N, long long nm, double* scratch, double* evals, long long ld, long long col0, int* fail, cudaStream_t st,
                 std::string* info, int h_ld = 0, long long h_mstride = 0, long long h_off = 0, bool tridiag_only = false,
                 const C64Build* build = nullptr);

// N <= 32 with eigenvectors (no gate): pair-layout tridiagonalisation keeping the reflectors -> vec32_kernel
// (QL with accumulation, sort, back-transformation).  evecs: [N][ld][N] complex (reference layout [band, k, comp]).
// scratch: >= nk * vec32_bytes_per_k() bytes.  Returns the number of kernels launched or -1 (text in *info).
// 32 < N <= 64 with eigenvectors: H [nm][N][N] (K1 output or caller matrices, not modified) -> tridiag_cta64 keeping the
// reflectors -> vecql64 -> vecbt64.  scratch: >= nm * vec64_bytes_per_k() bytes.  evecs [N][ld][N].
size_t vec64_bytes_per_k();
int launch_vec64(const cplx* H, int N, long long nm, double* evals, cplx* evecs, long long ld, long long col0, int* fail, void* scratch,
                 cudaStream_t st, std::string* info, const C64Build* build = nullptr);

// sktb_heev for N <= 32: caller-supplied matrices H [nm][N][N] (lower triangle read, input not modified) through the
// pair kernels in LOAD mode; evecs may be NULL.  scratch: >= nm * vec32_bytes_per_k() bytes.
int launch_heev32(const cplx* H, int N, long long nm, double* evals, cplx* evecs, long long ld, long long col0, int* fail, void* scratch,
                  cudaStream_t st, std::string* info);
size_t vec32_bytes_per_k();
int launch_vec32(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N, double* evals,
                 cplx* evecs, long long ld, long long col0, int* flags, long long flag_off, int* fail, void* scratch, cudaStream_t st,
                 std::string* info);

// returns number of kernels launched or -1 on error (text in *info).  scratch: >= 32 * 2*NC*8 bytes.
int launch_solve(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N,
                 double* evals, long long ld, long long col0, int* flags, long long flag_off, int* fail, double* scratch, size_t scratch_bytes,
                 cudaStream_t st, std::string* info);

}  // namespace small
}  // namespace sktb
