// Eigenvectors for N > 64 (config C5: the 512-orbital ribbon with its edge-projected LDOS; any model beyond the
// register-resident kernels).  Replaces row tracking (R <- R H_j per reflector plus a serial QL whose every rotation is
// applied to the tracked rows: 2-3 % of the FP64 peak) by the three-stage scheme the N <= 64 paths already use, with the
// tridiagonal eigenvectors from INVERSE ITERATION instead of an accumulated QL (for N = 512 the rotation replay alone
// would stream a 2 MB Z through ~5e5 dependent rotations per matrix):
//
//   1. heev_generic_kernel (panel formulation, phase 1: tridiagonalisation only) additionally keeps its reflectors v_j,
//      tau_j (4 MB per 512-orbital matrix, +0.5 % of its traffic) and the tridiagonal (d, e);
//   2. stein_kernel: one eigenpair of the real symmetric tridiagonal per THREAD - LAPACK's dstebz + dstein scheme: split
//      into unreduced blocks, eigenvalue r of a block by bisection on the Sturm count, eigenvector by inverse iteration
//      (LU with partial pivoting of T - lambda I as in dlagtf, solves with perturbed tiny pivots as in dlagts job -1,
//      deterministic pseudo-random start), modified Gram-Schmidt inside TIGHT clusters, one first-order Loewdin correction
//      over dstein's 1e-3 ||T|| clusters; global ascending order by rank.  O(N^2) work per matrix (tools/proto_stein.py is the
//      scalar model of the iteration and of the two-level orthogonalisation);
//   3. vecbt_big_kernel: U = H_0 H_1 ... H_{N-2} Z, one eigenvector per WARP held in registers (lane l owns components
//      l, l+32, ...: N/32 complex values), the reflectors staged panel-wise in shared memory and shared by the CTA's
//      warps; per reflector one fused dot product, one warp reduction, one update.  8 N^3 flops per matrix at register
//      speed; output straight into the reference's [band, k, component] layout, or only the rows an LDOS asks for.
// FP64 CUDA cores only.  (included by sktb.cu)
#pragma once
#include "common.cuh"

namespace sktb {

struct BigVecArgs {
    int N; long long nm;
    const double* de;         // [nm][2][N]  d | e of the tridiagonal
    double* evals; long long ld, col0;         // out: ascending eigenvalues, evals[band * ld + col0 + mat] (stein_kernel: bisection per unreduced block)
    double* Z;                // [nm][N][N]  component-major: Z[i * N + n] = component i of tridiagonal eigenvector n
    double* lu;               // [nm][4][N][N] scratch of the factorisations (a | b | c | d2), same indexing
    unsigned char* pin;       // [nm][N][N] pivot flags
    const cplx* vstore;       // [nm][N][N]  reflector j in row j (entries i <= j are zero, entry j+1 is one)
    const cplx* taustore;     // [nm][N]
    cplx* vec_out; long long vs_band, vs_k, vcol0;   // vec_out[band * vs_band + (vcol0 + mat) * vs_k + r]
    int n_rows; const int* rows;   // n_rows == 0: all N components (r = component); else only components rows[r]
};

SKTB_D double stein_uniform(unsigned a, unsigned b, unsigned c) {          // deterministic start vector in (-1, 1)
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return ((double)h + 0.5) * (2.0 / 4294967296.0) - 1.0;
}

__global__ void __launch_bounds__(1024) stein_kernel(BigVecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, T = blockDim.x, tid = threadIdx.x;
    double* d = reinterpret_cast<double*>(smem_raw);
    double* e = d + N;                 // e[k] couples k, k+1; negligible entries are set to zero (block boundaries)
    double* w = e + N;                 // eigenvalue of thread n (n-th position of its block, ascending inside the block)
    int* blo = reinterpret_cast<int*>(w + N);   // first / last index of the unreduced block that contains index n
    int* bhi = blo + N;
    double2* de2 = reinterpret_cast<double2*>(bhi + N);      // [N] (d_i, e_{i-1}^2) scaled by an exact power of two: product-form Sturm counts
    const double eps = 2.220446049250313e-16;
    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        __syncthreads();
        const double* de = a.de + (size_t)mat * 2 * N;
        for (int i = tid; i < N; i += T) { d[i] = de[i]; e[i] = (i < N - 1) ? de[N + i] : 0.0; }
        __syncthreads();
        double onenrm = 0.0;
        for (int i = 0; i < N; ++i) onenrm = fmax(onenrm, fabs(d[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0) + fabs(e[i]));
        __syncthreads();
        // ---- split into unreduced blocks (dstebz / dstein work block by block: eigenvectors of different blocks are
        //      orthogonal by support, so exact degeneracies ACROSS blocks - diagonal, zero, decoupled matrices - cost nothing)
        for (int i = tid; i < N - 1; i += T) if (fabs(e[i]) <= eps * onenrm) e[i] = 0.0;
        __syncthreads();
        int sex = 0;
        if (onenrm > 0.0) frexp(onenrm, &sex);
        const double ssc = scalbn(1.0, -sex), issc = scalbn(1.0, sex);
        for (int i = tid; i < N; i += T) { const double el = i > 0 ? e[i - 1] * ssc : 0.0; de2[i] = make_double2(d[i] * ssc, el * el); }
        __syncthreads();
        double* Zm = a.Z + (size_t)mat * N * N;
        double* La = a.lu + (size_t)mat * 4 * N * N;
        double* Lb = La + (size_t)N * N;
        double* Lc = Lb + (size_t)N * N;
        double* Ld = Lc + (size_t)N * N;
        unsigned char* Pn = a.pin + (size_t)mat * N * N;
        int b0 = 0, b1 = 0;
        double lam = 0.0;
        if (tid < N) {
            b0 = tid; b1 = tid;
            while (b0 > 0 && e[b0 - 1] != 0.0) --b0;
            while (b1 < N - 1 && e[b1] != 0.0) ++b1;
            blo[tid] = b0; bhi[tid] = b1;
            // ---- eigenvalue r = tid - b0 of the block by bisection on the Sturm count (dstebz recurrence) ----
            const int r = tid - b0;
            if (b1 == b0) lam = d[b0];
            else {
                double gl = 1e300, gu = -1e300, tn = 0.0;
                for (int i = b0; i <= b1; ++i) {
                    const double el = i > b0 ? fabs(e[i - 1]) : 0.0, er = i < b1 ? fabs(e[i]) : 0.0;
                    gl = fmin(gl, d[i] - el - er); gu = fmax(gu, d[i] + el + er);
                    tn = fmax(tn, fmax(fabs(d[i]), fmax(el, er)));
                }
                const int nbk = b1 - b0 + 1;
                // bisection in the scaled units of de2 (|T| <= 1): product-form Sturm count, see sturm_count_prod
                const double pivmin = 1e-290;
                double lo = gl * ssc - (2.0 * eps * (tn * ssc) * nbk + 2.0 * pivmin), hi = gu * ssc + (2.0 * eps * (tn * ssc) * nbk + 2.0 * pivmin);
                for (int it = 0; it < 128; ++it) {
                    const double x = 0.5 * (lo + hi);
                    if (!(hi - lo > 2.0 * eps * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) || x <= lo || x >= hi) break;
                    if (sturm_count_prod(de2, b0, b1, x) > r) hi = x; else lo = x;
                }
                lo *= issc; hi *= issc;
                lam = 0.5 * (lo + hi);
            }
            w[tid] = lam;
        }
        __syncthreads();
        // ---- inverse iteration inside the block.  Two levels: TIGHT clusters (gaps <= 1e-10 ||T||: degeneracies that survive the
        //      block splitting) are computed by the thread of their first eigenvalue, one vector after the other, with modified
        //      Gram-Schmidt inside the iteration (dstein); everything else independently, one thread per eigenvector, followed
        //      by TWO first-order symmetric (Loewdin) corrections over the PAIRWISE window |w_m - w_n| <= 2e-4 ||T||:
        //      independent inverse iterations overlap by ~eps ||T|| / gap (below 1e-12 outside the window, up to 2e-6 inside),
        //      each correction squares the overlap of a window.  (dstein's serial Gram-Schmidt over its 1e-3 ||T|| clusters
        //      would leave one thread with tens of vectors for the dense spectrum of a 512-orbital ribbon.) ----
        const double ptol = 2.0e-4 * onenrm, tight = 1.0e-10 * onenrm;
        const double tol = onenrm > 0.0 ? eps * onenrm : eps;
        const bool head = tid < N && (tid == b0 || w[tid] - w[tid - 1] > tight);
        if (tid < N) for (int i = 0; i < N; ++i) if (i < b0 || i > b1) Zm[(size_t)i * N + tid] = 0.0;
        if (head && b1 == b0) Zm[(size_t)b0 * N + tid] = 1.0;
        else if (head) {
            const int nbk = b1 - b0 + 1;
            double xjm = 0.0;
            for (int j = tid; j <= b1 && (j == tid || w[j] - w[j - 1] <= tight); ++j) {
                double xj = w[j];
                if (j > tid) {
                    // dstein separates the shifts of a cluster by 10 eps |xj|; here with an ABSOLUTE floor (10 eps ||T||): the
                    // bisection above runs to relative precision and returns e.g. +-1.3e-38 for the two edge states of a wide
                    // ribbon - with (numerically) equal shifts both solves amplify the same vector and Gram-Schmidt is left with noise
                    const double pertol = 10.0 * eps * fmax(fabs(xj), onenrm);
                    if (xj - xjm < pertol) xj = xjm + pertol;
                }
                xjm = xj;
                // ---- dlagtf: T_block - xj I = P L U ----
                double ak = d[b0] - xj, bk = e[b0];
                double scale1 = fabs(ak) + fabs(bk);
                for (int k = b0; k < b1; ++k) {
                    double ck = e[k], ak1 = d[k + 1] - xj, bk1 = (k < b1 - 1) ? e[k + 1] : 0.0, d2k = 0.0;
                    const double scale2 = fabs(ck) + fabs(ak1) + fabs(bk1);
                    const double piv1 = ak == 0.0 ? 0.0 : fabs(ak) / scale1;
                    unsigned char pv = 0;
                    double sa = ak, sb = bk;
                    if (ck == 0.0) scale1 = scale2;
                    else if (fabs(ck) / scale2 <= piv1) {
                        scale1 = scale2;
                        ck = ck / ak;
                        ak1 -= ck * bk;
                    } else {
                        pv = 1;
                        const double mult = ak / ck;
                        sa = ck;
                        const double temp = ak1;
                        ak1 = bk - mult * temp;
                        if (k < b1 - 1) { d2k = bk1; bk1 = -mult * d2k; }
                        sb = temp;
                        ck = mult;
                    }
                    const size_t at = (size_t)k * N + j;
                    La[at] = sa; Lb[at] = sb; Lc[at] = ck; Ld[at] = d2k; Pn[at] = pv;
                    ak = ak1; bk = bk1;
                }
                La[(size_t)b1 * N + j] = ak;
                const double a_last = ak;
                // ---- inverse iteration ----
                for (int i = b0; i <= b1; ++i) Zm[(size_t)i * N + j] = stein_uniform((unsigned)i, (unsigned)j, (unsigned)N);
                int nrmchk = 0;
                const double dtpcrt = sqrt(0.1 / nbk);
                for (int it = 0; it < 8; ++it) {
                    double s1 = 0.0;
                    for (int i = b0; i <= b1; ++i) s1 += fabs(Zm[(size_t)i * N + j]);
                    const double scl = (double)nbk * fmax(onenrm, 1e-300) * fmax(eps, fabs(a_last)) / fmax(s1, 1e-300);
                    // forward: y <- L^-1 P y (with the scaling of the start vector folded in)
                    // (the loads of four steps are issued together: the recurrence is latency-bound on its global-memory operands)
                    double yprev = Zm[(size_t)b0 * N + j] * scl;
                    int k = b0 + 1;
#define SKTB_FWD(YK, CV, PV, ROW)                                                                   \
                    {                                                                               \
                        double yk = (YK) * scl;                                                     \
                        if (PV) { const double t = yprev; yprev = yk; yk = t - (CV) * yk; }         \
                        else yk -= (CV) * yprev;                                                    \
                        Zm[(size_t)(ROW) * N + j] = yprev;                                          \
                        yprev = yk;                                                                 \
                    }
                    for (; k + 3 <= b1; k += 4) {
                        const size_t lk = (size_t)(k - 1) * N + j;
                        const double y0 = Zm[lk + N], y1 = Zm[lk + 2 * (size_t)N], y2 = Zm[lk + 3 * (size_t)N], y3 = Zm[lk + 4 * (size_t)N];
                        const double c0 = Lc[lk], c1 = Lc[lk + N], c2 = Lc[lk + 2 * (size_t)N], c3 = Lc[lk + 3 * (size_t)N];
                        const unsigned char p0 = Pn[lk], p1 = Pn[lk + N], p2 = Pn[lk + 2 * (size_t)N], p3 = Pn[lk + 3 * (size_t)N];
                        SKTB_FWD(y0, c0, p0, k - 1) SKTB_FWD(y1, c1, p1, k) SKTB_FWD(y2, c2, p2, k + 1) SKTB_FWD(y3, c3, p3, k + 2)
                    }
                    for (; k <= b1; ++k) {
                        const size_t lk = (size_t)(k - 1) * N + j;
                        SKTB_FWD(Zm[(size_t)k * N + j], Lc[lk], Pn[lk], k - 1)
                    }
#undef SKTB_FWD
                    Zm[(size_t)b1 * N + j] = yprev;
                    // backward: x <- U^-1 y, tiny pivots perturbed to +-tol
                    double x1 = 0.0, x2 = 0.0;
                    int kb = b1;
#define SKTB_BWD(KK, TV, LBV, LDV, LAV)                                                             \
                    {                                                                               \
                        double t = (TV) * rs;                                                       \
                        if ((KK) <= b1 - 1) t -= (LBV) * x1;                                        \
                        if ((KK) <= b1 - 2) t -= (LDV) * x2;                                        \
                        double akk = (LAV);                                                         \
                        if (fabs(akk) < tol) akk = akk >= 0.0 ? tol : -tol;                         \
                        t /= akk;                                                                   \
                        if (fabs(t) > 1.0e150) {       /* exponentially localised vectors (edge states) span hundreds of decades: rescale */ \
                            for (int kk = b0; kk <= b1; ++kk) Zm[(size_t)kk * N + j] *= 1.0e-150;   /* (solution so far and the right-hand side still ahead) */ \
                            t *= 1.0e-150; x1 *= 1.0e-150; x2 *= 1.0e-150; rs *= 1.0e-150;          /* rs: right-hand sides of this block already in registers */ \
                        }                                                                           \
                        Zm[(size_t)(KK) * N + j] = t;                                               \
                        x2 = x1; x1 = t;                                                            \
                    }
                    for (; kb - 3 >= b0; kb -= 4) {
                        const size_t at = (size_t)kb * N + j;
                        const double t0 = Zm[at], t1 = Zm[at - N], t2 = Zm[at - 2 * (size_t)N], t3 = Zm[at - 3 * (size_t)N];
                        const double lb0 = Lb[at], lb1 = Lb[at - N], lb2 = Lb[at - 2 * (size_t)N], lb3 = Lb[at - 3 * (size_t)N];
                        const double ld0 = Ld[at], ld1 = Ld[at - N], ld2 = Ld[at - 2 * (size_t)N], ld3 = Ld[at - 3 * (size_t)N];
                        const double la0 = La[at], la1 = La[at - N], la2 = La[at - 2 * (size_t)N], la3 = La[at - 3 * (size_t)N];
                        double rs = 1.0;
                        SKTB_BWD(kb, t0, lb0, ld0, la0) SKTB_BWD(kb - 1, t1, lb1, ld1, la1) SKTB_BWD(kb - 2, t2, lb2, ld2, la2) SKTB_BWD(kb - 3, t3, lb3, ld3, la3)
                    }
                    for (; kb >= b0; --kb) {
                        const size_t at = (size_t)kb * N + j;
                        double rs = 1.0;
                        SKTB_BWD(kb, Zm[at], Lb[at], Ld[at], La[at])
                    }
#undef SKTB_BWD
                    // modified Gram-Schmidt against the earlier vectors of the tight cluster
                    for (int q = tid; q < j; ++q) {
                        double dot = 0.0;
                        for (int i = b0; i <= b1; ++i) dot = fma(Zm[(size_t)i * N + j], Zm[(size_t)i * N + q], dot);
                        for (int i = b0; i <= b1; ++i) Zm[(size_t)i * N + j] = fma(-dot, Zm[(size_t)i * N + q], Zm[(size_t)i * N + j]);
                    }
                    double mx = 0.0;
                    for (int i = b0; i <= b1; ++i) mx = fmax(mx, fabs(Zm[(size_t)i * N + j]));
                    if (mx >= dtpcrt && ++nrmchk >= 3) break;
                }
                double mxf = 0.0;
                for (int i = b0; i <= b1; ++i) mxf = fmax(mxf, fabs(Zm[(size_t)i * N + j]));
                const double imx = mxf > 0.0 ? 1.0 / mxf : 0.0;     // scaled 2-norm: the squares of a 1e160 entry would overflow
                double nn = 0.0;
                for (int i = b0; i <= b1; ++i) { const double z = Zm[(size_t)i * N + j] * imx; nn = fma(z, z, nn); }
                const double inv = nn > 0.0 ? imx / sqrt(nn) : 0.0;
                for (int i = b0; i <= b1; ++i) Zm[(size_t)i * N + j] *= inv;
            }
        }
        __syncthreads();
        // ---- two Loewdin passes (Zm -> Lb -> final), global ascending rank (ties by index): the corrected vector goes to
        //      column `rank` of the (now free) first factor array = the final Z, the eigenvalue to evals[rank] ----
        int wlo = tid, whi = tid;
        if (tid < N) {
            const int j = tid;
            while (wlo > b0 && w[j] - w[wlo - 1] <= ptol) --wlo;
            while (whi < b1 && w[whi + 1] - w[j] <= ptol) ++whi;
            for (int i = 0; i < N; ++i) Lb[(size_t)i * N + j] = Zm[(size_t)i * N + j];
            for (int m = wlo; m <= whi; ++m) {
                if (m == j) continue;
                double dot = 0.0;
                for (int i = b0; i <= b1; ++i) dot = fma(Zm[(size_t)i * N + m], Zm[(size_t)i * N + j], dot);
                const double h = -0.5 * dot;
                for (int i = b0; i <= b1; ++i) Lb[(size_t)i * N + j] = fma(h, Zm[(size_t)i * N + m], Lb[(size_t)i * N + j]);
            }
        }
        __syncthreads();
        if (tid < N) {
            const int j = tid;
            const double wj = w[j];
            int rank = 0;
            for (int m = 0; m < N; ++m) { const double wm = w[m]; rank += (wm < wj) || (wm == wj && m < j); }
            a.evals[(long long)rank * a.ld + a.col0 + mat] = wj;
            for (int i = 0; i < N; ++i) La[(size_t)i * N + rank] = Lb[(size_t)i * N + j];
            for (int m = wlo; m <= whi; ++m) {
                if (m == j) continue;
                double dot = 0.0;
                for (int i = b0; i <= b1; ++i) dot = fma(Lb[(size_t)i * N + m], Lb[(size_t)i * N + j], dot);
                const double h = -0.5 * dot;
                for (int i = b0; i <= b1; ++i) La[(size_t)i * N + rank] = fma(h, Lb[(size_t)i * N + m], La[(size_t)i * N + rank]);
            }
        }
    }
}

// U = H_0 ... H_{N-2} Z for BW eigenvectors per CTA (one per warp); NQ = ceil(N / 32) components per lane.
constexpr int VBT_PANEL = 8;        // reflectors staged per shared-memory panel
// EV eigenvectors per warp: every reflector entry read from shared memory feeds EV dot products and EV updates (with one
// vector per warp the kernel is shared-memory-bandwidth bound: ncu showed the LSU data pipe 75 % busy against 49 % FP64).
template <int NQ, int BW, int EV>
__global__ void __launch_bounds__(BW * 32, (NQ * EV <= 16) ? 2 : 1) vecbt_big_kernel(BigVecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* Vp = reinterpret_cast<cplx*>(smem_raw);                         // [VBT_PANEL][NQ * 32] reflector panel
    cplx* tp = Vp + (size_t)VBT_PANEL * NQ * 32;                          // [VBT_PANEL] tau
    const int groups = (N + BW * EV - 1) / (BW * EV);
    for (long long work = blockIdx.x; work < a.nm * groups; work += gridDim.x) {
        const long long mat = work / groups;
        const int n0 = ((int)(work - mat * groups) * BW + warp) * EV;     // this warp's first eigenvector
        const double* Zm = a.lu + (size_t)mat * 4 * N * N;               // stein_kernel's final (corrected) vectors
        const cplx* Vs = a.vstore + (size_t)mat * N * N;
        const cplx* Ts = a.taustore + (size_t)mat * N;
        cplx u[EV][NQ];
#pragma unroll
        for (int ev = 0; ev < EV; ++ev) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = 32 * q + lane;
                u[ev][q] = cmake((n0 + ev < N && i < N) ? Zm[(size_t)i * N + n0 + ev] : 0.0, 0.0);
            }
        }
        // reflectors N-2 ... 0 in panels of VBT_PANEL (descending)
        for (int jhi = N - 2; jhi >= 0; jhi -= VBT_PANEL) {
            const int jlo = max(jhi - VBT_PANEL + 1, 0), np = jhi - jlo + 1;
            __syncthreads();
            for (int t = tid; t < np * NQ * 32; t += BW * 32) {
                const int s = t / (NQ * 32), i = t - s * (NQ * 32);
                Vp[t] = i < N ? Vs[(size_t)(jlo + s) * N + i] : cmake(0.0, 0.0);
            }
            if (tid < np) tp[tid] = Ts[jlo + tid];
            __syncthreads();
            const int q0 = (jlo + 1) >> 5;                                // components below 32 q0 are untouched by this panel
            for (int s = np - 1; s >= 0; --s) {
                const cplx* v = Vp + (size_t)s * NQ * 32 + lane;
                constexpr bool KEEP = NQ * EV <= 8;                        // else the reflector is re-read for the update (registers)
                constexpr int NR = KEEP ? NQ : 1;
                cplx vr[NR];
                cplx dacc[EV][2];                                           // independent chains (a single accumulator is a 2 NQ-deep
#pragma unroll                                                              // dependent DFMA chain per reflector)
                for (int ev = 0; ev < EV; ++ev) dacc[ev][0] = dacc[ev][1] = cmake(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) {
                        const cplx vq = v[32 * q];
                        if (KEEP) vr[q % NR] = vq;
#pragma unroll
                        for (int ev = 0; ev < EV; ++ev) cfmac(dacc[ev][q & 1], vq, u[ev][q]);                        // v^H u
                    }
                }
                cplx wv[EV];
#pragma unroll
                for (int ev = 0; ev < EV; ++ev) {
                    cplx dot = cadd(dacc[ev][0], dacc[ev][1]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o); }
                    wv[ev] = cmul(tp[s], dot);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) {
                        const cplx vq = KEEP ? vr[q % NR] : v[32 * q];
#pragma unroll
                        for (int ev = 0; ev < EV; ++ev) cfms(u[ev][q], wv[ev], vq);                                  // u -= tau (v^H u) v
                    }
                }
            }
        }
        // ---- output: the reference's [band, k, component] layout ----
#pragma unroll
        for (int ev = 0; ev < EV; ++ev) {
            if (n0 + ev < N) {
                cplx* out = a.vec_out + (long long)(n0 + ev) * a.vs_band + (a.vcol0 + mat) * a.vs_k;
#pragma unroll
                for (int q = 0; q < NQ; ++q) { const int i = 32 * q + lane; if (i < N) out[i] = u[ev][q]; }
            }
        }
    }
}

// Selected components only (LDOS / spectral maps: a handful of rows i of U = Q Z): row i of Q = e_i^T H_0 H_1 ... H_{N-2} costs
// one pass over the reflectors (p = Q^H e_i: p <- H_j^H p, j ascending - the same register layout and panel staging as above,
// one ROW per warp), then U[i, n] = sum_k conj(p_k) Z[k, n] for all n.  O(|S| N^2) per matrix instead of the 8 N^3 of the full
// back-transformation: the edge-projected LDOS of the 512-orbital ribbon needs 4 rows.
template <int NQ, int BW>
__global__ void __launch_bounds__(BW * 32, NQ <= 16 ? 2 : 1) rowsq_big_kernel(BigVecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* Vp = reinterpret_cast<cplx*>(smem_raw);                         // [VBT_PANEL][NQ * 32] reflector panel
    cplx* tp = Vp + (size_t)VBT_PANEL * NQ * 32;                          // [VBT_PANEL] tau
    static_assert(BW <= VBT_PANEL, "the staging of the finished rows aliases the reflector panel");
    const int groups = (a.n_rows + BW - 1) / BW;
    for (long long work = blockIdx.x; work < a.nm * groups; work += gridDim.x) {
        const long long mat = work / groups;
        const int r = (int)(work - mat * groups) * BW + warp;             // this warp's selected row
        const bool live = r < a.n_rows;
        const int irow = live ? a.rows[r] : 0;
        const double* Zm = a.lu + (size_t)mat * 4 * N * N;
        const cplx* Vs = a.vstore + (size_t)mat * N * N;
        const cplx* Ts = a.taustore + (size_t)mat * N;
        cplx u[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) u[q] = cmake((live && 32 * q + lane == irow) ? 1.0 : 0.0, 0.0);
        for (int jlo = 0; jlo <= N - 2; jlo += VBT_PANEL) {
            const int np = min(VBT_PANEL, N - 1 - jlo);
            __syncthreads();
            for (int t = tid; t < np * NQ * 32; t += BW * 32) {
                const int s = t / (NQ * 32), i = t - s * (NQ * 32);
                Vp[t] = i < N ? Vs[(size_t)(jlo + s) * N + i] : cmake(0.0, 0.0);
            }
            if (tid < np) tp[tid] = Ts[jlo + tid];
            __syncthreads();
            const int q0 = (jlo + 1) >> 5;
            for (int s = 0; s < np; ++s) {
                const cplx* v = Vp + (size_t)s * NQ * 32 + lane;
                cplx dacc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) dacc[c] = cmake(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) cfmac(dacc[q & 3], v[32 * q], u[q]);
                }
                cplx dot = cadd(cadd(dacc[0], dacc[1]), cadd(dacc[2], dacc[3]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o); }
                const cplx wv = cmul(cconj(tp[s]), dot);                  // H_j^H = I - conj(tau) v v^H
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) cfms(u[q], wv, v[32 * q]);
                }
            }
        }
        __syncthreads();                                                   // every warp is done with the last reflector panel
        cplx* mine = Vp + (size_t)warp * NQ * 32;
#pragma unroll
        for (int q = 0; q < NQ; ++q) mine[32 * q + lane] = u[q];           // p = Q^H e_i: row i of Q is p^H
        __syncwarp();
        if (live) {
            for (int n = lane; n < N; n += 32) {
                double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
                int k = 0;
                for (; k + 1 < N; k += 2) {
                    const double z0 = Zm[(size_t)k * N + n], z1 = Zm[(size_t)(k + 1) * N + n];
                    const cplx p0 = mine[k], p1 = mine[k + 1];
                    ax = fma(p0.x, z0, ax); ay = fma(-p0.y, z0, ay); bx = fma(p1.x, z1, bx); by = fma(-p1.y, z1, by);
                }
                if (k < N) { const double z0 = Zm[(size_t)k * N + n]; const cplx p0 = mine[k]; ax = fma(p0.x, z0, ax); ay = fma(-p0.y, z0, ay); }
                a.vec_out[(long long)n * a.vs_band + (a.vcol0 + mat) * a.vs_k + r] = cmake(ax + bx, ay + by);
            }
        }
        __syncwarp();
    }
}

inline size_t vecbt_big_smem(int NQ) { return ((size_t)VBT_PANEL * NQ * 32 + VBT_PANEL) * sizeof(cplx); }
inline size_t bigvec_bytes_per_matrix(int N) {          // Z | lu[4] | pin | reflectors | tau | (d, e)
    const size_t nn = (size_t)N * N;
    return nn * 8 + 4 * nn * 8 + ((nn + 15) & ~(size_t)15) + nn * 16 + (size_t)N * 16 + (size_t)2 * N * 8 + 64;
}

}  // namespace sktb
