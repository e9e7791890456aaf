// Eigenvectors for N > 64 (config C5: the 512-orbital ribbon with its edge-projected LDOS; any model beyond the
// register-resident kernels).  Replaces row tracking (R <- R H_j per reflector plus a serial QL whose every rotation is
// applied to the tracked rows: 2-3 % of the FP64 peak) by the three-stage scheme the N <= 64 paths already use, with the
// tridiagonal eigenvectors from INVERSE ITERATION instead of an accumulated QL (for N = 512 the rotation replay alone
// would stream a 2 MB Z through ~5e5 dependent rotations per matrix):
//
//   1. heev_generic_kernel (panel formulation, eigenvalues by bisection) additionally keeps its reflectors v_j, tau_j
//      (4 MB per 512-orbital matrix, +0.5 % of its traffic) and the tridiagonal (d, e);
//   2. stein_kernel: one eigenvector of the real symmetric tridiagonal per THREAD - LAPACK dstein's scheme: LU with
//      partial pivoting of T - lambda I (dlagtf), solves with perturbed tiny pivots (dlagts, job = -1), random start,
//      modified Gram-Schmidt inside clusters of eigenvalues closer than 1e-3 ||T||_1 (the thread of a cluster's first
//      eigenvalue computes the cluster's vectors one after the other).  O(N^2) work per matrix.  Negligible off-diagonals
//      are kept instead of splitting T into unreduced blocks: the solves decouple numerically (tools/proto_stein.py is
//      the scalar model, checked on exactly degenerate, split and zero-diagonal matrices);
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
    const double* evals; long long ld, col0;   // sorted eigenvalues: evals[band * ld + col0 + mat]
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
    double* e = d + N;
    double* w = e + N;
    const double eps = 2.220446049250313e-16;
    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        __syncthreads();
        const double* de = a.de + (size_t)mat * 2 * N;
        for (int i = tid; i < N; i += T) { d[i] = de[i]; e[i] = (i < N - 1) ? de[N + i] : 0.0; w[i] = a.evals[(long long)i * a.ld + a.col0 + mat]; }
        __syncthreads();
        double onenrm = 0.0;
        for (int i = 0; i < N; ++i) onenrm = fmax(onenrm, fabs(d[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0) + fabs(e[i]));
        const double ortol = 1.0e-3 * onenrm;
        const double tol = onenrm > 0.0 ? eps * onenrm : eps;
        double* Zm = a.Z + (size_t)mat * N * N;
        double* La = a.lu + (size_t)mat * 4 * N * N;
        double* Lb = La + (size_t)N * N;
        double* Lc = Lb + (size_t)N * N;
        double* Ld = Lc + (size_t)N * N;
        unsigned char* Pn = a.pin + (size_t)mat * N * N;
        // Two levels.  TIGHT clusters (gaps <= 1e-7 ||T||: degeneracies) are computed by the thread of their first eigenvalue,
        // one vector after the other, with modified Gram-Schmidt inside the iteration (dstein).  Everything else runs
        // independently, one thread per eigenvector, and ONE first-order symmetric (Loewdin) correction over the LOOSE cluster
        // (gaps <= 1e-3 ||T||, dstein's criterion) follows: independent inverse iterations are orthogonal to
        // ~eps ||T|| / gap <= 2e-9 there, the correction squares that.  (A serial dstein over the loose clusters of a dense
        // 512-level spectrum would leave one thread with tens of vectors and O(c^2 N) Gram-Schmidt work.)
        const double tight = 1.0e-7 * onenrm;
        const bool head = tid < N && (tid == 0 || w[tid] - w[tid - 1] > tight);
        if (head) {
            double xjm = 0.0;
            for (int j = tid; j < N && (j == tid || w[j] - w[j - 1] <= tight); ++j) {
                double xj = w[j];
                if (j > tid) {
                    const double pertol = 10.0 * fabs(eps * xj);
                    if (xj - xjm < pertol) xj = xjm + pertol;
                }
                xjm = xj;
                // ---- dlagtf: T - xj I = P L U ----
                double ak = d[0] - xj, bk = N > 1 ? e[0] : 0.0;
                double scale1 = fabs(ak) + fabs(bk);
                for (int k = 0; k < N - 1; ++k) {
                    double ck = e[k], ak1 = d[k + 1] - xj, bk1 = (k < N - 2) ? e[k + 1] : 0.0, d2k = 0.0;
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
                        if (k < N - 2) { d2k = bk1; bk1 = -mult * d2k; }
                        sb = temp;
                        ck = mult;
                    }
                    const size_t at = (size_t)k * N + j;
                    La[at] = sa; Lb[at] = sb; Lc[at] = ck; Ld[at] = d2k; Pn[at] = pv;
                    ak = ak1; bk = bk1;
                }
                La[(size_t)(N - 1) * N + j] = ak;
                const double a_last = ak;
                // ---- inverse iteration ----
                for (int i = 0; i < N; ++i) Zm[(size_t)i * N + j] = stein_uniform((unsigned)i, (unsigned)j, (unsigned)N);
                int nrmchk = 0;
                const double dtpcrt = sqrt(0.1 / N);
                for (int it = 0; it < 8; ++it) {
                    double s1 = 0.0;
                    for (int i = 0; i < N; ++i) s1 += fabs(Zm[(size_t)i * N + j]);
                    const double scl = (double)N * fmax(onenrm, 1e-300) * fmax(eps, fabs(a_last)) / fmax(s1, 1e-300);
                    // forward: y <- L^-1 P y (with the scaling of the start vector folded in)
                    double yprev = Zm[j] * scl;
                    for (int k = 1; k < N; ++k) {
                        const size_t lk = (size_t)(k - 1) * N + j;
                        double yk = Zm[(size_t)k * N + j] * scl;
                        const double c = Lc[lk];
                        if (Pn[lk]) { const double t = yprev; yprev = yk; yk = t - c * yk; }
                        else yk -= c * yprev;
                        Zm[(size_t)(k - 1) * N + j] = yprev;
                        yprev = yk;
                    }
                    Zm[(size_t)(N - 1) * N + j] = yprev;
                    // backward: x <- U^-1 y, tiny pivots perturbed to +-tol
                    double x1 = 0.0, x2 = 0.0;
                    for (int k = N - 1; k >= 0; --k) {
                        const size_t at = (size_t)k * N + j;
                        double t = Zm[at];
                        if (k <= N - 2) t -= Lb[at] * x1;
                        if (k <= N - 3) t -= Ld[at] * x2;
                        double akk = La[at];
                        if (fabs(akk) < tol) akk = akk >= 0.0 ? tol : -tol;
                        t /= akk;
                        Zm[at] = t;
                        x2 = x1; x1 = t;
                    }
                    // modified Gram-Schmidt against the earlier vectors of the cluster
                    for (int q = tid; q < j; ++q) {
                        double dot = 0.0;
                        for (int i = 0; i < N; ++i) dot = fma(Zm[(size_t)i * N + j], Zm[(size_t)i * N + q], dot);
                        for (int i = 0; i < N; ++i) Zm[(size_t)i * N + j] = fma(-dot, Zm[(size_t)i * N + q], Zm[(size_t)i * N + j]);
                    }
                    double mx = 0.0;
                    for (int i = 0; i < N; ++i) mx = fmax(mx, fabs(Zm[(size_t)i * N + j]));
                    if (mx >= dtpcrt && ++nrmchk >= 3) break;
                }
                double nn = 0.0;
                for (int i = 0; i < N; ++i) { const double z = Zm[(size_t)i * N + j]; nn = fma(z, z, nn); }
                const double inv = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
                for (int i = 0; i < N; ++i) Zm[(size_t)i * N + j] *= inv;
            }
        }
        __syncthreads();
        // ---- Loewdin correction over the loose cluster, into the (now free) first factor array: the final Z ----
        if (tid < N) {
            const int j = tid;
            int lo = j, hi = j;
            while (lo > 0 && w[lo] - w[lo - 1] <= ortol) --lo;
            while (hi < N - 1 && w[hi + 1] - w[hi] <= ortol) ++hi;
            for (int i = 0; i < N; ++i) La[(size_t)i * N + j] = Zm[(size_t)i * N + j];
            for (int m = lo; m <= hi; ++m) {
                if (m == j) continue;
                double dot = 0.0;
                for (int i = 0; i < N; ++i) dot = fma(Zm[(size_t)i * N + m], Zm[(size_t)i * N + j], dot);
                const double h = -0.5 * dot;
                for (int i = 0; i < N; ++i) La[(size_t)i * N + j] = fma(h, Zm[(size_t)i * N + m], La[(size_t)i * N + j]);
            }
        }
    }
}

// U = H_0 ... H_{N-2} Z for BW eigenvectors per CTA (one per warp); NQ = ceil(N / 32) components per lane.
constexpr int VBT_PANEL = 8;        // reflectors staged per shared-memory panel
template <int NQ, int BW>
__global__ void __launch_bounds__(BW * 32) vecbt_big_kernel(BigVecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* Vp = reinterpret_cast<cplx*>(smem_raw);                         // [VBT_PANEL][NQ * 32] reflector panel
    cplx* tp = Vp + (size_t)VBT_PANEL * NQ * 32;                          // [VBT_PANEL] tau
    static_assert(BW <= VBT_PANEL, "the staging of finished eigenvectors aliases the reflector panel");
    cplx* ub = Vp;                                                        // [BW][NQ * 32] staging of finished eigenvectors (row selection)
    const int groups = (N + BW - 1) / BW;
    for (long long work = blockIdx.x; work < a.nm * groups; work += gridDim.x) {
        const long long mat = work / groups;
        const int n = (int)(work - mat * groups) * BW + warp;             // this warp's eigenvector
        const bool live = n < N;
        const double* Zm = a.lu + (size_t)mat * 4 * N * N;               // stein_kernel's final (corrected) vectors
        const cplx* Vs = a.vstore + (size_t)mat * N * N;
        const cplx* Ts = a.taustore + (size_t)mat * N;
        cplx u[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = 32 * q + lane;
            u[q] = cmake((live && i < N) ? Zm[(size_t)i * N + n] : 0.0, 0.0);
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
                constexpr int NR = NQ <= 16 ? NQ : 1;                      // NQ = 32: the reflector is re-read for the update (registers)
                cplx vr[NR];
                cplx dot = cmake(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) { const cplx vq = v[32 * q]; if (NQ <= 16) vr[q % NR] = vq; cfmac(dot, vq, u[q]); }      // v^H u
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o); }
                const cplx wv = cmul(tp[s], dot);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (q >= q0) cfms(u[q], wv, NQ <= 16 ? vr[q % NR] : v[32 * q]);       // u -= tau (v^H u) v
                }
            }
        }
        // ---- output: the reference's [band, k, component] layout, or the selected components only ----
        if (a.n_rows == 0) {
            if (live) {
                cplx* out = a.vec_out + (long long)n * a.vs_band + (a.vcol0 + mat) * a.vs_k;
#pragma unroll
                for (int q = 0; q < NQ; ++q) { const int i = 32 * q + lane; if (i < N) out[i] = u[q]; }
            }
        } else {
            __syncthreads();                                               // every warp is done with the last reflector panel
            cplx* mine = ub + (size_t)warp * NQ * 32;
#pragma unroll
            for (int q = 0; q < NQ; ++q) mine[32 * q + lane] = u[q];
            __syncwarp();
            if (live) {
                cplx* out = a.vec_out + (long long)n * a.vs_band + (a.vcol0 + mat) * a.vs_k;
                for (int r = lane; r < a.n_rows; r += 32) out[r] = mine[a.rows[r]];
            }
            __syncwarp();
        }
    }
}

inline size_t vecbt_big_smem(int NQ) { return ((size_t)VBT_PANEL * NQ * 32 + VBT_PANEL) * sizeof(cplx); }
inline size_t bigvec_bytes_per_matrix(int N) {          // Z | lu[4] | pin | reflectors | tau | (d, e)
    const size_t nn = (size_t)N * N;
    return nn * 8 + 4 * nn * 8 + ((nn + 15) & ~(size_t)15) + nn * 16 + (size_t)N * 16 + (size_t)2 * N * 8 + 64;
}

}  // namespace sktb
