// Two-stage tridiagonalisation for large matrices (208 <= N <= 528: eigenvalues, and selected eigenvector rows for LDOS / spectral maps).
//
// The one-stage Householder reduction (kernels_generic.cuh) reads the trailing matrix once per reflector: 16 N^3 / 3 bytes
// per matrix (0.8-0.9 GB at N = 512 against 4 MB of matrix) - it is HBM-bound at 11 % of the FP64 peak.  Here:
//   stage 1  dense -> band of width SB = 16 (LAPACK zhetrd_he2hb idea): QR of a 16-column panel below the band (panel in
//            registers, a warp per column pair and row half), compact-WY block reflector Q = I - V T V^H, two-sided update of the
//            trailing matrix A <- Q^H A Q = A - V W^H - W V^H with X = (A V) T, W = X - 1/2 V (T^H V^H X).  Both matrix-sized
//            operations (Y = A V and the rank-32 update) and the small products are complex GEMMs on the FP64 tensor instruction
//            (mma.sync.m8n8k4.f64 -> DMMA.8x8x4): measured on B200 it shares the DFMA pipe but reaches the full 37 TFLOP/s
//            with 8x fewer issue slots and no register-operand limit (tools/probe/dmma_probe.cu).  The trailing matrix is
//            read twice and written once per PANEL (32 m^2 bytes, lower triangle only): N^3/(3*16)*32 B = 90 MB at N = 512.
//   stage 2  band -> tridiagonal by bulge chasing (zhbtrd / sb2st idea): sweep s annihilates column s of the band and chases
//            the 16x16 bulge down the band; one warp per sweep, sweeps pipelined two hops apart inside a CTA (progress
//            counters in shared memory), the band (2*SB diagonals with the fill, 262 KB at N = 512) stays in L2.
//   then     bisection on the tridiagonal matrix with the product-form Sturm count (bisect_prod_kernel), or - for projections -
//            inverse iteration (stein_kernel) and rows_kernel: rows of U = Q1 Q2 Z from the reflectors both stages keep.
// Every reflector follows zlarfg (H^H x = beta e1, H = I - tau v v^H, beta real), the band layout is AB[j][d] = A[j+d][j].
// tools/proto_twostage.py is the numpy model of both stages (same blocking, conventions and index arithmetic).
// Replaces, for large matrices, the LAPACK call behind numpy.linalg.eigh at pysktb/hamiltonian.py:119.
#pragma once
#include "common.cuh"

namespace sktb {
namespace twostage {

constexpr int SB = 16;            // band width after stage 1
constexpr int ABW = 2 * SB;       // stored diagonals per column (d = 0 .. 2 SB - 1: the bulge chase fills up to 2 SB - 1)
constexpr int S1_WARPS = 16;
constexpr int S1_THREADS = S1_WARPS * 32;

constexpr int YP_MAX = 8;        // column parts of the X = A V product (load balance; partial sums go through global memory)

struct Stage1Args {
    cplx* H; int N; long long nm;   // [nm][N][N] row-major, lower triangle read (zheevd UPLO='L'), destroyed
    cplx* AB;                       // [nm][N][ABW]
    cplx* Yp;                       // [nm][YP_MAX][NR][SB]: partial products A22 V per column part
    cplx* Wg;                       // [nm][NR][SB]: X = A22 V T, then W
    cplx* Gg;                       // [nm][S1_WARPS][SB*SB]: per-warp partial sums of G = V^H X
    int NR, mp;                     // NR = N rounded up to 8; mp: column stride of the panel in shared memory (odd)
    cplx* Tstore;                   // NULL, or [nm][npanels][SB*SB]: keep the block reflectors (V goes into the dead panel of H, T here)
    int npanels;
};
struct Stage2Args {
    cplx* AB; int N; long long nm;
    double* de;                     // [nm][2][N]: d, then e (N - 1 entries)
    cplx* Rstore; long long rstride;   // NULL, or [nm][rstride]: reflector k = off(s) + h of (sweep s, hop h) as 16 entries of v + tau
};
// number of bulge-chase reflectors of an N x N band: sum over sweeps of ceil((N - 1 - s) / SB)
SKTB_HD long long chase_reflectors(int N) { long long k = 0; for (int s = 0; s < N - 1; ++s) k += (N - 1 - s + SB - 1) / SB; return k; }
SKTB_HD int stage1_panels(int N) { return N > SB ? (N - SB + SB - 1) / SB : 0; }
// selected rows of the eigenvector matrix U = Q1 Q2 Z (Q1: block reflectors of stage 1, Q2: chase reflectors, Z: tridiagonal eigenvectors)
struct RowsArgs {
    int N; long long nm;
    const cplx* H;                  // [nm][N][N] after stage 1 with Tstore: V panels below the band
    const cplx* Tstore; int npanels;
    const cplx* Rstore; long long rstride;
    const double* Zf; long long zstride;      // Zf[mat * zstride + k * N + n]: component k of tridiagonal eigenvector n
    cplx* vec_out; long long vs_band, vs_k, vcol0;   // vec_out[n * vs_band + (vcol0 + mat) * vs_k + r] = U[rows[r], n]
    int n_rows; const int* rows;
};

SKTB_HD int stage1_mp(int N) { return (((N - SB + 7) / 8) * 8) | 1; }
SKTB_HD size_t stage1_smem_bytes(int N, int ring_depth) {
    return ((size_t)(SB + 2) * stage1_mp(N) + 4 * SB * SB + 5 * SB + (size_t)S1_WARPS * ring_depth * 64) * sizeof(cplx) + 64;
}
SKTB_HD size_t scratch_bytes_per_matrix(int N) {
    const size_t NR = (size_t)(N + 7) / 8 * 8;
    return ((size_t)N * ABW + (YP_MAX + 1) * NR * SB + (size_t)S1_WARPS * SB * SB) * sizeof(cplx) + (size_t)2 * N * sizeof(double);
}

// host: stage 1 + stage 2 on `nm` matrices; de[mat][2][N] receives the tridiagonal.  scratch: nm * scratch_bytes_per_matrix(N).
// Returns the number of kernel launches, or -1 when N does not fit (shared memory) - the caller falls back to the one-stage path.
// stages: bit 0 = stage 1 (H -> band in scratch), bit 1 = stage 2 (band -> de).  band_of(scratch) is the [nm][N][ABW] band array.
int launch_tridiag(cplx* H, int N, long long nm, void* scratch, double* de, cudaStream_t st, char* info, size_t info_len, int stages = 3);
inline cplx* band_of(void* scratch) { return reinterpret_cast<cplx*>(scratch); }
bool supported(int N);
// eigenvalues of the tridiagonal matrices de[mat][2][N] by bisection (product-form Sturm sequence), ascending: evals[n*ld + col0 + mat]
int launch_bisect(const double* de, int N, long long nm, double* evals, long long ld, long long col0, cudaStream_t st);
// keep variant: Tstore / Rstore (see Stage1Args / Stage2Args) are filled for launch_rows
int launch_tridiag_keep(cplx* H, int N, long long nm, void* scratch, double* de, cplx* Tstore, cplx* Rstore, long long rstride, cudaStream_t st, char* info, size_t info_len);
int launch_rows(const RowsArgs& a, cudaStream_t st);

#if defined(__CUDACC__) && defined(SKTB_TWOSTAGE_KERNELS)      // the kernels live in heev_twostage.cu only
// D(8x8) += A(8x4) B(4x8) in FP64; lane = 4 g + q: a = A[g][q], b = B[q][g], c0/c1 = C[g][2q], C[g][2q+1]
SKTB_D void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
SKTB_D double dneg(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
SKTB_D cplx ldcg(const cplx* p) { return __ldcg(p); }
SKTB_D cplx cwarp_sum(cplx v) { return cmake(warp_sum(v.x), warp_sum(v.y)); }

// zlarfg without the safmin loop, reciprocal square root / reciprocal by hardware seed + Newton (full double precision)
SKTB_D void fast_householder(cplx alpha, double xnorm2, double& beta, cplx& tau, cplx& scale) {
    if (xnorm2 == 0.0 && alpha.y == 0.0) { beta = alpha.x; tau = cmake(0.0, 0.0); scale = cmake(0.0, 0.0); return; }
    const double n2 = alpha.x * alpha.x + alpha.y * alpha.y + xnorm2;
    const double irt = fast_rsqrt(n2);
    const double nrm = n2 * irt;
    const bool neg = alpha.x >= 0.0;
    beta = neg ? -nrm : nrm;
    const double ib = neg ? -irt : irt;                                   // 1 / beta
    tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
    const double dx = alpha.x - beta, dy = alpha.y;
    const double id = fast_rcp(dx * dx + dy * dy);
    scale = cmake(dx * id, -dy * id);
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------------------------------------
// X tile (8 rows x 16 columns, complex) += A-operand (two k-steps a0, a1; CONJ: use conj(a)) times the B-operand rows
// 8J+2q, 8J+2q+1 of V (shared memory, column-major: V[row][n] = Ps[n*mp + row])
template <bool CONJ>
SKTB_D void s1_xtile(double (&Xr)[2][2], double (&Xi)[2][2], double (&Zr)[2][2], double (&Zi)[2][2], cplx a0, cplx a1,
                     const cplx* __restrict__ Ps, int mp, int rowB, int g) {
    // X = Xr + Zr + i (Xi + Zi): the two k-steps accumulate into separate registers (8 independent DMMA chains per warp)
    const cplx* bp = Ps + (size_t)g * mp + rowB;
    const cplx b00 = bp[0], b10 = bp[1], b01 = bp[(size_t)8 * mp], b11 = bp[(size_t)8 * mp + 1];
    const double a0i = CONJ ? a0.y : dneg(a0.y), a0j = CONJ ? dneg(a0.y) : a0.y;     // Xr += a.x b.x + a0i b.y ; Xi += a.x b.y + a0j b.x
    const double a1i = CONJ ? a1.y : dneg(a1.y), a1j = CONJ ? dneg(a1.y) : a1.y;
    dmma(Xr[0][0], Xr[0][1], a0.x, b00.x); dmma(Xi[0][0], Xi[0][1], a0.x, b00.y);
    dmma(Xr[1][0], Xr[1][1], a0.x, b01.x); dmma(Xi[1][0], Xi[1][1], a0.x, b01.y);
    dmma(Zr[0][0], Zr[0][1], a1.x, b10.x); dmma(Zi[0][0], Zi[0][1], a1.x, b10.y);
    dmma(Zr[1][0], Zr[1][1], a1.x, b11.x); dmma(Zi[1][0], Zi[1][1], a1.x, b11.y);
    dmma(Xr[0][0], Xr[0][1], a0i, b00.y);  dmma(Xi[0][0], Xi[0][1], a0j, b00.x);
    dmma(Xr[1][0], Xr[1][1], a0i, b01.y);  dmma(Xi[1][0], Xi[1][1], a0j, b01.x);
    dmma(Zr[0][0], Zr[0][1], a1i, b10.y);  dmma(Zi[0][0], Zi[0][1], a1j, b10.x);
    dmma(Zr[1][0], Zr[1][1], a1i, b11.y);  dmma(Zi[1][0], Zi[1][1], a1j, b11.x);
}

// asynchronous 16-byte copy global -> shared (LDGSTS), zero-filled when !ok; per-thread commit groups
SKTB_D void cp_async16(cplx* smem_dst, const cplx* gsrc, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
SKTB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
SKTB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// rank-2k tile update: C -= V_I W_J^H + W_I V_J^H accumulated straight into the C fragment (Cr, Ci) - the signs go into the B operands (sign-bit
// flips on the integer pipe), so the tile has no FP64 epilogue: with the pipe full of DMMAs every DADD of an epilogue waits in line behind them
// (ncu: a quarter of the update phase was 'math pipe throttle' on 12 DADD per tile).
//   Re: Cr += Vr (-Wr) + Vi (-Wi) + Wr (-Vr) + Wi (-Vi)      Im: Ci += Vi (-Wr) + Vr Wi + Wi (-Vr) + Wr Vi
SKTB_D void s1_r2k(double (&Cr)[2], double (&Ci)[2], cplx vI, cplx wI, cplx v2, cplx w2) {
    const double nwx = dneg(w2.x), nwy = dneg(w2.y), nvx = dneg(v2.x), nvy = dneg(v2.y);
    dmma(Cr[0], Cr[1], vI.x, nwx); dmma(Ci[0], Ci[1], vI.y, nwx);
    dmma(Cr[0], Cr[1], vI.y, nwy); dmma(Ci[0], Ci[1], vI.x, w2.y);
    dmma(Cr[0], Cr[1], wI.x, nvx); dmma(Ci[0], Ci[1], wI.y, nvx);
    dmma(Cr[0], Cr[1], wI.y, nvy); dmma(Ci[0], Ci[1], wI.x, v2.y);
}
#ifdef SKTB_TS_PROFILE
__device__ unsigned long long g_s1_prof[32];
#define S1_MARK(i) { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_s1_prof[i], (unsigned long long)(now_ - tmark_)); tmark_ = now_; } }
#define S1_WMARK(i) { if (lane == 0) { const long long now_ = clock64(); atomicAdd(&g_s1_prof[i], (unsigned long long)(now_ - wmark_)); } }
#define S1_WSTART() { wmark_ = clock64(); }
#else
#define S1_MARK(i)
#define S1_WMARK(i)
#define S1_WSTART()
#endif
template <int RD>      // depth of the per-warp ring of asynchronously prefetched 8 x 8 tiles (shared memory, 1 KB per tile)
__global__ void __launch_bounds__(S1_THREADS, 1) stage1_kernel(Stage1Args a) {
#ifdef SKTB_TS_PROFILE
    long long tmark_ = clock64(), wmark_ = 0;
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, mp = a.mp, NR = a.NR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    cplx* Ps = reinterpret_cast<cplx*>(smem_raw);            // [SB][mp] panel, column-major; V after the QR
    cplx* Ts = Ps + (size_t)SB * mp;                         // [SB][SB] T (upper triangular), Ts[k*SB+n]
    cplx* Rs = Ts + SB * SB;                                 // [SB][SB] R, Rs[r*SB+c]
    cplx* Ss = Rs + SB * SB;                                 // [SB][SB] G, then 1/2 S = 1/2 T^H G
    cplx* Zs = Ss + SB * SB;                                 // [SB][SB] Zs[w*SB+c] = V_w^H v_c (w < c), for T
    cplx* part = Zs + SB * SB;                               // [2][SB] raw dots with the pivot column, one per row half
    cplx* rowc = part + 2 * SB;                              // [2][SB] entries of the current pivot row of every column (double-buffered)
    cplx* taus = rowc + 2 * SB;                              // [SB]
    cplx* shadow = taus + SB;                                // [2][mp] the current pivot column, updated and unscaled
    cplx* ring = shadow + (size_t)2 * mp + (size_t)warp * RD * 64 + lane;   // this lane's slots: ring[(stage * 2 + e) * 32]

    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        cplx* A = a.H + (size_t)mat * N * N;
        cplx* AB = a.AB + (size_t)mat * N * ABW;
        cplx* Yp = a.Yp + (size_t)mat * YP_MAX * NR * SB;
        cplx* Wg = a.Wg + (size_t)mat * NR * SB;
        cplx* Gg = a.Gg + (size_t)mat * S1_WARPS * SB * SB;
        int c0 = 0;
        for (;; c0 += SB) {
            const int m = N - c0 - SB;                       // rows below the band of this panel
            if (m <= 0) break;
            const int mr = (m + 7) / 8 * 8;
            const int nt = mr / 8;
            cplx* A22 = A + (size_t)(c0 + SB) * N + (c0 + SB);
            // ---- 1 + 2. Householder QR of the panel with the panel in REGISTERS.  Warp (k2 = warp >> 1, h = warp & 1) owns columns 2 k2,
            //         2 k2 + 1 on the rows r = 64 i + 32 h + lane (i < 8): 16 complex per lane.  Per column c: the pivot column (shared
            //         memory, written by its owners in the previous step) is read ONCE into registers, the two dots of the own columns
            //         with it go through part[h][w]; after the barrier every warp forms the reflector, updates its columns from the cached
            //         pivot and publishes the entries of row c+1 (rowc) and, if it owns it, the next pivot column.  (With the panel in shared
            //         memory every column step streams the whole panel through the 128 B/clk pipe: 5 kclk per column, 17 % of the kernel.) ----
            __syncthreads();
            S1_MARK(0)
            const int k2 = warp >> 1, hq = warp & 1, w0 = 2 * k2;
            cplx col0[8], col1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 64 * i + 32 * hq + lane;
                col0[i] = col1[i] = cmake(0.0, 0.0);
                if (r < m) { const cplx* ap = A + (size_t)(c0 + SB + r) * N + c0 + w0; col0[i] = ldcg(ap); col1[i] = ldcg(ap + 1); }
            }
            for (int idx = tid; idx < 4 * SB * SB; idx += S1_THREADS) Ts[idx] = cmake(0.0, 0.0);   // T, R, S, Z
            if (k2 == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const int r = 64 * i + 32 * hq + lane; if (r < m) shadow[r] = col0[i]; }
            }
            if (hq == 0 && lane == 0) { rowc[w0] = col0[0]; rowc[w0 + 1] = col1[0]; }
            __syncthreads();
            S1_MARK(1)
            const int ncol = m < SB ? m : SB;                 // columns that have a diagonal row
            const int ns = (m + 63) >> 6;                     // live 64-row slots
            for (int c = 0; c < ncol; ++c) {
                const cplx* pc = shadow + (size_t)(c & 1) * mp;
                cplx pcr[8];
                cplx d0 = cmake(0.0, 0.0), d1 = cmake(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 64 * i + 32 * hq + lane;
                    pcr[i] = cmake(0.0, 0.0);
                    if (i < ns) {                                               // (uniform) slots beyond the panel carry no work
                        if (r > c && r < m) pcr[i] = pc[r];
                        cfmac(d0, pcr[i], col0[i]); cfmac(d1, pcr[i], col1[i]);
                    }
                }
                d0 = cwarp_sum(d0); d1 = cwarp_sum(d1);
                if (lane == 0) { part[hq * SB + w0] = d0; part[hq * SB + w0 + 1] = d1; }
                S1_MARK(15)
                __syncthreads();
                S1_MARK(12)
                double beta; cplx tau, scale;
                fast_householder(pc[c], part[c].x + part[SB + c].x, beta, tau, scale);
                const cplx ctau = cconj(tau);
                const cplx* rc = rowc + (c & 1) * SB;         // entries of row c of every column (before this step)
                cplx* rn = rowc + ((c + 1) & 1) * SB;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int w = w0 + j;
                    cplx* col = j ? col1 : col0;
                    const cplx raw = cadd(part[w], part[SB + w]);
                    const cplx ac = rc[w];
                    if (w > c) {
                        const cplx coef = cmul(ctau, cadd(ac, cmulc(scale, raw)));     // conj(tau) v^H a_w
                        const cplx cs2 = cmul(coef, scale);
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (i < ns) cfms(col[i], cs2, pcr[i]);
                        if (hq == 0 && lane == 0) Rs[c * SB + w] = csub(ac, coef);
                    } else if (w == c) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = 64 * i + 32 * hq + lane;
                            col[i] = r == c ? cmake(1.0, 0.0) : cmul(pcr[i], scale);   // rows < c: pcr = 0
                        }
                        if (hq == 0 && lane == 0) { taus[c] = tau; Rs[c * SB + c] = cmake(beta, 0.0); }
                    } else {
                        if (hq == 0 && lane == 0) Zs[w * SB + c] = cadd(cconj(ac), cmul(scale, cconj(raw)));   // V_w^H v_c
                    }
                }
                // publish row c+1 of the own columns and, for its owners, the next pivot column
                if (c + 1 < ncol) {
                    const int rr = c + 1;
                    if (((rr >> 5) & 1) == hq && (rr & 31) == lane) {
                        const int ii = rr >> 6;
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (i == ii) { rn[w0] = col0[i]; rn[w0 + 1] = col1[i]; }
                    }
                    if ((rr >> 1) == k2) {
                        cplx* nx = shadow + (size_t)(rr & 1) * mp;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = 64 * i + 32 * hq + lane;
                            if (r > c && r < m) nx[r] = (rr & 1) ? col1[i] : col0[i];
                        }
                    }
                }
                S1_MARK(16)
                __syncthreads();
                S1_MARK(13)
            }
            // V -> shared memory (column-major) for the GEMM operands; columns without a diagonal row (m < SB) and rows m .. mr-1 are zero
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 64 * i + 32 * hq + lane;
                if (r < mr) {
                    const bool live = r < m;
                    const cplx v0 = (live && w0 < ncol) ? col0[i] : cmake(0.0, 0.0), v1 = (live && w0 + 1 < ncol) ? col1[i] : cmake(0.0, 0.0);
                    Ps[(size_t)w0 * mp + r] = v0;
                    Ps[(size_t)(w0 + 1) * mp + r] = v1;
                    if (a.Tstore && live) { cplx* vp = A + (size_t)(c0 + SB + r) * N + c0 + w0; vp[0] = v0; vp[1] = v1; }   // V into the dead panel of H
                }
            }
            __syncthreads();
            // T[t][c] = -tau_c sum_{k=t}^{c-1} T[t][k] z[k][c], T[t][t] = tau_t: the rows of T are independent - one half-warp per row
            // (lane k holds T[t][k]), 15 uniform column steps with a 16-lane reduction each
            if (warp < SB / 2) {
                const int t = 2 * warp + (lane >> 4), k = lane & 15;
                cplx Tk = (k == t && t < ncol) ? taus[t] : cmake(0.0, 0.0);
                for (int c = 1; c < ncol; ++c) {
                    cplx pr = cmake(0.0, 0.0);
                    if (k >= t && k < c) pr = cmul(Tk, Zs[k * SB + c]);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) { pr.x += __shfl_xor_sync(0xffffffffu, pr.x, o); pr.y += __shfl_xor_sync(0xffffffffu, pr.y, o); }
                    if (k == c && c > t) { const cplx tc = taus[c]; Tk = cmul(cmake(-tc.x, -tc.y), pr); }
                }
                Ts[t * SB + k] = Tk;
                if (a.Tstore) a.Tstore[((size_t)mat * a.npanels + c0 / SB) * SB * SB + t * SB + k] = Tk;
            }
            S1_MARK(2)
            // ---- 3. band output of the 16 columns of this panel: diagonal block (lower) + R; the rest of the column is zero ----
            {
                const int c = tid >> 5, d = tid & 31;
                const int j = c0 + c;
                cplx val = cmake(0.0, 0.0);
                if (d <= SB - 1 - c) {
                    val = ldcg(A + (size_t)(j + d) * N + j);
                    if (d == 0) val.y = 0.0;
                } else if (d >= SB - c && d <= SB) {
                    const int r = d - SB + c;
                    if (r < m) val = Rs[r * SB + c];
                }
                AB[(size_t)j * ABW + d] = val;
            }
            S1_MARK(3)
            S1_WSTART()
            // ---- 4. Y = A22 V (lower triangle of A22 read: tile (I,J) for J <= I, tile (J,I)^H for J > I).  Work units: (row tile
            //         I, column part p) dealt round-robin to the warps; P parts so that the last round is nearly full ----
            int P = 1;
            {
                int best = 1 << 30;
                for (int pp = 1; pp <= YP_MAX && pp <= nt; ++pp) {
                    const int units = nt * pp, rounds = (units + S1_WARPS - 1) / S1_WARPS;
                    const int idle = (rounds * S1_WARPS - units) * 1024 / (rounds * S1_WARPS) + pp * 8;   // idle share (x1024) + a price per part
                    if (idle < best) { best = idle; P = pp; }
                }
            }
            {
                // one stream of tiles per warp across its units; the tiles are copied to the warp's ring RD - 1 tiles ahead (cp.async)
                const int nunits = nt * P;
                int uI = warp, iI = 0, pI = 0, jI = 0, j1I = 0;                 // issue cursor
                int uC = warp, iC = 0, pC = 0, jC = 0, j1C = 0;                 // compute cursor
                auto unit_of = [&](int u, int& I, int& p, int& J, int& J1) {
                    I = u / P; p = u - I * P;
                    J = (int)((long long)p * nt / P); J1 = (int)((long long)(p + 1) * nt / P);
                };
                if (uI < nunits) { unit_of(uI, iI, pI, jI, j1I); unit_of(uC, iC, pC, jC, j1C); }
                auto issue = [&](int stage) {
                    if (uI < nunits) {
                        const int I = iI, J = jI;
                        const int row = 8 * I + g;
                        cplx* dst = ring + stage * 64;
                        if (J < I) {                                            // A22[row][8J + 2q + e]
                            const cplx* ap = A22 + (size_t)row * N + 8 * J + 2 * q;
                            const bool ok = row < m;
                            cp_async16(dst, ok ? ap : A22, ok); cp_async16(dst + 32, ok ? ap + 1 : A22, ok);
                        } else if (J > I) {                                     // A22[8J + 2q + e][8I + g] (conjugated at use)
                            const int r0 = 8 * J + 2 * q;
                            const cplx* ap = A22 + (size_t)r0 * N + row;
                            const bool ok0 = r0 < m, ok1 = r0 + 1 < m;
                            cp_async16(dst, ok0 ? ap : A22, ok0); cp_async16(dst + 32, ok1 ? ap + N : A22, ok1);
                        } else {                                                // diagonal tile: the stored (lower) one of (row, col), (col, row)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int col = 8 * I + 2 * q + e;
                                const bool ok = row < m && col < m;
                                const cplx* ap = row >= col ? A22 + (size_t)row * N + col : A22 + (size_t)col * N + row;
                                cp_async16(dst + 32 * e, ok ? ap : A22, ok);
                            }
                        }
                        if (++jI >= j1I) { uI += S1_WARPS; if (uI < nunits) unit_of(uI, iI, pI, jI, j1I); }
                    }
                    cp_async_commit();
                };
#pragma unroll
                for (int k = 0; k < RD - 1; ++k) issue(k);
                double Xr[2][2] = {{0, 0}, {0, 0}}, Xi[2][2] = {{0, 0}, {0, 0}}, Zr[2][2] = {{0, 0}, {0, 0}}, Zi[2][2] = {{0, 0}, {0, 0}};
                int stage = 0;
                while (uC < nunits) {
                    issue(stage == 0 ? RD - 1 : stage - 1);
                    cp_async_wait<RD - 1>();
                    cplx a0 = ring[stage * 64], a1 = ring[stage * 64 + 32];
                    const int I = iC, J = jC;
                    if (J > I) { a0.y = dneg(a0.y); a1.y = dneg(a1.y); }
                    else if (J == I) {
                        const int row = 8 * I + g, col = 8 * I + 2 * q;
                        if (row < col) a0.y = dneg(a0.y); else if (row == col) a0.y = 0.0;
                        if (row < col + 1) a1.y = dneg(a1.y); else if (row == col + 1) a1.y = 0.0;
                    }
                    s1_xtile<false>(Xr, Xi, Zr, Zi, a0, a1, Ps, mp, 8 * J + 2 * q, g);
                    if (++jC >= j1C) {
                        // partial Y tile -> Yp[p] (C layout: row g, columns 8t + 2q, 8t + 2q + 1)
                        cplx* xp = Yp + ((size_t)pC * NR + 8 * I + g) * SB + 2 * q;
                        xp[0] = cmake(Xr[0][0] + Zr[0][0], Xi[0][0] + Zi[0][0]); xp[1] = cmake(Xr[0][1] + Zr[0][1], Xi[0][1] + Zi[0][1]);
                        xp[8] = cmake(Xr[1][0] + Zr[1][0], Xi[1][0] + Zi[1][0]); xp[9] = cmake(Xr[1][1] + Zr[1][1], Xi[1][1] + Zi[1][1]);
#pragma unroll
                        for (int t2 = 0; t2 < 2; ++t2) { Xr[t2][0] = Xr[t2][1] = Xi[t2][0] = Xi[t2][1] = Zr[t2][0] = Zr[t2][1] = Zi[t2][0] = Zi[t2][1] = 0.0; }
                        uC += S1_WARPS;
                        if (uC < nunits) unit_of(uC, iC, pC, jC, j1C);
                    }
                    stage = stage + 1 == RD ? 0 : stage + 1;
                }
                cp_async_wait<0>();
            }
            S1_WMARK(10)
            __syncthreads();
            S1_MARK(4)
            // ---- 5. X = (sum of the parts) T: 8 x 16 row tiles on DMMA, k index of (lane q, k-step s) = 2q + (s & 1) + 8 (s >> 1) ----
            for (int I = warp; I < nt; I += S1_WARPS) {
                const int row = 8 * I + g;
                cplx y[4];
                {
                    const cplx* yp = Yp + (size_t)row * SB + 2 * q;
                    y[0] = yp[0]; y[1] = yp[1]; y[2] = yp[8]; y[3] = yp[9];
                    for (int p = 1; p < P; ++p) {
                        const cplx* y2 = yp + (size_t)p * NR * SB;
                        y[0] = cadd(y[0], y2[0]); y[1] = cadd(y[1], y2[1]); y[2] = cadd(y[2], y2[8]); y[3] = cadd(y[3], y2[9]);
                    }
                }
                double Xr[2][2] = {{0, 0}, {0, 0}}, Xi[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int k = 2 * q + (s4 & 1) + 8 * (s4 >> 1);
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        if (s4 >= 2 && t2 == 0) continue;                          // T is upper triangular: rows 8.. of columns ..7 are zero
                        const cplx tb = Ts[k * SB + g + 8 * t2];
                        dmma(Xr[t2][0], Xr[t2][1], y[s4].x, tb.x); dmma(Xi[t2][0], Xi[t2][1], y[s4].x, tb.y);
                        dmma(Xr[t2][0], Xr[t2][1], dneg(y[s4].y), tb.y); dmma(Xi[t2][0], Xi[t2][1], y[s4].y, tb.x);
                    }
                }
                cplx* xp = Wg + (size_t)row * SB + 2 * q;
                xp[0] = cmake(Xr[0][0], Xi[0][0]); xp[1] = cmake(Xr[0][1], Xi[0][1]);
                xp[8] = cmake(Xr[1][0], Xi[1][0]); xp[9] = cmake(Xr[1][1], Xi[1][1]);
            }
            __syncthreads();
            S1_MARK(5)
            // ---- 6. G = V^H X (16x16), per-warp partial sums over the warp's rows, then S = T^H G ----
            {
                double Gr[2][2][2], Gi[2][2][2];
#pragma unroll
                for (int t1 = 0; t1 < 2; ++t1)
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) { Gr[t1][t2][0] = Gr[t1][t2][1] = Gi[t1][t2][0] = Gi[t1][t2][1] = 0.0; }
                for (int kk = warp; kk < mr / 8; kk += S1_WARPS) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int rk = 8 * kk + 2 * q + e;                  // rows 2q + e: conflict-free with the odd column stride
                        const cplx vv[2] = {Ps[(size_t)g * mp + rk], Ps[(size_t)(g + 8) * mp + rk]};
                        const cplx xx[2] = {Wg[(size_t)rk * SB + g], Wg[(size_t)rk * SB + g + 8]};
#pragma unroll
                        for (int t1 = 0; t1 < 2; ++t1)
#pragma unroll
                            for (int t2 = 0; t2 < 2; ++t2) {
                                // conj(v) x: re = vr xr + vi xi, im = vr xi - vi xr
                                dmma(Gr[t1][t2][0], Gr[t1][t2][1], vv[t1].x, xx[t2].x); dmma(Gr[t1][t2][0], Gr[t1][t2][1], vv[t1].y, xx[t2].y);
                                dmma(Gi[t1][t2][0], Gi[t1][t2][1], vv[t1].x, xx[t2].y); dmma(Gi[t1][t2][0], Gi[t1][t2][1], dneg(vv[t1].y), xx[t2].x);
                            }
                    }
                }
                cplx* gp = Gg + (size_t)warp * SB * SB;
#pragma unroll
                for (int t1 = 0; t1 < 2; ++t1)
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        gp[(8 * t1 + g) * SB + 8 * t2 + 2 * q] = cmake(Gr[t1][t2][0], Gi[t1][t2][0]);
                        gp[(8 * t1 + g) * SB + 8 * t2 + 2 * q + 1] = cmake(Gr[t1][t2][1], Gi[t1][t2][1]);
                    }
            }
            __syncthreads();
            if (tid < SB * SB) {
                cplx s = cmake(0.0, 0.0);
                for (int w = 0; w < S1_WARPS; ++w) s = cadd(s, Gg[(size_t)w * SB * SB + tid]);
                Ss[tid] = s;
            }
            __syncthreads();
            cplx sval = cmake(0.0, 0.0);
            if (tid < SB * SB) {                              // S[i][j] = sum_{k<=i} conj(T[k][i]) G[k][j]
                const int i = tid >> 4, j = tid & 15;
                for (int k = 0; k <= i; ++k) cfmac(sval, Ts[k * SB + i], Ss[k * SB + j]);
            }
            __syncthreads();
            if (tid < SB * SB) Ss[tid] = cscale(0.5, sval);   // 1/2 S
            __syncthreads();
            S1_MARK(6)
            // ---- 7. W = X - V (1/2 S) on DMMA, same tiling as step 5 ----
            for (int I = warp; I < nt; I += S1_WARPS) {
                const int row = 8 * I + g;
                cplx* xp = Wg + (size_t)row * SB + 2 * q;
                const cplx x00 = xp[0], x01 = xp[1], x10 = xp[8], x11 = xp[9];
                double Xr[2][2] = {{x00.x, x01.x}, {x10.x, x11.x}}, Xi[2][2] = {{x00.y, x01.y}, {x10.y, x11.y}};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int k = 2 * q + (s4 & 1) + 8 * (s4 >> 1);
                    const cplx v = Ps[(size_t)k * mp + row];
                    const double nvx = dneg(v.x), nvy = dneg(v.y);
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        const cplx sb = Ss[k * SB + g + 8 * t2];
                        // X -= v s: re -= vx sx - vy sy, im -= vx sy + vy sx
                        dmma(Xr[t2][0], Xr[t2][1], nvx, sb.x); dmma(Xi[t2][0], Xi[t2][1], nvx, sb.y);
                        dmma(Xr[t2][0], Xr[t2][1], v.y, sb.y); dmma(Xi[t2][0], Xi[t2][1], nvy, sb.x);
                    }
                }
                xp[0] = cmake(Xr[0][0], Xi[0][0]); xp[1] = cmake(Xr[0][1], Xi[0][1]);
                xp[8] = cmake(Xr[1][0], Xi[1][0]); xp[9] = cmake(Xr[1][1], Xi[1][1]);
            }
            __syncthreads();
            S1_MARK(7)
            S1_WSTART()
            // ---- 8. A22 -= V W^H + W V^H on the lower tiles; the linearised tile list is cut into 16 equal ranges ----
            {
                const int T = nt * (nt + 1) / 2;
                int t = (int)((long long)warp * T / S1_WARPS);
                const int t1 = (int)((long long)(warp + 1) * T / S1_WARPS);
                int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((I + 1) * (I + 2) / 2 <= t) ++I;
                while (I * (I + 1) / 2 > t) --I;
                int J = t - I * (I + 1) / 2;
                // C tiles: cp.async ring, RD - 1 tiles ahead; W_J operands: registers, one tile ahead (two buffers, no register rotation)
                int Ic = I, Jc = J, tc = t;                              // issue cursor of the C ring
                auto issue_c = [&](int stage) {
                    if (tc < t1) {
                        const int row = 8 * Ic + g, col = 8 * Jc + 2 * q;
                        const cplx* cp = A22 + (size_t)row * N + col;
                        const bool ok0 = row < m && col < m, ok1 = row < m && col + 1 < m;
                        cplx* dst = ring + stage * 64;
                        cp_async16(dst, ok0 ? cp : A22, ok0); cp_async16(dst + 32, ok1 ? cp + 1 : A22, ok1);
                        ++tc; if (++Jc > Ic) { ++Ic; Jc = 0; }
                    }
                    cp_async_commit();
                };
#pragma unroll
                for (int k = 0; k < RD - 1; ++k) issue_c(k);
                int stage = 0;
                int In = I, Jn = J;                                      // position of the W operands to prefetch next
                cplx wA0, wA1, wA2, wA3, wB0, wB1, wB2, wB3, wC0, wC1, wC2, wC3;      // three buffers: the W operands are loaded two tiles ahead
                wA0 = wA1 = wA2 = wA3 = wB0 = wB1 = wB2 = wB3 = wC0 = wC1 = wC2 = wC3 = cmake(0.0, 0.0);
                auto load_w = [&](int Jj, cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
                    const cplx* wj = Wg + (size_t)(8 * Jj + g) * SB + 2 * q;      // k index of (q, s) = 2q + (s & 1) + 8 (s >> 1)
                    x0 = wj[0]; x1 = wj[1]; x2 = wj[8]; x3 = wj[9];
                };
                if (t < t1) { load_w(Jn, wA0, wA1, wA2, wA3); if (++Jn > In) { ++In; Jn = 0; } }
                if (t + 1 < t1) { load_w(Jn, wB0, wB1, wB2, wB3); if (++Jn > In) { ++In; Jn = 0; } }
                int Irow = -1;
                cplx vI0, vI1, vI2, vI3, wI0, wI1, wI2, wI3;
                vI0 = vI1 = vI2 = vI3 = wI0 = wI1 = wI2 = wI3 = cmake(0.0, 0.0);
#define SKTB_S1_TILE(W0, W1, W2, W3)                                                                                       \
                    {                                                                                                      \
                        if (I != Irow) {                                                                                   \
                            const int rw = 8 * I + g;                                                                      \
                            vI0 = Ps[(size_t)(2 * q) * mp + rw]; vI1 = Ps[(size_t)(2 * q + 1) * mp + rw];                  \
                            vI2 = Ps[(size_t)(2 * q + 8) * mp + rw]; vI3 = Ps[(size_t)(2 * q + 9) * mp + rw];              \
                            const cplx* wp = Wg + (size_t)rw * SB + 2 * q;                                                 \
                            wI0 = wp[0]; wI1 = wp[1]; wI2 = wp[8]; wI3 = wp[9];                                            \
                            Irow = I;                                                                                      \
                        }                                                                                                  \
                        issue_c(stage == 0 ? RD - 1 : stage - 1);                                                          \
                        cp_async_wait<RD - 1>();                                                                           \
                        const cplx C0 = ring[stage * 64], C1 = ring[stage * 64 + 32];                                      \
                        stage = stage + 1 == RD ? 0 : stage + 1;                                                           \
                        const int row = 8 * I + g, col = 8 * J + 2 * q;                                                    \
                        double Cr[2] = {C0.x, C1.x}, Ci[2] = {C0.y, C1.y};                                                 \
                        const cplx* vp = Ps + (size_t)(2 * q) * mp + 8 * J + g;                                            \
                        s1_r2k(Cr, Ci, vI0, wI0, vp[0], W0);                                                               \
                        s1_r2k(Cr, Ci, vI1, wI1, vp[mp], W1);                                                              \
                        s1_r2k(Cr, Ci, vI2, wI2, vp[(size_t)8 * mp], W2);                                                  \
                        s1_r2k(Cr, Ci, vI3, wI3, vp[(size_t)9 * mp], W3);                                                  \
                        cplx* cp = A22 + (size_t)row * N + col;                                                            \
                        if (row < m && col < m) cp[0] = cmake(Cr[0], Ci[0]);                                               \
                        if (row < m && col + 1 < m) cp[1] = cmake(Cr[1], Ci[1]);                                           \
                        if (++J > I) { ++I; J = 0; }                                                                       \
                    }
                for (; t < t1; t += 3) {
                    if (t + 2 < t1) { load_w(Jn, wC0, wC1, wC2, wC3); if (++Jn > In) { ++In; Jn = 0; } }
                    SKTB_S1_TILE(wA0, wA1, wA2, wA3)
                    if (t + 1 < t1) {
                        if (t + 3 < t1) { load_w(Jn, wA0, wA1, wA2, wA3); if (++Jn > In) { ++In; Jn = 0; } }
                        SKTB_S1_TILE(wB0, wB1, wB2, wB3)
                    }
                    if (t + 2 < t1) {
                        if (t + 4 < t1) { load_w(Jn, wB0, wB1, wB2, wB3); if (++Jn > In) { ++In; Jn = 0; } }
                        SKTB_S1_TILE(wC0, wC1, wC2, wC3)
                    }
                }
                cp_async_wait<0>();
#undef SKTB_S1_TILE
                S1_WMARK(11)
            }
            S1_MARK(8)
        }
        // ---- remaining columns c0 .. N-1: inside the band already ----
        __syncthreads();
        {
            const int c = tid >> 5, d = tid & 31;
            const int j = c0 + c;
            if (j < N) {
                cplx val = cmake(0.0, 0.0);
                if (j + d < N && d <= SB) { val = ldcg(A + (size_t)(j + d) * N + j); if (d == 0) val.y = 0.0; }
                AB[(size_t)j * ABW + d] = val;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 2: bulge chasing.  One warp per sweep; lane = (r = lane & 15, hh = lane >> 4) owns row r, columns 8hh .. 8hh+7 of
// the 16 x 16 blocks.  prog[s] = hops completed by sweep s (1 << 30 when finished).  Hop h >= 1 of sweep s touches rows
// s+16h+1 .. s+16h+16 and columns s+16h-15 .. s+16h+16; hop h+1 of sweep s-1 reaches row s+16h+16 in those columns, hop h+2 of
// sweep s-1 starts at row s+16h+32: sweep s may run hop h once sweep s-1 has completed hop h + 1 (all synchronisation is inside one
// CTA: block-scope fences).  Per hop a warp loads the off-diagonal block B (registers) and the lower triangle of the diagonal
// block D (into its shared-memory slab, mirrored from there) in one go, applies the previous reflector from the right, builds the
// new one from B's first column, applies it from the left and to D from both sides.  Latency-bound by design: 4 CTAs x 4 warps per SM (a
// warp that has to wait for its predecessor sweep idles; fewer warps per matrix and more matrices per SM keep the warps busy).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef S2_MIN_CTAS
#define S2_MIN_CTAS 4      // measured, matrices/s at N = 512 / 256 (no spills up to 128 registers): 4 CTAs x 4 warps 4.9e4 / 1.8e5, 2 x 8 4.8e4 / 1.3e5, 1 x 16 3.3e4 / 7.3e4,
                           // 8 x 2 3.0e4 / 1.8e5; with 8 warps: 3 CTAs at 80 registers 4.5e4, 4 CTAs at 64 registers (spills) 3.3e4
#endif
#ifndef S2_NWARPS
#define S2_NWARPS 4
#endif
constexpr int S2_WARPS = S2_NWARPS;
constexpr int S2_DONE = 1 << 30;
#ifndef S2_SPIN_NS
#define S2_SPIN_NS 100
#endif
constexpr int S2_DS = SB + 1;                       // row stride of the D slab (conflict-free rows and columns)
constexpr int S2_SLAB = SB * S2_DS + 3 * SB;        // per warp: D block, v_a, v_b, w
SKTB_HD size_t stage2_smem_bytes(int N) { return (size_t)S2_WARPS * S2_SLAB * sizeof(cplx) + (size_t)2 * N * sizeof(int); }

// lower triangle of the diagonal block rows/cols r1 .. r1+L-1 -> slab Ds[r][c] (c <= r), zero padded to 16 x 16
template <bool FULL>
SKTB_D void s2_load_diag(const cplx* __restrict__ AB, int r1, int L, int r, int hh, cplx* Ds) {
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        if (c <= r) {
            cplx x = cmake(0.0, 0.0);
            if (FULL || (r < L && c < L)) x = ldcg(AB + (r1 + c) * ABW + (r - c));
            if (c == r) x.y = 0.0;
            Ds[r * S2_DS + c] = x;
        }
    }
}
// A <- H^H A H on the diagonal block held in the slab, reflector v (shared, zero padded), tau; stores the lower part
template <bool FULL>
SKTB_D void s2_two_sided(cplx* __restrict__ AB, int r1, int L, const cplx* vS, cplx* wS, const cplx* Ds, cplx tau, int r, int hh) {
    cplx D[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        D[cc] = c <= r ? Ds[r * S2_DS + c] : cconj(Ds[c * S2_DS + r]);
    }
    cplx y = cmake(0.0, 0.0);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) cfma(y, D[cc], vS[8 * hh + cc]);
    y.x += __shfl_xor_sync(0xffffffffu, y.x, 16); y.y += __shfl_xor_sync(0xffffffffu, y.y, 16);
    const cplx x = cmul(tau, y);
    const cplx vr = vS[r];
    cplx dot = cmulc(x, vr);                                  // conj(x_r) v_r
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o); }
    const cplx alpha = cmul(cmake(-0.5 * tau.x, -0.5 * tau.y), dot);
    cplx w = x; cfma(w, alpha, vr);
    if (hh == 0) wS[r] = w;
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        if (c <= r) {
            const cplx wc = wS[c], vc = vS[c];
            // D[r][c] -= v_r conj(w_c) + w_r conj(v_c)
            D[cc].x -= vr.x * wc.x + vr.y * wc.y + w.x * vc.x + w.y * vc.y;
            D[cc].y -= vr.y * wc.x - vr.x * wc.y + w.y * vc.x - w.x * vc.y;
            if (FULL || r < L) __stcg(AB + (r1 + c) * ABW + (r - c), D[cc]);
        }
    }
}

// zlarfg on a column held by lanes hh == 0 (x_r, r < L; zero beyond): returns v_r (own row, both halves), beta, tau
SKTB_D cplx s2_reflector(cplx x, int r, int L, double& beta, cplx& tau) {
    double n2 = (r >= 1 && r < L) ? x.x * x.x + x.y * x.y : 0.0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    n2 = __shfl_sync(0xffffffffu, n2, 0);
    const cplx alpha = cmake(__shfl_sync(0xffffffffu, x.x, 0), __shfl_sync(0xffffffffu, x.y, 0));
    cplx scale;
    fast_householder(alpha, n2, beta, tau, scale);
    cplx v = r == 0 ? cmake(1.0, 0.0) : ((r < L) ? cmul(x, scale) : cmake(0.0, 0.0));
    v.x = __shfl_sync(0xffffffffu, v.x, r); v.y = __shfl_sync(0xffffffffu, v.y, r);   // lanes hh == 1 take the value of lane r
    return v;
}

// one chase hop: B = A[r1 .. r1+L2-1][r0 .. r0+L-1], D = A[r1.., r1..]; in: reflector (vS, tau) of rows r0..; out: (v2S, tau2) of rows r1..
template <bool FULL>
SKTB_D void s2_hop(cplx* __restrict__ AB, int r0, int L, int r1, int L2, const cplx* vS, cplx* v2S, cplx* wS, cplx* Ds, cplx tau, cplx& tau2,
                   int r, int hh) {
    cplx Bk[8];
    const bool rok = FULL || r < L2;
    const int boff = r0 * ABW + L + r;                         // element (r, c) at boff + c * (ABW - 1)
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        Bk[cc] = (FULL || (rok && c < L)) ? ldcg(AB + boff + c * (ABW - 1)) : cmake(0.0, 0.0);
    }
    s2_load_diag<FULL>(AB, r1, L2, r, hh, Ds);                 // independent of B: all loads of the hop are in flight together
    // B <- B H = B - tau (B v) v^H
    cplx t = cmake(0.0, 0.0);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) cfma(t, Bk[cc], vS[8 * hh + cc]);
    t.x += __shfl_xor_sync(0xffffffffu, t.x, 16); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16);
    const cplx tt = cmul(tau, t);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const cplx vc = vS[8 * hh + cc];                       // B[r][c] -= tt conj(v_c)
        Bk[cc].x -= tt.x * vc.x + tt.y * vc.y;
        Bk[cc].y -= tt.y * vc.x - tt.x * vc.y;
    }
    // new reflector from column 0 of B
    double beta2;
    const cplx v2 = s2_reflector(hh == 0 ? Bk[0] : cmake(0.0, 0.0), r, L2, beta2, tau2);
    if (hh == 0) v2S[r] = v2;
    // B <- H2^H B = B - conj(tau2) v2 (v2^H B), four columns at a time
    const cplx cv = cmul(cconj(tau2), v2);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        cplx u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = cmulc(v2, Bk[4 * half + k]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { u[k].x += __shfl_xor_sync(0xffffffffu, u[k].x, o); u[k].y += __shfl_xor_sync(0xffffffffu, u[k].y, o); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cfms(Bk[4 * half + k], cv, u[k]);
    }
    if (hh == 0) Bk[0] = r == 0 ? cmake(beta2, 0.0) : cmake(0.0, 0.0);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        if (FULL || (rok && c < L)) __stcg(AB + boff + c * (ABW - 1), Bk[cc]);
    }
    __syncwarp();                                              // v2S and the D slab are complete
    s2_two_sided<FULL>(AB, r1, L2, v2S, wS, Ds, tau2, r, hh);
}

__global__ void __launch_bounds__(S2_WARPS * 32, S2_MIN_CTAS) stage2_kernel(Stage2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane & 15, hh = lane >> 4;
    cplx* Ds = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * S2_SLAB;
    cplx* vecs = Ds + SB * S2_DS;                              // v_a, v_b, w
    volatile int* prog = reinterpret_cast<volatile int*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)S2_WARPS * S2_SLAB);
    int* roff = const_cast<int*>(prog) + N;                    // first reflector index of every sweep (recording only)
    if (a.Rstore) {
        if (tid == 0) { int k = 0; for (int s = 0; s < N - 1; ++s) { roff[s] = k; k += (N - 1 - s + SB - 1) / SB; } }
        __syncthreads();
    }

    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        cplx* AB = a.AB + (size_t)mat * N * ABW;
        cplx* Rst = a.Rstore ? a.Rstore + (size_t)mat * a.rstride : nullptr;
        __syncthreads();
        for (int s = tid; s < N; s += S2_WARPS * 32) prog[s] = 0;
        __syncthreads();
        for (int s = warp; s < N - 1; s += S2_WARPS) {
            cplx* vS = vecs; cplx* v2S = vecs + SB; cplx* wS = vecs + 2 * SB;
            int h = 0;
            auto wait_for = [&](int need) {
                if (s > 0) { while (prog[s - 1] < need) { __nanosleep(S2_SPIN_NS); } }      // sleeping leaves the issue slots to the warps that chase
                __syncwarp();
                __threadfence_block();
            };
            auto publish = [&](int done) {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) prog[s] = done;
            };
            // ---- hop 0: annihilate column s below the subdiagonal ----
            int L = min(SB, N - 1 - s);
            wait_for(h + 2);
            cplx x = (hh == 0 && r < L) ? ldcg(AB + s * ABW + 1 + r) : cmake(0.0, 0.0);
            s2_load_diag<false>(AB, s + 1, L, r, hh, Ds);
            double beta; cplx tau;
            cplx v = s2_reflector(x, r, L, beta, tau);
            if (hh == 0 && r < L) __stcg(AB + s * ABW + 1 + r, r == 0 ? cmake(beta, 0.0) : cmake(0.0, 0.0));
            if (hh == 0) vS[r] = v;
            if (Rst && hh == 0) { cplx* rp = Rst + (size_t)roff[s] * (SB + 1); rp[r] = v; if (r == 0) rp[SB] = tau; }
            __syncwarp();
            int r0 = s + 1;
            s2_two_sided<false>(AB, r0, L, vS, wS, Ds, tau, r, hh);
            publish(++h);
            // ---- hops 1..: chase the bulge ----
            while (true) {
                const int r1 = r0 + L;
                if (r1 >= N) break;
                const int L2 = min(SB, N - r1);
                wait_for(h + 2);
                cplx tau2;
                if (L == SB && L2 == SB) s2_hop<true>(AB, r0, L, r1, L2, vS, v2S, wS, Ds, tau, tau2, r, hh);
                else s2_hop<false>(AB, r0, L, r1, L2, vS, v2S, wS, Ds, tau, tau2, r, hh);
                if (Rst && hh == 0) { cplx* rp = Rst + (size_t)(roff[s] + h) * (SB + 1); rp[r] = v2S[r]; if (r == 0) rp[SB] = tau2; }
                publish(++h);
                cplx* tmp = vS; vS = v2S; v2S = tmp;
                tau = tau2; L = L2; r0 = r1;
            }
            publish(S2_DONE);
        }
        __syncthreads();
        double* de = a.de + (size_t)mat * 2 * N;
        for (int j = tid; j < N; j += S2_WARPS * 32) {
            de[j] = __ldcg(AB + (size_t)j * ABW).x;
            de[N + j] = j < N - 1 ? __ldcg(AB + (size_t)j * ABW + 1).x : 0.0;
        }
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// Bisection on the tridiagonal matrix, one eigenvalue per thread, one CTA per matrix.  The Sturm count uses the PRODUCT form
// p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (3 FP64 operations per row; the quotient form q_i = d_i - x - e^2 / q_{i-1} of dstebz costs a
// division, ~10 FP64 operations with a Newton reciprocal - at N = 512 the bisection was as expensive as the band reduction).  The matrix is
// scaled by an exact power of two so that |d|, |e| <= 1, (p_{i-1}, p_i) are renormalised by a power of two whenever the exponent leaves
// +-300 (integer test every 8 rows), an exact zero is replaced by -2^-100 p_{i-1} (counted as a sign change, dstebz's pivmin rule).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bisect_prod_kernel(const double* __restrict__ de, int N, long long nm, double* __restrict__ evals,
                                                           long long ld, long long col0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* de2 = reinterpret_cast<double2*>(smem_raw);                  // [N] (d_r, e_{r-1}^2), scaled
    double* red = reinterpret_cast<double*>(de2 + N);                     // [3][32]
    const int tid = threadIdx.x, T = blockDim.x;
    for (long long mat = blockIdx.x; mat < nm; mat += gridDim.x) {
        const double* hd = de + (size_t)mat * 2 * N;
        __syncthreads();
        double di = 0.0, el = 0.0, er = 0.0;
        if (tid < N) {
            di = hd[tid];
            el = tid > 0 ? fabs(hd[N + tid - 1]) : 0.0;
            er = tid < N - 1 ? fabs(hd[N + tid]) : 0.0;
        }
        double emax = fmax(fabs(di), fmax(el, er));
        for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
        if ((tid & 31) == 0) red[64 + (tid >> 5)] = emax;
        __syncthreads();
        const int nw = (T + 31) >> 5;
        double tn = 0.0;
        for (int q = 0; q < nw; ++q) tn = fmax(tn, red[64 + q]);
        // exact power-of-two scale: 2^-ex with 2^(ex-1) <= tn < 2^ex  (tn = 0: no scaling)
        int ex = 0;
        if (tn > 0.0) frexp(tn, &ex);
        const double sc = scalbn(1.0, -ex), isc = scalbn(1.0, ex);
        di *= sc; el *= sc; er *= sc;
        double lo_b = 1e300, hi_b = -1e300;
        // rows interleaved: de2[r] = (d_r, e_{r-1}^2): one 16-byte broadcast load per Sturm step
        if (tid < N) { de2[tid] = make_double2(di, el * el); lo_b = di - el - er; hi_b = di + el + er; }
        for (int o = 16; o > 0; o >>= 1) {
            lo_b = fmin(lo_b, __shfl_xor_sync(0xffffffffu, lo_b, o));
            hi_b = fmax(hi_b, __shfl_xor_sync(0xffffffffu, hi_b, o));
        }
        __syncthreads();
        if ((tid & 31) == 0) { red[tid >> 5] = lo_b; red[32 + (tid >> 5)] = hi_b; }
        __syncthreads();
        double gl = 1e300, gu = -1e300;
        for (int q = 0; q < nw; ++q) { gl = fmin(gl, red[q]); gu = fmax(gu, red[32 + q]); }
        const double eps = 2.220446049250313e-16;
        const double pivmin = 1e-290;                                     // scaled units (|T| <= 1)
        gl -= 2.0 * eps * N + 2.0 * pivmin; gu += 2.0 * eps * N + 2.0 * pivmin;
        if (tid < N) {
            double lo = gl, hi = gu;
            for (int it = 0; it < 128; ++it) {
                const double x = 0.5 * (lo + hi);
                if (!(hi - lo > 2.0 * eps * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) || x <= lo || x >= hi) break;
                const int cnt = sturm_count_prod(de2, 0, N - 1, x);
                if (cnt > tid) hi = x; else lo = x;
            }
            evals[(long long)tid * ld + col0 + mat] = 0.5 * (lo + hi) * isc;
        }
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// Selected rows of U = Q1 Q2 Z after the two-stage reduction (LDOS / spectral projections need a few rows only): p = Q2^H Q1^H e_i by
// applying the block reflectors of stage 1 (V in the panels of H, T in Tstore) and then the chase reflectors in sweep / hop order,
// then U[i, n] = sum_k conj(p_k) Z[k, n].  One (matrix, row) per warp, p in shared memory.  O(N^2) per row.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ROWS_WARPS = 8;
__global__ void __launch_bounds__(ROWS_WARPS * 32) rows_kernel(RowsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* us = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * N;
    const long long total = a.nm * a.n_rows;
    for (long long work = (long long)blockIdx.x * ROWS_WARPS + warp; work < total; work += (long long)gridDim.x * ROWS_WARPS) {
        const long long mat = work / a.n_rows;
        const int r = (int)(work - mat * a.n_rows);
        const int irow = a.rows[r];
        const cplx* H = a.H + (size_t)mat * N * N;
        for (int i = lane; i < N; i += 32) us[i] = cmake(i == irow ? 1.0 : 0.0, 0.0);
        __syncwarp();
        // ---- Q1^H: u <- u - V T^H (V^H u), panel by panel ----
        for (int p = 0; p < a.npanels; ++p) {
            const int c0 = SB * p, R = c0 + SB, m = N - R;
            cplx y[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) y[c] = cmake(0.0, 0.0);
            for (int rr = lane; rr < m; rr += 32) {
                const cplx ur = us[R + rr];
                const cplx* vp = H + (size_t)(R + rr) * N + c0;
#pragma unroll
                for (int c = 0; c < SB; ++c) cfmac(y[c], vp[c], ur);
            }
#pragma unroll
            for (int c = 0; c < SB; ++c) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { y[c].x += __shfl_xor_sync(0xffffffffu, y[c].x, o); y[c].y += __shfl_xor_sync(0xffffffffu, y[c].y, o); }
            }
            // z = T^H y (lane i < 16 computes z_i), then every lane gets all of z
            const cplx* Tp = a.Tstore + ((size_t)mat * a.npanels + p) * SB * SB;
            cplx zi = cmake(0.0, 0.0);
            if (lane < SB) {
#pragma unroll
                for (int k = 0; k < SB; ++k) if (k <= lane) cfmac(zi, Tp[k * SB + lane], y[k]);
            }
            cplx z[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) { z[c].x = __shfl_sync(0xffffffffu, zi.x, c); z[c].y = __shfl_sync(0xffffffffu, zi.y, c); }
            for (int rr = lane; rr < m; rr += 32) {
                cplx ur = us[R + rr];
                const cplx* vp = H + (size_t)(R + rr) * N + c0;
#pragma unroll
                for (int c = 0; c < SB; ++c) cfms(ur, vp[c], z[c]);
                us[R + rr] = ur;
            }
            __syncwarp();
        }
        // ---- Q2^H: the chase reflectors in sweep / hop order, 16 lanes each (lanes >= 16 mirror) ----
        {
            const cplx* rp = a.Rstore + (size_t)mat * a.rstride;
            const int l = lane & 15;
            cplx vn = rp[l], tn = rp[SB];                                 // reflector 0
            for (int s = 0; s < N - 1; ++s) {
                for (int R0 = s + 1; R0 < N; R0 += SB) {
                    const cplx v = vn, tau = tn;
                    rp += SB + 1;
                    vn = rp[l]; tn = rp[SB];                              // next reflector (one slot of slack behind the last one)
                    const int L = min(SB, N - R0);
                    const cplx x = l < L ? us[R0 + l] : cmake(0.0, 0.0);
                    cplx dot = cmulc(v, x);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) { dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o); }
                    const cplx wv = cmul(cconj(tau), dot);
                    cplx xn = x; cfms(xn, wv, v);
                    if (lane < L) us[R0 + lane] = xn;
                    __syncwarp();
                }
            }
        }
        // ---- rows of U ----
        const double* Zm = a.Zf + (size_t)mat * a.zstride;
        for (int n = lane; n < N; n += 32) {
            double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
            int k = 0;
            for (; k + 1 < N; k += 2) {
                const double z0 = Zm[(size_t)k * N + n], z1 = Zm[(size_t)(k + 1) * N + n];
                const cplx p0 = us[k], p1 = us[k + 1];
                ax = fma(p0.x, z0, ax); ay = fma(-p0.y, z0, ay); bx = fma(p1.x, z1, bx); by = fma(-p1.y, z1, by);
            }
            if (k < N) { const double z0 = Zm[(size_t)k * N + n]; const cplx p0 = us[k]; ax = fma(p0.x, z0, ax); ay = fma(-p0.y, z0, ay); }
            a.vec_out[(long long)n * a.vs_band + (a.vcol0 + mat) * a.vs_k + r] = cmake(ax + bx, ay + by);
        }
        __syncwarp();
    }
}
#endif  // __CUDACC__

}  // namespace twostage
}  // namespace sktb
