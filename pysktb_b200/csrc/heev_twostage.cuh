// Two-stage tridiagonalisation for large matrices (128 <= N <= ~600, eigenvalues only).
//
// The one-stage Householder reduction (kernels_generic.cuh) reads the trailing matrix once per reflector: 16 N^3 / 3 bytes
// per matrix (0.9 GB at N = 512 against 4 MB of matrix) - it is HBM-bound at 11 % of the FP64 peak.  Here:
//   stage 1  dense -> band of width SB = 16 (LAPACK zhetrd_he2hb idea): QR of a 16-column panel below the band (panel in
//            shared memory, one warp per column), compact-WY block reflector Q = I - V T V^H, two-sided update of the trailing
//            matrix A <- Q^H A Q = A - V W^H - W V^H with X = A (V T), W = X - 1/2 V (T^H V^H X).  Both matrix-sized
//            operations (X = A V~ and the rank-32 update) are complex GEMMs on the FP64 tensor instruction
//            (mma.sync.m8n8k4.f64 -> DMMA.8x8x4): measured on B200 it shares the DFMA pipe but reaches the full 37 TFLOP/s
//            with 8x fewer issue slots and no register-operand limit (tools/probe/dmma_probe.cu).  The trailing matrix is
//            read twice and written once per PANEL (32 m^2 bytes, lower triangle only): N^3/(3*16)*32 B = 90 MB at N = 512.
//   stage 2  band -> tridiagonal by bulge chasing (zhbtrd / sb2st idea): sweep s annihilates column s of the band and chases
//            the 16x16 bulge down the band; one warp per sweep, sweeps pipelined three hops apart inside a CTA (progress
//            counters in shared memory), the band (2*SB diagonals with the fill, 262 KB at N = 512) stays in L2.
// Every reflector follows zlarfg (H^H x = beta e1, H = I - tau v v^H, beta real), the band layout is AB[j][d] = A[j+d][j].
// tools/proto_twostage.py is the numpy model of both stages (same blocking, conventions and index arithmetic).
// Replaces, for eigenvalues of large matrices, the LAPACK call behind numpy.linalg.eigh at pysktb/hamiltonian.py:119.
#pragma once
#include "common.cuh"

namespace sktb {
namespace twostage {

constexpr int SB = 16;            // band width after stage 1
constexpr int ABW = 2 * SB;       // stored diagonals per column (d = 0 .. 2 SB - 1: the bulge chase fills up to 2 SB - 1)
constexpr int S1_WARPS = 16;
constexpr int S1_THREADS = S1_WARPS * 32;

struct Stage1Args {
    cplx* H; int N; long long nm;   // [nm][N][N] row-major, lower triangle read (zheevd UPLO='L'), destroyed
    cplx* AB;                       // [nm][N][ABW]
    cplx* Vg; cplx* Vt; cplx* Wg;   // [nm][NR][SB] each (NR = N rounded up to 8): V, V T and X / W of the current panel
    int NR, mp;                     // mp: column stride of the panel in shared memory (odd)
};
struct Stage2Args {
    cplx* AB; int N; long long nm;
    double* de;                     // [nm][2][N]: d, then e (N - 1 entries)
};

SKTB_HD size_t stage1_smem_bytes(int N) {
    const int mp = (((N - SB + 7) / 8) * 8) | 1;
    return ((size_t)SB * mp + 3 * SB * SB + (size_t)S1_WARPS * SB * SB + 4 * SB) * sizeof(cplx) + 64;
}
SKTB_HD size_t scratch_bytes_per_matrix(int N) {
    const size_t NR = (size_t)(N + 7) / 8 * 8;
    return ((size_t)N * ABW + 3 * NR * SB) * sizeof(cplx) + (size_t)2 * N * sizeof(double);
}

// host: stage 1 + stage 2 on `nm` matrices; de[mat][2][N] receives the tridiagonal.  scratch: nm * scratch_bytes_per_matrix(N).
// Returns the number of kernel launches, or -1 when N does not fit (shared memory) - the caller falls back to the one-stage path.
// stages: bit 0 = stage 1 (H -> band in scratch), bit 1 = stage 2 (band -> de).  band_of(scratch) is the [nm][N][ABW] band array.
int launch_tridiag(cplx* H, int N, long long nm, void* scratch, double* de, cudaStream_t st, char* info, size_t info_len, int stages = 3);
inline cplx* band_of(void* scratch) { return reinterpret_cast<cplx*>(scratch); }
bool supported(int N);

#if defined(__CUDACC__) && defined(SKTB_TWOSTAGE_KERNELS)      // the kernels live in heev_twostage.cu only
// D(8x8) += A(8x4) B(4x8) in FP64; lane = 4 g + q: a = A[g][q], b = B[q][g], c0/c1 = C[g][2q], C[g][2q+1]
SKTB_D void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
SKTB_D double dneg(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
SKTB_D cplx ldcg(const cplx* p) { return __ldcg(p); }
SKTB_D cplx cwarp_sum(cplx v) { return cmake(warp_sum(v.x), warp_sum(v.y)); }

// ---------------------------------------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S1_THREADS, 1) stage1_kernel(Stage1Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, mp = a.mp, NR = a.NR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    cplx* Ps = reinterpret_cast<cplx*>(smem_raw);            // [SB][mp] panel, column-major; V after the QR
    cplx* Ts = Ps + (size_t)SB * mp;                         // [SB][SB] T (upper triangular), Ts[k*SB+n]
    cplx* Rs = Ts + SB * SB;                                 // [SB][SB] R, Rs[r*SB+c]
    cplx* Ss = Rs + SB * SB;                                 // [SB][SB] G, then S = T^H G
    cplx* Gp = Ss + SB * SB;                                 // [S1_WARPS][SB*SB] per-warp partials of G
    cplx* dots = Gp + (size_t)S1_WARPS * SB * SB;            // [SB]
    cplx* zs = dots + SB;                                    // [SB]
    cplx* taus = zs + SB;                                    // [SB]
    cplx* scs = taus + SB;                                   // [SB] scale of column c (the betas go straight into the diagonal of Rs)

    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        cplx* A = a.H + (size_t)mat * N * N;
        cplx* AB = a.AB + (size_t)mat * N * ABW;
        cplx* Vg = a.Vg + (size_t)mat * NR * SB;
        cplx* Vt = a.Vt + (size_t)mat * NR * SB;
        cplx* Wg = a.Wg + (size_t)mat * NR * SB;
        int c0 = 0;
        for (;; c0 += SB) {
            const int m = N - c0 - SB;                       // rows below the band of this panel
            if (m <= 0) break;
            const int mr = (m + 7) / 8 * 8;
            const int nt = mr / 8;
            cplx* A22 = A + (size_t)(c0 + SB) * N + (c0 + SB);
            // ---- 1. panel -> shared memory (column-major) ----
            __syncthreads();
            for (int idx = tid; idx < m * SB; idx += S1_THREADS) {
                const int r = idx >> 4, c = idx & 15;
                Ps[(size_t)c * mp + r] = ldcg(A + (size_t)(c0 + SB + r) * N + c0 + c);
            }
            for (int idx = tid; idx < 3 * SB * SB; idx += S1_THREADS) Ts[idx] = cmake(0.0, 0.0);   // T, R, S
            __syncthreads();
            // ---- 2. Householder QR of the panel, one warp per column ----
            cplx* myc = Ps + (size_t)warp * mp;
            const int ncol = m < SB ? m : SB;                 // columns that have a diagonal row
            for (int c = 0; c < ncol; ++c) {
                const cplx* pc = Ps + (size_t)c * mp;
                if (c > 0 && warp == c - 1) {                 // finalise column c-1: V entries below the diagonal
                    const cplx sc = scs[c - 1];
                    for (int r = c + lane; r < m; r += 32) myc[r] = cmul(myc[r], sc);
                    if (lane < c - 1) { Rs[lane * SB + (c - 1)] = myc[lane]; }
                    __syncwarp();
                    if (lane < c - 1) myc[lane] = cmake(0.0, 0.0);
                    if (lane == 0) myc[c - 1] = cmake(1.0, 0.0);
                    __syncwarp();
                }
                // raw_w = sum_{r>c} conj(P[r][c]) P[r][w]
                cplx acc = cmake(0.0, 0.0);
                for (int r = c + 1 + lane; r < m; r += 32) cfmac(acc, pc[r], myc[r]);
                acc = cwarp_sum(acc);
                if (lane == 0) dots[warp] = acc;
                __syncthreads();
                const cplx alpha = pc[c];
                double beta; cplx tau, scale;
                householder(alpha, dots[c].x, beta, tau, scale);
                const cplx ctau = cconj(tau);
                const cplx mydot = dots[warp];              // (dots is rewritten only after the barrier that ends this step)
                if (warp > c) {
                    const cplx s = cadd(myc[c], cmulc(scale, mydot));            // v^H a_w = a_w[c] + conj(scale) raw_w
                    const cplx coef = cmul(ctau, s);
                    const cplx cs2 = cmul(coef, scale);
                    for (int r = c + 1 + lane; r < m; r += 32) cfms(myc[r], cs2, pc[r]);
                    __syncwarp();
                    if (lane == 0) myc[c] = csub(myc[c], coef);
                } else if (warp < c) {
                    if (lane == 0) zs[warp] = cadd(cconj(myc[c]), cmul(scale, cconj(mydot)));   // V_w^H v_c
                } else {
                    if (lane == 0) { taus[c] = tau; scs[c] = scale; Rs[c * SB + c] = cmake(beta, 0.0); }
                }
                __syncthreads();
                // T[:c][c] = -tau T[:c][:c] z, T[c][c] = tau
                if (tid < c) {
                    cplx t = cmake(0.0, 0.0);
                    for (int k = tid; k < c; ++k) cfma(t, Ts[tid * SB + k], zs[k]);
                    Ts[tid * SB + c] = cmul(cmake(-tau.x, -tau.y), t);
                } else if (tid == c) Ts[c * SB + c] = tau;
            }
            __syncthreads();
            // finalise the columns the loop has not: w >= ncol - 1
            if (warp >= ncol - 1) {
                if (warp < ncol) {                            // last column with a diagonal row
                    const cplx sc = scs[warp];
                    for (int r = warp + 1 + lane; r < m; r += 32) myc[r] = cmul(myc[r], sc);
                    if (lane < warp) Rs[lane * SB + warp] = myc[lane];
                    __syncwarp();
                    if (lane < warp) myc[lane] = cmake(0.0, 0.0);
                    if (lane == 0) myc[warp] = cmake(1.0, 0.0);
                } else {                                      // no diagonal row (m < SB): the column is R, V column is zero
                    if (lane < m) Rs[lane * SB + warp] = myc[lane];
                    __syncwarp();
                    if (lane < m) myc[lane] = cmake(0.0, 0.0);
                }
            }
            __syncthreads();
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
            // ---- 4. V and V~ = V T to global memory, row-major [row][16] ----
            for (int r = tid; r < mr; r += S1_THREADS) {
                cplx v[SB];
#pragma unroll
                for (int k = 0; k < SB; ++k) v[k] = r < m ? Ps[(size_t)k * mp + r] : cmake(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < SB; ++k) Vg[(size_t)r * SB + k] = v[k];
#pragma unroll
                for (int n = 0; n < SB; ++n) {
                    cplx t = cmake(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k <= n; ++k) cfma(t, v[k], Ts[k * SB + n]);
                    Vt[(size_t)r * SB + n] = t;
                }
            }
            __syncthreads();
            // ---- 5. X = A22 V~ (lower triangle of A22 read: tile (I,J) for J <= I, tile (J,I)^H for J > I) ----
            const int np = (nt + 1) / 2;
            for (int pi = warp; pi < np; pi += S1_WARPS) {
                for (int half = 0; half < 2; ++half) {
                    const int I = half ? nt - 1 - pi : pi;
                    if (half && I == pi) break;
                    double Xr[2][2] = {{0, 0}, {0, 0}}, Xi[2][2] = {{0, 0}, {0, 0}};
                    const int row = 8 * I + g;
                    const bool rok = row < m;
                    // row orientation, J < I
#pragma unroll 2
                    for (int J = 0; J < I; ++J) {
                        const cplx* ap = A22 + (size_t)row * N + 8 * J + 2 * q;
                        cplx a0 = rok ? ldcg(ap) : cmake(0.0, 0.0), a1 = rok ? ldcg(ap + 1) : cmake(0.0, 0.0);
                        const cplx* bp = Vt + (size_t)(8 * J + 2 * q) * SB + g;
                        const cplx b00 = bp[0], b01 = bp[8], b10 = bp[SB], b11 = bp[SB + 8];
                        dmma(Xr[0][0], Xr[0][1], a0.x, b00.x); dmma(Xr[0][0], Xr[0][1], dneg(a0.y), b00.y);
                        dmma(Xi[0][0], Xi[0][1], a0.x, b00.y); dmma(Xi[0][0], Xi[0][1], a0.y, b00.x);
                        dmma(Xr[1][0], Xr[1][1], a0.x, b01.x); dmma(Xr[1][0], Xr[1][1], dneg(a0.y), b01.y);
                        dmma(Xi[1][0], Xi[1][1], a0.x, b01.y); dmma(Xi[1][0], Xi[1][1], a0.y, b01.x);
                        dmma(Xr[0][0], Xr[0][1], a1.x, b10.x); dmma(Xr[0][0], Xr[0][1], dneg(a1.y), b10.y);
                        dmma(Xi[0][0], Xi[0][1], a1.x, b10.y); dmma(Xi[0][0], Xi[0][1], a1.y, b10.x);
                        dmma(Xr[1][0], Xr[1][1], a1.x, b11.x); dmma(Xr[1][0], Xr[1][1], dneg(a1.y), b11.y);
                        dmma(Xi[1][0], Xi[1][1], a1.x, b11.y); dmma(Xi[1][0], Xi[1][1], a1.y, b11.x);
                    }
                    // diagonal tile
                    {
                        cplx av[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 8 * I + 2 * q + e;
                            cplx x = cmake(0.0, 0.0);
                            if (row < m && col < m) {
                                if (row > col) x = ldcg(A22 + (size_t)row * N + col);
                                else if (row < col) x = cconj(ldcg(A22 + (size_t)col * N + row));
                                else x = cmake(ldcg(A22 + (size_t)row * N + row).x, 0.0);
                            }
                            av[e] = x;
                        }
                        const cplx* bp = Vt + (size_t)(8 * I + 2 * q) * SB + g;
                        const cplx b00 = bp[0], b01 = bp[8], b10 = bp[SB], b11 = bp[SB + 8];
                        const cplx a0 = av[0], a1 = av[1];
                        dmma(Xr[0][0], Xr[0][1], a0.x, b00.x); dmma(Xr[0][0], Xr[0][1], dneg(a0.y), b00.y);
                        dmma(Xi[0][0], Xi[0][1], a0.x, b00.y); dmma(Xi[0][0], Xi[0][1], a0.y, b00.x);
                        dmma(Xr[1][0], Xr[1][1], a0.x, b01.x); dmma(Xr[1][0], Xr[1][1], dneg(a0.y), b01.y);
                        dmma(Xi[1][0], Xi[1][1], a0.x, b01.y); dmma(Xi[1][0], Xi[1][1], a0.y, b01.x);
                        dmma(Xr[0][0], Xr[0][1], a1.x, b10.x); dmma(Xr[0][0], Xr[0][1], dneg(a1.y), b10.y);
                        dmma(Xi[0][0], Xi[0][1], a1.x, b10.y); dmma(Xi[0][0], Xi[0][1], a1.y, b10.x);
                        dmma(Xr[1][0], Xr[1][1], a1.x, b11.x); dmma(Xr[1][0], Xr[1][1], dneg(a1.y), b11.y);
                        dmma(Xi[1][0], Xi[1][1], a1.x, b11.y); dmma(Xi[1][0], Xi[1][1], a1.y, b11.x);
                    }
                    // column orientation, J > I: A-operand [g][k = 2q+e] = conj(A22[8J+2q+e][8I+g])
                    const int colI = 8 * I + g;
                    const bool cok = colI < m;
#pragma unroll 2
                    for (int J = I + 1; J < nt; ++J) {
                        const int r0 = 8 * J + 2 * q;
                        const cplx* ap = A22 + (size_t)r0 * N + colI;
                        cplx a0 = (cok && r0 < m) ? ldcg(ap) : cmake(0.0, 0.0), a1 = (cok && r0 + 1 < m) ? ldcg(ap + N) : cmake(0.0, 0.0);
                        const cplx* bp = Vt + (size_t)r0 * SB + g;
                        const cplx b00 = bp[0], b01 = bp[8], b10 = bp[SB], b11 = bp[SB + 8];
                        // conj(a): re = a.x, im = -a.y
                        dmma(Xr[0][0], Xr[0][1], a0.x, b00.x); dmma(Xr[0][0], Xr[0][1], a0.y, b00.y);
                        dmma(Xi[0][0], Xi[0][1], a0.x, b00.y); dmma(Xi[0][0], Xi[0][1], dneg(a0.y), b00.x);
                        dmma(Xr[1][0], Xr[1][1], a0.x, b01.x); dmma(Xr[1][0], Xr[1][1], a0.y, b01.y);
                        dmma(Xi[1][0], Xi[1][1], a0.x, b01.y); dmma(Xi[1][0], Xi[1][1], dneg(a0.y), b01.x);
                        dmma(Xr[0][0], Xr[0][1], a1.x, b10.x); dmma(Xr[0][0], Xr[0][1], a1.y, b10.y);
                        dmma(Xi[0][0], Xi[0][1], a1.x, b10.y); dmma(Xi[0][0], Xi[0][1], dneg(a1.y), b10.x);
                        dmma(Xr[1][0], Xr[1][1], a1.x, b11.x); dmma(Xr[1][0], Xr[1][1], a1.y, b11.y);
                        dmma(Xi[1][0], Xi[1][1], a1.x, b11.y); dmma(Xi[1][0], Xi[1][1], dneg(a1.y), b11.x);
                    }
                    // X tile -> Wg (C layout: row g, columns 8t + 2q, 8t + 2q + 1)
                    cplx* xp = Wg + (size_t)row * SB + 2 * q;
                    xp[0] = cmake(Xr[0][0], Xi[0][0]); xp[1] = cmake(Xr[0][1], Xi[0][1]);
                    xp[8] = cmake(Xr[1][0], Xi[1][0]); xp[9] = cmake(Xr[1][1], Xi[1][1]);
                }
            }
            __syncthreads();
            // ---- 6. G = V^H X (16x16), per-warp partial sums over the warp's rows, then S = T^H G ----
            {
                double Gr[2][2][2], Gi[2][2][2];
#pragma unroll
                for (int t1 = 0; t1 < 2; ++t1)
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) { Gr[t1][t2][0] = Gr[t1][t2][1] = Gi[t1][t2][0] = Gi[t1][t2][1] = 0.0; }
                for (int ks = warp; ks < mr / 4; ks += S1_WARPS) {
                    const int rk = 4 * ks + q;
                    const cplx va = Vg[(size_t)rk * SB + g], vb = Vg[(size_t)rk * SB + g + 8];
                    const cplx xa = Wg[(size_t)rk * SB + g], xb = Wg[(size_t)rk * SB + g + 8];
                    const cplx vv[2] = {va, vb}, xx[2] = {xa, xb};
#pragma unroll
                    for (int t1 = 0; t1 < 2; ++t1)
#pragma unroll
                        for (int t2 = 0; t2 < 2; ++t2) {
                            // conj(v) x: re = vr xr + vi xi, im = vr xi - vi xr
                            dmma(Gr[t1][t2][0], Gr[t1][t2][1], vv[t1].x, xx[t2].x); dmma(Gr[t1][t2][0], Gr[t1][t2][1], vv[t1].y, xx[t2].y);
                            dmma(Gi[t1][t2][0], Gi[t1][t2][1], vv[t1].x, xx[t2].y); dmma(Gi[t1][t2][0], Gi[t1][t2][1], dneg(vv[t1].y), xx[t2].x);
                        }
                }
                cplx* gp = Gp + (size_t)warp * SB * SB;
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
                for (int w = 0; w < S1_WARPS; ++w) s = cadd(s, Gp[(size_t)w * SB * SB + tid]);
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
            // ---- 7. W = X - 1/2 V S, one row per thread ----
            for (int r = tid; r < mr; r += S1_THREADS) {
                cplx x[SB];
#pragma unroll
                for (int n = 0; n < SB; ++n) x[n] = Wg[(size_t)r * SB + n];
                if (r < m) {
                    for (int k = 0; k < SB; ++k) {
                        const cplx vk = Ps[(size_t)k * mp + r];
#pragma unroll
                        for (int n = 0; n < SB; ++n) cfms(x[n], vk, Ss[k * SB + n]);
                    }
                }
#pragma unroll
                for (int n = 0; n < SB; ++n) Wg[(size_t)r * SB + n] = x[n];
            }
            __syncthreads();
            // ---- 8. A22 -= V W^H + W V^H on the lower tiles ----
            for (int pi = warp; pi < np; pi += S1_WARPS) {
                for (int half = 0; half < 2; ++half) {
                    const int I = half ? nt - 1 - pi : pi;
                    if (half && I == pi) break;
                    const int row = 8 * I + g;
                    const bool rok = row < m;
                    cplx vI[4], wI[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) { vI[s] = Vg[(size_t)row * SB + 4 * s + q]; wI[s] = Wg[(size_t)row * SB + 4 * s + q]; }
#pragma unroll 2
                    for (int J = 0; J <= I; ++J) {
                        const int col = 8 * J + 2 * q;
                        cplx* cp = A22 + (size_t)row * N + col;
                        const bool ok0 = rok && col < m, ok1 = rok && col + 1 < m;
                        cplx c0v = ok0 ? ldcg(cp) : cmake(0.0, 0.0), c1v = ok1 ? ldcg(cp + 1) : cmake(0.0, 0.0);
                        double Pr[2] = {0, 0}, Pp[2] = {0, 0}, Pm[2] = {0, 0};
                        const cplx* vj = Vg + (size_t)(8 * J + g) * SB + q;
                        const cplx* wj = Wg + (size_t)(8 * J + g) * SB + q;
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const cplx v2 = vj[4 * s], w2 = wj[4 * s];
                            dmma(Pr[0], Pr[1], vI[s].x, w2.x); dmma(Pr[0], Pr[1], vI[s].y, w2.y);
                            dmma(Pr[0], Pr[1], wI[s].x, v2.x); dmma(Pr[0], Pr[1], wI[s].y, v2.y);
                            dmma(Pp[0], Pp[1], vI[s].y, w2.x); dmma(Pp[0], Pp[1], wI[s].y, v2.x);
                            dmma(Pm[0], Pm[1], vI[s].x, w2.y); dmma(Pm[0], Pm[1], wI[s].x, v2.y);
                        }
                        c0v.x -= Pr[0]; c0v.y -= Pp[0] - Pm[0];
                        c1v.x -= Pr[1]; c1v.y -= Pp[1] - Pm[1];
                        if (ok0) cp[0] = c0v;
                        if (ok1) cp[1] = c1v;
                    }
                }
            }
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
// the 16 x 16 blocks.  prog[s] = hops completed by sweep s (1 << 30 when finished); sweep s may run hop h once sweep s-1
// has completed hop h + 2.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int S2_WARPS = 8;
constexpr int S2_DONE = 1 << 30;

struct HopVec { cplx* v; cplx* w; };

// A <- H^H A H on the diagonal block rows/cols r1 .. r1+L-1 (lower storage), reflector v (shared, length 16, zero padded), tau
SKTB_D void s2_two_sided(cplx* AB, int r1, int L, const cplx* vS, cplx* wS, cplx tau, int r, int hh) {
    cplx D[8];
    const bool rok = r < L;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        cplx x = cmake(0.0, 0.0);
        if (rok && c < L) {
            if (r >= c) { x = ldcg(AB + (size_t)(r1 + c) * ABW + (r - c)); if (r == c) x.y = 0.0; }
            else x = cconj(ldcg(AB + (size_t)(r1 + r) * ABW + (c - r)));
        }
        D[cc] = x;
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
    __syncwarp();
    if (hh == 0) wS[r] = w;
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const int c = 8 * hh + cc;
        const cplx wc = wS[c], vc = vS[c];
        // D[r][c] -= v_r conj(w_c) + w_r conj(v_c)
        D[cc].x -= vr.x * wc.x + vr.y * wc.y + w.x * vc.x + w.y * vc.y;
        D[cc].y -= vr.y * wc.x - vr.x * wc.y + w.y * vc.x - w.x * vc.y;
        if (rok && c <= r) __stcg(AB + (size_t)(r1 + c) * ABW + (r - c), D[cc]);
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
    householder(alpha, n2, beta, tau, scale);
    cplx v = r == 0 ? cmake(1.0, 0.0) : ((r < L) ? cmul(x, scale) : cmake(0.0, 0.0));
    v.x = __shfl_sync(0xffffffffu, v.x, r); v.y = __shfl_sync(0xffffffffu, v.y, r);   // lanes hh == 1 take the value of lane r
    return v;
}

__global__ void __launch_bounds__(S2_WARPS * 32, 2) stage2_kernel(Stage2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane & 15, hh = lane >> 4;
    cplx* slab = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 3 * SB;     // v_a, v_b, w
    volatile int* prog = reinterpret_cast<volatile int*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)S2_WARPS * 3 * SB);

    for (long long mat = blockIdx.x; mat < a.nm; mat += gridDim.x) {
        cplx* AB = a.AB + (size_t)mat * N * ABW;
        __syncthreads();
        for (int s = tid; s < N; s += S2_WARPS * 32) prog[s] = 0;
        __syncthreads();
        for (int s = warp; s < N - 1; s += S2_WARPS) {
            cplx* vS = slab; cplx* v2S = slab + SB; cplx* wS = slab + 2 * SB;
            int h = 0;
            auto wait_for = [&](int need) {
                if (s > 0) { while (prog[s - 1] < need) { } }
                __syncwarp();
                __threadfence();
            };
            auto publish = [&](int done) {
                __threadfence();
                __syncwarp();
                if (lane == 0) prog[s] = done;
            };
            // ---- hop 0: annihilate column s below the subdiagonal ----
            int L = min(SB, N - 1 - s);
            wait_for(h + 3);
            cplx x = (hh == 0 && r < L) ? ldcg(AB + (size_t)s * ABW + 1 + r) : cmake(0.0, 0.0);
            double beta; cplx tau;
            cplx v = s2_reflector(x, r, L, beta, tau);
            if (hh == 0 && r < L) __stcg(AB + (size_t)s * ABW + 1 + r, r == 0 ? cmake(beta, 0.0) : cmake(0.0, 0.0));
            __syncwarp();
            if (hh == 0) vS[r] = v;
            __syncwarp();
            int r0 = s + 1;
            s2_two_sided(AB, r0, L, vS, wS, tau, r, hh);
            publish(++h);
            // ---- hops 1..: chase the bulge ----
            while (true) {
                const int r1 = r0 + L;
                if (r1 >= N) break;
                const int L2 = min(SB, N - r1);
                wait_for(h + 3);
                cplx Bk[8];
                const bool rok = r < L2;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = 8 * hh + cc;
                    Bk[cc] = (rok && c < L) ? ldcg(AB + (size_t)(r0 + c) * ABW + (L - c + r)) : cmake(0.0, 0.0);
                }
                // B <- B H = B - tau (B v) v^H
                cplx t = cmake(0.0, 0.0);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) cfma(t, Bk[cc], vS[8 * hh + cc]);
                t.x += __shfl_xor_sync(0xffffffffu, t.x, 16); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16);
                const cplx tt = cmul(tau, t);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const cplx vc = vS[8 * hh + cc];                 // B[r][c] -= tt conj(v_c)
                    Bk[cc].x -= tt.x * vc.x + tt.y * vc.y;
                    Bk[cc].y -= tt.y * vc.x - tt.x * vc.y;
                }
                // new reflector from column 0 of B
                double beta2; cplx tau2;
                const cplx v2 = s2_reflector(hh == 0 ? Bk[0] : cmake(0.0, 0.0), r, L2, beta2, tau2);
                __syncwarp();
                if (hh == 0) v2S[r] = v2;
                // B <- H2^H B = B - conj(tau2) v2 (v2^H B)
                cplx u[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    u[cc] = cmulc(v2, Bk[cc]);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) { u[cc].x += __shfl_xor_sync(0xffffffffu, u[cc].x, o); u[cc].y += __shfl_xor_sync(0xffffffffu, u[cc].y, o); }
                }
                const cplx cv = cmul(cconj(tau2), v2);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = 8 * hh + cc;
                    cfms(Bk[cc], cv, u[cc]);
                    if (c == 0) Bk[cc] = r == 0 ? cmake(beta2, 0.0) : cmake(0.0, 0.0);
                    if (rok && c < L) __stcg(AB + (size_t)(r0 + c) * ABW + (L - c + r), Bk[cc]);
                }
                __syncwarp();
                s2_two_sided(AB, r1, L2, v2S, wS, tau2, r, hh);
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
#endif  // __CUDACC__

}  // namespace twostage
}  // namespace sktb
