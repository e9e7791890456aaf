// K5: the reference's dense-inverse Green's function, G(E,k) = inv((E + i eta) I - H(k))  (greens.py:70, :455), for models whose
// H(k) is NOT Hermitian (f-orbital blocks, d-shell SOC).  For Hermitian H(k) the diagonal of G is a Lorentzian
// sum over eigenpairs (K2 + K3, one eigensolve per k); for a non-Hermitian H(k) it is not - the reference inverts the FULL matrix
// get_ham returns, upper triangle included - so this kernel does what the reference does: one complex inverse per (k, E).
//
// One CTA per (k, E) item (persistent grid, items dealt round-robin), the N x N matrix in shared memory (row stride N + 1: column
// reads touch all banks), in-place Gauss-Jordan with IMPLICIT row pivoting - step c takes the unused row p_c with the largest
// |Re| + |Im| entry of column c (LAPACK izamax's measure) and no rows are swapped:
//     stage:   prow[j] = A[p][j] / piv (prow[c] = 1 / piv),  fcol[r] = A[r][c] (fcol[p] = 0)                      | barrier
//     update:  A[r][j] = (j == c ? 0 : A[r][j]) - fcol[r] prow[j]   (r != p),   A[p][j] = prow[j]                  | barrier
// Every warp repeats the pivot search (N loads) instead of a third barrier.  With P the permutation that puts row p_c at position
// c the sweep leaves (P A)^-1 = A^-1 P^T in P A's row numbering:  A^-1[i][m] = R[p_i][q_m], q the inverse permutation; only the
// diagonal entries the caller's index groups name are read (tools/proto_gj.py is the numpy model of this file).
// Cost: 8 N^3 flops per item, bound by shared-memory bandwidth (one 16 B load + store per 4 DFMA).  Matrices that do not fit
// shared memory (N > 118) run the same code on a per-CTA global workspace (L2-resident; slow, correct).
#pragma once
#include "common.cuh"

namespace sktb {

struct GfArgs {
    const cplx* H;          // [nk][N][N] row-major (K1 output: get_ham's full matrix)
    int N; long long nk;
    const double* E; int nE; double eta;
    int n_groups;           // 0: trace
    const int* gptr;        // [n_groups + 1] device
    const int* grows;       // row indices (with multiplicity, as the reference's loops visit them)
    double* part;           // [nk][max(n_groups,1)][nE]:  -(1/pi) sum_{i in group} Im G_ii
    cplx* ws;               // NULL: matrix in shared memory, else [gridDim.x][N (N + 1)]
    cplx* Gout;             // non-NULL: the full inverse of every item, [nk][nE][N][N] row-major
};

constexpr int GF_THREADS = 256;

__global__ void __launch_bounds__(GF_THREADS) gf_inverse_kernel(GfArgs a) {
    extern __shared__ __align__(16) unsigned char gf_smem[];
    const int N = a.N, ld = N + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = GF_THREADS / 32;
    cplx* base = reinterpret_cast<cplx*>(gf_smem);
    cplx* A = a.ws ? a.ws + (size_t)blockIdx.x * N * ld : base;
    cplx* prow = a.ws ? base : base + (size_t)N * ld;
    cplx* fcol = prow + N;
    int* p_of_step = reinterpret_cast<int*>(fcol + N);
    int* step_of_row = p_of_step + N;
    const int ng = a.n_groups > 0 ? a.n_groups : 1;
    const long long n_items = a.nk * a.nE;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long k = item / a.nE;
        const int e = (int)(item - k * a.nE);
        const cplx z = cmake(a.E[e], a.eta);
        const cplx* Hk = a.H + (size_t)k * N * N;
        __syncthreads();                                   // previous item's read-out is complete
        for (int r = warp; r < N; r += n_warps)
            for (int j = lane; j < N; j += 32) {
                const cplx h = Hk[(size_t)r * N + j];
                A[r * ld + j] = r == j ? cmake(z.x - h.x, z.y - h.y) : cmake(-h.x, -h.y);
            }
        __syncthreads();
        unsigned long long used = 0ull;                    // bit t: row lane + 32 t has been a pivot row (identical in every warp)
        for (int c = 0; c < N; ++c) {
            // ---- pivot search, repeated by every warp ----
            double best = -1.0; int bp = 0x7fffffff;
            for (int r = lane, t = 0; r < N; r += 32, ++t) {
                if ((used >> t) & 1ull) continue;
                const cplx v = A[r * ld + c];
                const double m = fabs(v.x) + fabs(v.y);
                if (m > best) { best = m; bp = r; }         // rows ascend: the first maximum wins
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
            }
            const int p = bp;
            if ((p & 31) == lane) used |= 1ull << (p >> 5);
            // ---- stage the scaled pivot row and the factor column ----
            const cplx piv = A[p * ld + c];
            const double den = piv.x * piv.x + piv.y * piv.y;
            const cplx inv = cmake(piv.x / den, -piv.y / den);
            for (int j = tid; j < N; j += GF_THREADS) {
                prow[j] = j == c ? inv : cmul(A[p * ld + j], inv);
                fcol[j] = j == p ? cmake(0.0, 0.0) : A[j * ld + c];
            }
            if (tid == 0) { p_of_step[c] = p; step_of_row[p] = c; }
            __syncthreads();
            // ---- rank-1 update of all rows but the pivot row, which takes the scaled row ----
            for (int j = lane; j < N; j += 32) {
                const cplx pj = prow[j];
                for (int r = warp; r < N; r += n_warps) {
                    cplx v = A[r * ld + j];
                    if (j == c) v = cmake(0.0, 0.0);
                    cfms(v, fcol[r], pj);
                    A[r * ld + j] = r == p ? pj : v;
                }
            }
            __syncthreads();
        }
        // ---- read-out: -(1/pi) sum_{i in group} Im G_ii,  G_ii = R[p_i][q_i] ----
        for (int g = warp; g < ng; g += n_warps) {
            double s = 0.0;
            if (a.n_groups > 0) {
                for (int t = a.gptr[g] + lane; t < a.gptr[g + 1]; t += 32) {
                    const int i = a.grows[t];
                    s += A[p_of_step[i] * ld + step_of_row[i]].y;
                }
            } else {
                for (int i = lane; i < N; i += 32) s += A[p_of_step[i] * ld + step_of_row[i]].y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && a.part) a.part[((size_t)k * ng + g) * a.nE + e] = -s / 3.141592653589793;
        }
        if (a.Gout) {
            cplx* G = a.Gout + (size_t)item * N * N;
            for (int i = warp; i < N; i += n_warps)
                for (int m = lane; m < N; m += 32) G[(size_t)i * N + m] = A[p_of_step[i] * ld + step_of_row[m]];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// N <= 32 (the north-star model T among them): ONE MATRIX PER WARP, lane r owns row r in registers (32 complex = 128 registers;
// smaller matrices are padded with an identity block, which the elimination never mixes with the rest).  The same Gauss-Jordan sweep
// with implicit row pivoting, arranged so that one code body serves every step without dynamic register indices:
//   * the register window ROTATES: step c reads its column from register 0, every update writes its result one register down
//     (a DFMA's destination is free), and the finished column of the inverse enters at register 31 - after 32 steps register j
//     holds column j of R;
//   * the pivot row is not scaled before it is broadcast: lane p stores its raw row (predicated STS.128), every lane folds 1/piv
//     into its own factor, g_r = f_r / piv, and the pivot lane uses g_p = 1 - 1/piv, so that the one update  B -= g * prow  also
//     turns the pivot row into prow / piv; the new column is  r == p ? 1/piv : -g_r;
//   * the pivot search is one REDUX: key = (float bits of |Re| + |Im|, low 5 bits replaced by 31 - lane), 0 for used rows
//     (an FP32 measure decides between near-equal candidates differently from LAPACK's FP64 izamax; either choice is stable).
// Per step: 124 DFMA against 31 broadcast LDS.128 + 31 single-lane STS.128; no barrier, no spill at 3 x 4 warps per SM.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int GFW_WARPS = 4;

__global__ void __launch_bounds__(GFW_WARPS * 32, 3) gf_inverse_warp_kernel(GfArgs a) {
    __shared__ __align__(16) cplx s_prow[GFW_WARPS][32];
    __shared__ double s_diag[GFW_WARPS][32];
    const int N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ng = a.n_groups > 0 ? a.n_groups : 1;
    const long long n_items = a.nk * a.nE;
    cplx* prow = s_prow[warp];
    double* diag = s_diag[warp];
#pragma unroll 1
    for (long long item = (long long)blockIdx.x * GFW_WARPS + warp; item < n_items; item += (long long)gridDim.x * GFW_WARPS) {
        const long long k = item / a.nE;
        const int e = (int)(item - k * a.nE);
        const cplx z = cmake(a.E[e], a.eta);
        const cplx* Hr = a.H + (size_t)k * N * N + (size_t)min(lane, N - 1) * N;
        cplx B[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const cplx h = Hr[min(j, N - 1)];
            cplx v = cmake(-h.x, -h.y);
            if (j == lane) { v.x += z.x; v.y += z.y; }
            if (j >= N || lane >= N) v = cmake(j == lane ? 1.0 : 0.0, 0.0);
            B[j] = v;
        }
        bool used = false;
        int my_step = 0;
#pragma unroll 1
        for (int c = 0; c < 32; ++c) {
            const cplx f = B[0];
            const float mag = fmaxf((float)(fabs(f.x) + fabs(f.y)), 1.17549435e-38f);
            const unsigned key = used ? 0u : ((__float_as_uint(mag) & 0xffffffe0u) | (unsigned)(31 - lane));
            const int p = 31 - (int)(__reduce_max_sync(0xffffffffu, key) & 31u);
            const bool is_p = lane == p;
            if (is_p) { used = true; my_step = c; }
            const cplx piv = cmake(__shfl_sync(0xffffffffu, f.x, p), __shfl_sync(0xffffffffu, f.y, p));
            const double rd = fast_rcp(fma(piv.x, piv.x, piv.y * piv.y));
            const cplx inv = cmake(piv.x * rd, -piv.y * rd);
            cplx g = cmul(f, inv);
            if (is_p) g = cmake(1.0 - inv.x, -inv.y);
            __syncwarp();                                   // the previous step's readers of prow are done
            if (is_p) {
#pragma unroll
                for (int m = 1; m < 32; ++m) prow[m] = B[m];
            }
            __syncwarp();
#pragma unroll
            for (int m = 1; m < 32; ++m) {
                const cplx pm = prow[m];
                cplx v = B[m];
                cfms(v, g, pm);
                B[m - 1] = v;
            }
            B[31] = is_p ? inv : cmake(-g.x, -g.y);
        }
        // lane r was the pivot row of step i = my_step:  G_ii = R[p_i][q_i] = register step_of_row[i] of this lane
        const int q = __shfl_sync(0xffffffffu, my_step, my_step);
        double gi = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) gi = j == q ? B[j].y : gi;
        __syncwarp();
        diag[my_step] = gi;
        __syncwarp();
        for (int g = 0; g < ng; ++g) {
            double s = 0.0;
            if (a.n_groups > 0) {
                for (int t = a.gptr[g] + lane; t < a.gptr[g + 1]; t += 32) s += diag[a.grows[t]];
            } else {
                s = lane < N ? diag[lane] : 0.0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.part[((size_t)k * ng + g) * a.nE + e] = -s / 3.141592653589793;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// 32 < N <= 96 (C4: N = 64): one matrix per 256-thread CTA with the matrix DISTRIBUTED OVER THE REGISTERS of the CTA - thread
// (warp w, lane l) owns rows w + 8 t (t < TR) x columns l + 32 u (u < TC), 16 complex values at N = 64 - so the rank-1 update
// touches no shared memory for the matrix (the shared-memory kernel above moves 32 B per 4 DFMA through the LSU and is bound by
// it).  Per step: the owners of column c stage it (double-buffered), barrier, every warp repeats the pivot search on the staged
// column, the owners of row p stage the raw pivot row, barrier, update with the folded factors g_r of the warp kernel.
// ---------------------------------------------------------------------------------------------------------------------------
template <int TR, int TC>
__global__ void __launch_bounds__(GF_THREADS, (TR * TC <= 16 ? 2 : 1)) gf_inverse_cta_kernel(GfArgs a) {
    constexpr int NR = 8 * TR, NCOL = 32 * TC;
    __shared__ __align__(16) cplx s_fcol[2][NR];
    __shared__ __align__(16) cplx s_prow[NCOL];
    __shared__ int s_p_of_step[NR], s_step_of_row[NR];
    __shared__ double s_diag[NR];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = GF_THREADS / 32;
    const int ng = a.n_groups > 0 ? a.n_groups : 1;
    const long long n_items = a.nk * a.nE;
#pragma unroll 1
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long k = item / a.nE;
        const int e = (int)(item - k * a.nE);
        const cplx z = cmake(a.E[e], a.eta);
        const cplx* Hk = a.H + (size_t)k * N * N;
        cplx B[TR][TC];
#pragma unroll
        for (int t = 0; t < TR; ++t) {
            const int r = warp + 8 * t;
#pragma unroll
            for (int u = 0; u < TC; ++u) {
                const int j = lane + 32 * u;
                const cplx h = Hk[(size_t)min(r, N - 1) * N + min(j, N - 1)];
                cplx v = cmake(-h.x, -h.y);
                if (r == j) { v.x += z.x; v.y += z.y; }
                if (r >= N || j >= N) v = cmake(0.0, 0.0);
                B[t][u] = v;
            }
        }
        unsigned long long used = 0ull;                    // bit q: row lane + 32 q has been a pivot row (identical in every warp)
#pragma unroll 1
        for (int c = 0; c < N; ++c) {
            cplx* fcol = s_fcol[c & 1];
            const int uc = c >> 5;
            if (lane == (c & 31)) {
#pragma unroll
                for (int t = 0; t < TR; ++t) {
                    cplx v = B[t][0];
#pragma unroll
                    for (int u = 1; u < TC; ++u) v = uc == u ? B[t][u] : v;
                    fcol[warp + 8 * t] = v;
                }
            }
            __syncthreads();
            double best = -1.0; int bp = 0x7fffffff;
            for (int r = lane, q = 0; r < N; r += 32, ++q) {
                if ((used >> q) & 1ull) continue;
                const cplx v = fcol[r];
                const double m = fabs(v.x) + fabs(v.y);
                if (m > best) { best = m; bp = r; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
            }
            const int p = bp;
            if ((p & 31) == lane) used |= 1ull << (p >> 5);
            if (warp == (p & 7)) {
                const int tp = p >> 3;
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    cplx v = B[0][u];
#pragma unroll
                    for (int t = 1; t < TR; ++t) v = tp == t ? B[t][u] : v;
                    s_prow[lane + 32 * u] = v;
                }
            }
            if (tid == 0) { s_p_of_step[c] = p; s_step_of_row[p] = c; }
            const cplx piv = fcol[p];
            const double rd = 1.0 / fma(piv.x, piv.x, piv.y * piv.y);
            const cplx inv = cmake(piv.x * rd, -piv.y * rd);
            __syncthreads();
            cplx pr[TC];
#pragma unroll
            for (int u = 0; u < TC; ++u) pr[u] = s_prow[lane + 32 * u];
#pragma unroll
            for (int t = 0; t < TR; ++t) {
                const int r = warp + 8 * t;
                cplx g = cmul(fcol[r], inv);
                if (r == p) g = cmake(1.0 - inv.x, -inv.y);
                const cplx colv = r == p ? inv : cmake(-g.x, -g.y);
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    cplx v = B[t][u];
                    cfms(v, g, pr[u]);
                    B[t][u] = (lane + 32 * u) == c ? colv : v;
                }
            }
        }
        // read-out: G_ii = R[p_i][q_i]; the owner of (row r, column step_of_row[i]) with i = step_of_row[r] (p_i = r) publishes it
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TR; ++t) {
            const int r = warp + 8 * t;
            if (r < N) {
                const int i = s_step_of_row[r];
                const int jj = s_step_of_row[i];
                if ((jj & 31) == lane) {
                    double v = B[t][0].y;
#pragma unroll
                    for (int u = 1; u < TC; ++u) v = (jj >> 5) == u ? B[t][u].y : v;
                    s_diag[i] = v;
                }
            }
        }
        __syncthreads();
        for (int g = warp; g < ng; g += n_warps) {
            double s = 0.0;
            if (a.n_groups > 0) {
                for (int q = a.gptr[g] + lane; q < a.gptr[g + 1]; q += 32) s += s_diag[a.grows[q]];
            } else {
                for (int i = lane; i < N; i += 32) s += s_diag[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.part[((size_t)k * ng + g) * a.nE + e] = -s / 3.141592653589793;
        }
        __syncthreads();                                   // s_diag / step tables are rewritten by the next item
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Tile layout (32 < N <= 64, C4: four warps per matrix; the one-warp instantiation for N <= 32 is kept behind SKTB_GF_TILE1=1 -
// measured slower than the row-per-lane kernel, 99 against 76 ms on T: 440 instead of 250 instructions per step).  The row-per-lane kernel
// above broadcasts the pivot row out of ONE lane: 31 single-lane STS.128 + 31 LDS.128 per step, and ncu shows the warps waiting on
// the MIO queue for a third of their time.  Here a lane owns an 8 x 4 TILE of the matrix - rows a + AR t (t < 8), columns
// 4 (BC w + b) + u (u < 4), lane = (a, b) with AR x BC = 32 row / column groups per warp - so both broadcast vectors are spread
// over several lanes: the pivot column is staged by the AR lanes that hold it (8 STS.128) and read back per row group (8 LDS.128),
// the pivot row by the BC lanes of its row group (4 STS.128, warp-uniform switch on the tile row) and read back per column group
// (4 LDS.128): 24 multi-lane LSU instructions per 128 DFMA.  The register window rotates by one column per step in EVERY lane
// (period 4), which keeps the column of step c in register 0 of column group (c / 4); with four warps per matrix only the staged
// pivot column crosses warps: one CTA barrier per step (double-buffered), the pivot row stays inside each warp.
// ---------------------------------------------------------------------------------------------------------------------------
template <int NW>
struct GfTile {
    static constexpr int AR = NW == 1 ? 4 : 8;        // row groups per warp
    static constexpr int BC = 32 / AR;                 // column groups per warp
    static constexpr int NP = 8 * AR;                  // padded order: 32 or 64
    static constexpr int MPC = 4 / NW;                 // matrices per 128-thread CTA
    struct Smem {
        cplx fcol[2][NP];
        cplx prow[NW][4 * BC];
        double diag[NP];
        int p_of_step[NP], step_of_row[NP];
        cplx piv[2];
        int pivot_row[2];
    };
};

#ifndef SKTB_GF_TILE_MINB
#define SKTB_GF_TILE_MINB 2        // measured: 255 registers / 8 warps per SM beats 168 registers / 12 warps (spills and moves in the step loop)
#endif
template <int NW>
__global__ void __launch_bounds__(128, SKTB_GF_TILE_MINB) gf_inverse_tile_kernel(GfArgs a) {
    using T = GfTile<NW>;
    constexpr int AR = T::AR, BC = T::BC, NP = T::NP, MPC = T::MPC;
    __shared__ __align__(16) typename T::Smem s_all[MPC];
    const int N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mi = NW == 1 ? warp : 0, w = NW == 1 ? 0 : warp;          // matrix slot in the CTA, warp within the matrix
    typename T::Smem& S = s_all[mi];
    const int ra = lane / BC, cb = lane % BC;                           // row group, column group
    const int col0 = 4 * (BC * w + cb);                                 // first column of this lane's tile
    const int ng = a.n_groups > 0 ? a.n_groups : 1;
    const long long n_items = a.nk * a.nE;
    const int n_steps = (N + 3) & ~3;                                   // whole rotation periods: the window ends in natural order
    auto msync = [&]() { if (NW == 1) __syncwarp(); else __syncthreads(); };
#pragma unroll 1
    for (long long item0 = (long long)blockIdx.x * MPC; item0 < n_items; item0 += (long long)gridDim.x * MPC) {
        long long item = item0 + mi;
        const bool live = item < n_items;
        if (!live) item = n_items - 1;
        const long long k = item / a.nE;
        const int e = (int)(item - k * a.nE);
        const cplx z = cmake(a.E[e], a.eta);
        const cplx* Hk = a.H + (size_t)k * N * N;
        cplx B[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int r = ra + AR * t;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = col0 + u;
                const cplx h = Hk[(size_t)min(r, N - 1) * N + min(j, N - 1)];
                cplx v = cmake(-h.x, -h.y);
                if (r == j) { v.x += z.x; v.y += z.y; }
                if (r >= N || j >= N) v = cmake(r == j ? 1.0 : 0.0, 0.0);
                B[t][u] = v;
            }
        }
        unsigned used = 0u;                                             // bit t: row ra + AR t has been a pivot row
#pragma unroll 1
        for (int c = 0; c < n_steps; ++c) {
            const int buf = c & 1;
            const int gc = c >> 2;                                      // global column group of column c
            const bool in_col = gc == BC * w + cb;
            if (NW == 1 || w == gc / BC) {                              // the warp that holds column c: pivot search + staging
                unsigned key = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const cplx f = B[t][0];
                    const float mag = fmaxf((float)(fabs(f.x) + fabs(f.y)), 1.17549435e-38f);
                    const unsigned kt = (__float_as_uint(mag) & 0xffffffc0u) | (unsigned)(63 - (ra + AR * t));
                    if (in_col && !((used >> t) & 1u)) key = max(key, kt);
                    if (in_col) S.fcol[buf][ra + AR * t] = f;
                }
                key = __reduce_max_sync(0xffffffffu, key);
                __syncwarp();                                       // the staging stores are visible to lane 0
                if (lane == 0) {
                    const int p = 63 - (int)(key & 63u);
                    S.pivot_row[buf] = p; S.p_of_step[c] = p; S.step_of_row[p] = c;
                    const cplx pv = S.fcol[buf][p];
                    S.piv[buf] = pv;
                    S.fcol[buf][p] = cmake(pv.x - 1.0, pv.y);           // (piv - 1) / piv = 1 - 1/piv: the pivot row's factor needs no special case
                }
            }
            msync();
            const int p = S.pivot_row[buf];
            const int ap = p % AR, tp = p / AR;
            if (ra == ap) used |= 1u << tp;
            const cplx piv = S.piv[buf];
            const double rd = fast_rcp(fma(piv.x, piv.x, piv.y * piv.y));
            const cplx inv = cmake(piv.x * rd, -piv.y * rd);
            if (ra == ap) {                                             // this row group holds the pivot row: stage its 4 columns per lane
                cplx* dst = S.prow[w] + 4 * cb;
                switch (tp) {                                           // warp-uniform
                case 0: dst[0] = B[0][0]; dst[1] = B[0][1]; dst[2] = B[0][2]; dst[3] = B[0][3]; break;
                case 1: dst[0] = B[1][0]; dst[1] = B[1][1]; dst[2] = B[1][2]; dst[3] = B[1][3]; break;
                case 2: dst[0] = B[2][0]; dst[1] = B[2][1]; dst[2] = B[2][2]; dst[3] = B[2][3]; break;
                case 3: dst[0] = B[3][0]; dst[1] = B[3][1]; dst[2] = B[3][2]; dst[3] = B[3][3]; break;
                case 4: dst[0] = B[4][0]; dst[1] = B[4][1]; dst[2] = B[4][2]; dst[3] = B[4][3]; break;
                case 5: dst[0] = B[5][0]; dst[1] = B[5][1]; dst[2] = B[5][2]; dst[3] = B[5][3]; break;
                case 6: dst[0] = B[6][0]; dst[1] = B[6][1]; dst[2] = B[6][2]; dst[3] = B[6][3]; break;
                default: dst[0] = B[7][0]; dst[1] = B[7][1]; dst[2] = B[7][2]; dst[3] = B[7][3]; break;
                }
            }
            __syncwarp();
            cplx pr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) pr[u] = S.prow[w][4 * cb + u];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const bool is_p = ra == ap && t == tp;
                const cplx g = cmul(S.fcol[buf][ra + AR * t], inv);
                cplx first = B[t][0];
                cfms(first, g, pr[0]);
#pragma unroll
                for (int u = 1; u < 4; ++u) {
                    cplx v = B[t][u];
                    cfms(v, g, pr[u]);
                    B[t][u - 1] = v;
                }
                const cplx colv = cmake((is_p ? 1.0 : 0.0) - g.x, -g.y);   // pivot row: 1 - (1 - 1/piv) = 1/piv; other rows: -f_r / piv
                B[t][3] = in_col ? colv : first;
            }
            __syncwarp();                                               // prow is rewritten by the next step
        }
        // read-out: G_ii = R[p_i][q_i] - for row r = p_i (i = step_of_row[r]) the needed column is q_i = step_of_row[i]
        msync();
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int r = ra + AR * t;
            const int i = r < N ? S.step_of_row[r] : 0;             // rows of the identity padding beyond the last step were never pivots
            const int q = S.step_of_row[i];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r < N && col0 + u == q) S.diag[i] = B[t][u].y;
        }
        msync();
        for (int g = w; g < ng; g += NW) {
            double s = 0.0;
            if (a.n_groups > 0) {
                for (int q = a.gptr[g] + lane; q < a.gptr[g + 1]; q += 32) s += S.diag[a.grows[q]];
            } else {
                for (int i = lane; i < N; i += 32) s += S.diag[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && live) a.part[((size_t)k * ng + g) * a.nE + e] = -s / 3.141592653589793;
        }
        msync();                                                        // diag / step tables are rewritten by the next item
    }
}

inline size_t gf_smem_bytes(int N, bool in_smem) {
    return (in_smem ? (size_t)N * (N + 1) * sizeof(cplx) : 0) + (size_t)2 * N * sizeof(cplx) + (size_t)2 * N * sizeof(int);
}

}  // namespace sktb
