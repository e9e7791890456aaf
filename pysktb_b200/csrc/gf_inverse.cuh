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

inline size_t gf_smem_bytes(int N, bool in_smem) {
    return (in_smem ? (size_t)N * (N + 1) * sizeof(cplx) : 0) + (size_t)2 * N * sizeof(cplx) + (size_t)2 * N * sizeof(int);
}

}  // namespace sktb
