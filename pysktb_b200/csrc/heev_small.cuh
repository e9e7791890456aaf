// K1+K2 fused, register-resident, for N <= 32 (configs C1-C3 and the north-star 32-orbital model):
// H(k) is assembled straight into registers and never touches HBM; per k-point the only traffic is
// 24 B of k in (0 B for meshes: K0 is fused too) and 8*N B of eigenvalues out.
//
// Mapping.  L = 4/8/16/32 lanes own one matrix (32/L matrices per warp side by side); lane i holds row i of the
// mirrored-lower Hermitian matrix  A_eff[i][k] = H[i][k] (k<i), conj(H[k][i]) (k>i), Re H[i][i]  - LAPACK
// zheevd(UPLO='L') semantics, which is what the reference's numpy call reads (hamiltonian.py:119).  The
// mirrored tables are prepared on the host (build_tables) so no cross-lane transposition is needed:
//   Lt[b][kappa][iota] = T_b[iota][kappa] (kappa <= iota) | T_b[kappa][iota] (kappa > iota, used with conj phase).
// Householder tridiagonalisation is fully unrolled over the column index (static register indexing);
// v and w are exchanged through a 1 KB shared-memory buffer per warp (broadcast LDS.128), norms and dots
// by xor-shuffles inside the L-lane group.  After 32 matrices the warp transposes (d,e) through shared
// memory and every lane runs the implicit-shift QL of ITS OWN matrix (no idle lanes, rsqrt-based rotations),
// sorts its eigenvalues with a register bitonic network and the warp stores them coalesced along k.
// FP64 CUDA cores only: the work is batched rank-2 updates on 32x32 Hermitian matrices, not a GEMM.
#pragma once
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
    double rec[9];
    double asym_bound = 0.0;
    int soc_any = 0;
};

struct SmallArgs {
    SmallTables t;
    const double* k;            // [nk][3] or NULL -> mesh
    int mesh0, mesh1, mesh2;
    long long first, nk;
    int N, soc;
    double* evals; long long ld, col0;
    int* flags; int* fail;
};

inline bool supports(int N) { return N >= 1 && N <= 32; }

constexpr int SMALL_WARPS = 4;
constexpr int DSTRIDE = 33;   // odd stride: conflict-free both for the per-row stores and the per-lane QL

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
SKTB_D int tql_lane(double* d, double* e, int n) {
    const double eps = 2.220446049250313e-16;
#define D_(i) d[(i) * DSTRIDE]
#define E_(i) e[(i) * DSTRIDE]
    E_(n - 1) = 0.0;
    int fail = 0;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            double dm = fabs(D_(m));
            for (; m < n - 1; ++m) {
                double dn = fabs(D_(m + 1));
                if (fabs(E_(m)) <= eps * (dm + dn)) break;
                dm = dn;
            }
            if (m == l) break;
            if (++iter > 80) { fail = 1; break; }
            double el = E_(l), dl = D_(l);
            double g = (D_(l + 1) - dl) / (2.0 * el);
            double r = sqrt(fma(g, g, 1.0));
            g = D_(m) - dl + el / (g + (g >= 0.0 ? r : -r));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            bool under = false;
            double di1 = D_(m);                       // d[i+1]
            for (; i >= l; --i) {
                double ei = E_(i), di = D_(i);
                double f = s * ei, b = c * ei;
                double h2 = fma(f, f, g * g);
                if (h2 == 0.0) { E_(i + 1) = 0.0; D_(i + 1) = di1 - p; E_(m) = 0.0; under = true; break; }
                double ir = rsqrt(h2);
                r = h2 * ir;
                E_(i + 1) = r;
                s = f * ir; c = g * ir;
                g = di1 - p;
                r = fma(di - g, s, 2.0 * c * b);
                p = s * r;
                D_(i + 1) = g + p;
                g = fma(c, r, -b);
                di1 = di;
            }
            if (under) continue;
            D_(l) = di1 - p; E_(l) = g; E_(m) = 0.0;
        }
    }
#undef D_
#undef E_
    return fail;
}

template <int L>
SKTB_D double group_sum(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// NC: compile-time column count (>= N, multiple of 4); L: lanes per matrix (power of two >= NC);
// SOC: kron(h, I2) + soc_mat assembly; GATE: evaluate the reference's Hermiticity gate per k.
// ---------------------------------------------------------------------------------------------------------
template <int NC, int L, bool SOC, bool GATE>
__global__ void __launch_bounds__(SMALL_WARPS * 32) heev_small_kernel(SmallArgs a) {
    constexpr int G = 32 / L;                 // matrices side by side in a warp
    constexpr int P2 = L;                     // bitonic width
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = lane & (L - 1), g = lane / L;

    // ---- shared memory carve-up: CTA-wide tables, then per-warp scratch ---------------------------------
    double* sLt = reinterpret_cast<double*>(smem_raw);
    double* sDt = sLt + (size_t)nb * n * n;
    double* sOn = sDt + (GATE ? (size_t)nb * n * n : 0);
    double* sBv = sOn + n;
    double* sSA = sBv + 3 * nb;
    size_t off = (size_t)((sSA + ((GATE && SOC) ? N * N : 0)) - sLt);
    off = (off + 1) & ~(size_t)1;              // 16-byte align
    cplx* sSt = reinterpret_cast<cplx*>(sLt + off);
    cplx* wbase = sSt + (SOC ? N * N : 0);
    const size_t per_warp_cplx = (size_t)2 * 32 + (size_t)G * nb + (size_t)NC * DSTRIDE;   // v,w | phases | d,e (2*NC*DSTRIDE doubles)
    cplx* myw = wbase + per_warp_cplx * warp;
    cplx* vbuf = myw;
    cplx* wbuf = myw + 32;
    cplx* phs = myw + 64 + (size_t)g * nb;
    double* dbuf = reinterpret_cast<double*>(myw + 64 + (size_t)G * nb);
    double* ebuf = dbuf + NC * DSTRIDE;

    for (int t = tid; t < nb * n * n; t += blockDim.x) { sLt[t] = a.t.Lt[t]; if (GATE) sDt[t] = a.t.Dt[t]; }
    for (int t = tid; t < n; t += blockDim.x) sOn[t] = a.t.onsite[t];
    for (int t = tid; t < 3 * nb; t += blockDim.x) sBv[t] = a.t.bvec[t];
    if (SOC) for (int t = tid; t < N * N; t += blockDim.x) { sSt[t] = a.t.St[t]; if (GATE) sSA[t] = a.t.SAt[t]; }
    __syncthreads();

    const long long n_batches = (a.nk + 31) / 32;
    const long long gwarp = (long long)blockIdx.x * SMALL_WARPS + warp, nwarps = (long long)gridDim.x * SMALL_WARPS;
    const bool row_ok = i < N;
    const int mu = SOC ? (i >> 1) : i;
    const int spin = i & 1;

    for (long long batch = gwarp; batch < n_batches; batch += nwarps) {
        const long long kbase = batch * 32;
#pragma unroll 1
        for (int it = 0; it < L; ++it) {
            const int slot = it * G + g;
            long long kq = kbase + slot;
            const bool k_ok = kq < a.nk;
            if (!k_ok) kq = a.nk - 1;
            double k0, k1, k2;
            if (a.k) { k0 = a.k[3 * kq]; k1 = a.k[3 * kq + 1]; k2 = a.k[3 * kq + 2]; }
            else kmesh_point(a.first + kq, a.mesh0, a.mesh1, a.mesh2, k0, k1, k2);

            // ---- bond phases (one sincos per lane) ----
            {
                double c0 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[0]), __dmul_rn(k1, a.t.rec[3])), __dmul_rn(k2, a.t.rec[6]));
                double c1 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[1]), __dmul_rn(k1, a.t.rec[4])), __dmul_rn(k2, a.t.rec[7]));
                double c2 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[2]), __dmul_rn(k1, a.t.rec[5])), __dmul_rn(k2, a.t.rec[8]));
                for (int b = i; b < nb; b += L) {
                    double dot = __dadd_rn(__dadd_rn(__dmul_rn(c0, sBv[3 * b]), __dmul_rn(c1, sBv[3 * b + 1])), __dmul_rn(c2, sBv[3 * b + 2]));
                    cplx ph;
                    sincos(__dmul_rn(6.283185307179586, dot), &ph.y, &ph.x);
                    phs[b] = ph;
                }
            }
            __syncwarp();

            // ---- K1: row i of the mirrored-lower H(k), straight into registers ----
            cplx A[NC];
            double gate_max = -1.0;
            if (SOC) {
#pragma unroll
                for (int nu = 0; nu < NC / 2; ++nu) {
                    double re = 0.0, im = 0.0, gsum = 0.0;
                    if (row_ok && nu < n) {
                        for (int b = 0; b < nb; ++b) {
                            double t = sLt[((size_t)b * n + nu) * n + mu];
                            cplx ph = phs[b];
                            re = fma(t, ph.x, re); im = fma(t, ph.y, im);
                            if (GATE) gsum = fma(sDt[((size_t)b * n + nu) * n + mu], ph.x, gsum);
                        }
                        if (nu == mu) { re += sOn[mu]; im = 0.0; }
                        else if (nu > mu) im = -im;
                    }
                    cplx s0 = cmake(0.0, 0.0), s1 = s0;
                    if (row_ok && 2 * nu < N) {
                        s0 = sSt[(size_t)(2 * nu) * N + i]; s1 = sSt[(size_t)(2 * nu + 1) * N + i];
                        if (GATE) {
                            gate_max = fmax(gate_max, (spin == 0 ? gsum : 0.0) + sSA[(size_t)(2 * nu) * N + i]);
                            gate_max = fmax(gate_max, (spin == 1 ? gsum : 0.0) + sSA[(size_t)(2 * nu + 1) * N + i]);
                        }
                    }
                    A[2 * nu] = cmake((spin == 0 ? re : 0.0) + s0.x, (spin == 0 ? im : 0.0) + s0.y);
                    A[2 * nu + 1] = cmake((spin == 1 ? re : 0.0) + s1.x, (spin == 1 ? im : 0.0) + s1.y);
                }
            } else {
#pragma unroll
                for (int nu = 0; nu < NC; ++nu) {
                    double re = 0.0, im = 0.0, gsum = 0.0;
                    if (row_ok && nu < n) {
                        for (int b = 0; b < nb; ++b) {
                            double t = sLt[((size_t)b * n + nu) * n + mu];
                            cplx ph = phs[b];
                            re = fma(t, ph.x, re); im = fma(t, ph.y, im);
                            if (GATE) gsum = fma(sDt[((size_t)b * n + nu) * n + mu], ph.x, gsum);
                        }
                        if (nu == mu) { re += sOn[mu]; im = 0.0; }
                        else if (nu > mu) im = -im;
                        if (GATE) gate_max = fmax(gate_max, gsum);
                    }
                    A[nu] = cmake(re, im);
                }
            }
            if (GATE) {
                bool bad = gate_max > 1.0e-9;
                unsigned m = __ballot_sync(0xffffffffu, bad);
                unsigned gm = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (g * L));
                if (i == 0 && k_ok && a.flags) a.flags[a.col0 + kq] = (m & gm) ? 1 : 0;
            } else {
                if (i == 0 && k_ok && a.flags) a.flags[a.col0 + kq] = 0;
            }

            // ---- K2a: Householder tridiagonalisation, columns unrolled ----
            double dval = 0.0, eval = 0.0;
#pragma unroll
            for (int j = 0; j < NC - 1; ++j) {
                const cplx x = A[j];
                const bool below = i > j + 1;
                double xn = below ? fma(x.x, x.x, x.y * x.y) : 0.0;
                xn = group_sum<L>(xn);
                cplx alpha;
                alpha.x = __shfl_sync(0xffffffffu, x.x, j + 1, L);
                alpha.y = __shfl_sync(0xffffffffu, x.y, j + 1, L);
                dval = (i == j) ? x.x : dval;
                double beta; cplx tau, scale;
                householder(alpha, xn, beta, tau, scale);
                eval = (i == j) ? beta : eval;
                if (j == NC - 2) break;                       // the last reflector only fixes e[NC-2]
                cplx v = (i == j + 1) ? cmake(1.0, 0.0) : (below ? cmul(x, scale) : cmake(0.0, 0.0));
                vbuf[lane] = v;
                __syncwarp();
                cplx acc0 = cmake(0.0, 0.0), acc1 = acc0;
#pragma unroll
                for (int k = j + 1; k < NC; ++k) {
                    cplx vk = vbuf[g * L + k];
                    if ((k - j) & 1) cfma(acc0, A[k], vk); else cfma(acc1, A[k], vk);
                }
                cplx p = cmul(tau, cadd(acc0, acc1));
                if (i <= j) p = cmake(0.0, 0.0);
                cplx dp = cmulc(p, v);
                dp.x = group_sum<L>(dp.x); dp.y = group_sum<L>(dp.y);
                cplx a2 = cmul(cmake(-0.5 * tau.x, -0.5 * tau.y), dp);
                cplx w = p;
                cfma(w, a2, v);
                wbuf[lane] = w;
                __syncwarp();
                const cplx vc = v, wc = w;
#pragma unroll
                for (int k = j + 1; k < NC; ++k) {
                    cplx vk = vbuf[g * L + k], wk = wbuf[g * L + k];
                    cplx t = A[k];
                    // A[i][k] -= v_i conj(w_k) + w_i conj(v_k)
                    t.x = fma(-vc.x, wk.x, t.x); t.x = fma(-vc.y, wk.y, t.x);
                    t.y = fma(-vc.y, wk.x, t.y); t.y = fma(vc.x, wk.y, t.y);
                    t.x = fma(-wc.x, vk.x, t.x); t.x = fma(-wc.y, vk.y, t.x);
                    t.y = fma(-wc.y, vk.x, t.y); t.y = fma(wc.x, vk.y, t.y);
                    A[k] = t;
                }
                __syncwarp();
            }
            dval = (i == NC - 1) ? A[NC - 1].x : dval;
            if (i < NC) { dbuf[i * DSTRIDE + slot] = dval; ebuf[i * DSTRIDE + slot] = eval; }
            __syncwarp();
        }

        // ---- K2b: one tridiagonal matrix per lane ----
        if (tql_lane(dbuf + lane, ebuf + lane, N)) atomicAdd(a.fail, 1);
        double ev[P2];
#pragma unroll
        for (int q = 0; q < P2; ++q) ev[q] = (q < N && q < NC) ? dbuf[q * DSTRIDE + lane] : __longlong_as_double(0x7ff0000000000000LL);
        bitonic_sort<P2>(ev);
        if (kbase + lane < a.nk) {
#pragma unroll
            for (int q = 0; q < P2; ++q)
                if (q < N) a.evals[(long long)q * a.ld + a.col0 + kbase + lane] = ev[q];
        }
        __syncwarp();
    }
}

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
    if (table_doubles(n, nb, 2 * n <= 32 ? 2 * n : n, 2 * n <= 32, true) * 8 > 110 * 1024) { *why = "dense bond tables exceed the shared-memory budget"; return false; }
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
    if (!up(Lt, &out->Lt, owned) || !up(Dt, &out->Dt, owned) || !up(bvec, &out->bvec, owned) || !up(ons, &out->onsite, owned) ||
        !up(St, &out->St, owned) || !up(SAt, &out->SAt, owned)) { *why = "cudaMalloc failed"; return false; }
    return true;
}

template <int NC, int L, bool SOC, bool GATE>
static int launch_one(const SmallArgs& a, cudaStream_t st, std::string* info) {
    constexpr int G = 32 / L;
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N;
    size_t smem = table_doubles(n, nb, N, SOC, GATE) * 8 + (size_t)SMALL_WARPS * ((size_t)64 + (size_t)G * nb + (size_t)NC * DSTRIDE) * 16;
    auto kern = heev_small_kernel<NC, L, SOC, GATE>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(smem=" + std::to_string(smem) + ") failed"; return -1;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SMALL_WARPS * 32, smem);
    if (occ < 1) { *info = "kernel does not fit (smem=" + std::to_string(smem) + ")"; return -1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long batches = (a.nk + 31) / 32;
    long long want = (batches + SMALL_WARPS - 1) / SMALL_WARPS;
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sms * occ));
    kern<<<grid, SMALL_WARPS * 32, smem, st>>>(a);
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[256];
    snprintf(buf, sizeof buf, "heev_small<NC=%d,L=%d,soc=%d,gate=%d> N=%d grid=%u x %d thr, smem=%zuB, %d CTA/SM", NC, L, (int)SOC, (int)GATE, N, grid,
             SMALL_WARPS * 32, smem, occ);
    *info = buf;
    return 1;
}

template <int NC, int L>
static int launch_nc(const SmallArgs& a, bool gate, cudaStream_t st, std::string* info) {
    if (a.soc) return gate ? launch_one<NC, L, true, true>(a, st, info) : launch_one<NC, L, true, false>(a, st, info);
    return gate ? launch_one<NC, L, false, true>(a, st, info) : launch_one<NC, L, false, false>(a, st, info);
}

// returns number of kernels launched (1) or -1 on error (text in *info)
int launch_solve(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N,
                 double* evals, long long ld, long long col0, int* flags, int* fail, cudaStream_t st, std::string* info);

}  // namespace small
}  // namespace sktb
