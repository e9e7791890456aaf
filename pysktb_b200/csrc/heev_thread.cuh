// K0+K1+K2 for N <= 8 (config C1: sp3 + spin doubling; graphene-class models with SOC): ONE MATRIX PER THREAD.
// For matrices this small the lane-group kernels of heev_small.cuh spend most of their issue slots on cross-lane
// traffic (every reflector entry is a shuffle or a shared-memory broadcast, every norm a butterfly) and half of the
// lanes hold retired rows.  Here a thread owns the packed lower triangle of its H(k) in registers (36 complex for
// NC = 8), the whole Householder tridiagonalisation is straight-line code with static register indices, no shuffle,
// no barrier, no idle lane; shared memory only carries the k-independent tables (uniform-address broadcast loads) and
// the (d, e) of the per-lane QL (odd stride, conflict-free).  One kernel: K0 (mesh point) + K1 (bond phases,
// h(k), kron(h, I2) + soc_mat) + K2 (zhetd2(UPLO='L') semantics + square-root-free QL + sort) + coalesced stores along
// k (the reference's [band, k] layout).  FP64 CUDA cores only.
//
// Bond phases: exp(2 pi i k.d) through sincospi(2 k.d) (exact argument reduction, no multiplication by a rounded 2 pi),
// and a bond whose vector is the bitwise negative of an earlier bond of the table takes the complex conjugate of that
// bond's phase (k.(-d) = -(k.d) exactly) - half of the sincos evaluations for centrosymmetric neighbour shells.
// (included by heev_small.cuh inside namespace sktb::small)
#pragma once

constexpr int THREAD_CTA = 128;

SKTB_D constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }      // packed lower triangle, j <= i

// Householder tridiagonalisation of the Hermitian matrix whose lower triangle (real diagonal) is A, LAPACK zhetd2('L').
template <int NC>
SKTB_D void thread_tridiag(cplx (&A)[NC * (NC + 1) / 2], double (&d)[NC], double (&e)[NC]) {
#pragma unroll
    for (int j = 0; j < NC - 1; ++j) {
        d[j] = A[tri(j, j)].x;
        double xn = 0.0;
#pragma unroll
        for (int i = j + 2; i < NC; ++i) { xn = fma(A[tri(i, j)].x, A[tri(i, j)].x, xn); xn = fma(A[tri(i, j)].y, A[tri(i, j)].y, xn); }
        double beta; cplx tau, scale;
        pair_larfg(A[tri(j + 1, j)], xn, beta, tau, scale);
        e[j] = beta;
        if (j == NC - 2) break;                                   // the last reflector only fixes e[NC-2]
        cplx v[NC];                                               // v[i], i > j: v[j+1] = 1
        v[j + 1] = cmake(1.0, 0.0);
#pragma unroll
        for (int i = j + 2; i < NC; ++i) v[i] = cmul(A[tri(i, j)], scale);
        // p = tau * A22 v  (Hermitian mat-vec from the lower triangle)
        cplx p[NC];
#pragma unroll
        for (int i = j + 1; i < NC; ++i) {
            cplx acc = cmake(A[tri(i, i)].x * v[i].x, A[tri(i, i)].x * v[i].y);
#pragma unroll
            for (int k = j + 1; k < i; ++k) cfma(acc, A[tri(i, k)], v[k]);          // lower part of row i
#pragma unroll
            for (int k = i + 1; k < NC; ++k) cfmac(acc, A[tri(k, i)], v[k]);         // upper part = conj of column i
            p[i] = cmul(tau, acc);
        }
        // w = p - (tau/2) (p^H v) v
        cplx dot = cmake(0.0, 0.0);
#pragma unroll
        for (int i = j + 1; i < NC; ++i) cfmac(dot, p[i], v[i]);
        const cplx a2 = cmul(cmake(-0.5 * tau.x, -0.5 * tau.y), dot);
#pragma unroll
        for (int i = j + 1; i < NC; ++i) cfma(p[i], a2, v[i]);                       // p now holds w
        // A22 -= v w^H + w v^H  (lower triangle; the diagonal stays real)
#pragma unroll
        for (int i = j + 1; i < NC; ++i) {
#pragma unroll
            for (int k = j + 1; k < i; ++k) { cfms(A[tri(i, k)], v[i], cconj(p[k])); cfms(A[tri(i, k)], p[i], cconj(v[k])); }
            A[tri(i, i)].x -= 2.0 * fma(v[i].x, p[i].x, v[i].y * p[i].y);
        }
    }
    d[NC - 1] = A[tri(NC - 1, NC - 1)].x;
    e[NC - 1] = 0.0;
}

// SPLIT: stop after the tridiagonalisation and leave (d, e) in a.scratch ([k][2][NC], the record tql_small_kernel reads):
// the per-lane QL is a long dependent chain that wants many resident warps, the tridiagonalisation wants 200 registers.
template <int NC, bool SOC, bool SPLIT>
__global__ void __launch_bounds__(THREAD_CTA) solve_thread_kernel(SmallArgs a, double* __restrict__ evals, long long ld, long long col0, int* fail) {
    constexpr int NH = SOC ? NC / 2 : NC;               // orbital-space size bound
    constexpr int NT = NH * (NH + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tables, bonds in the order a.t.thread_order lists them (a bond followed by its inversion partner):
    // Ltri [nb][NT] packed lower hoppings (T_b[mu][nu], nu <= mu; zero beyond n) | bond vectors | onsite | conj flags | soc lower
    double* Ltri = reinterpret_cast<double*>(smem_raw);
    double* Bv = Ltri + (size_t)nb * NT;
    double* On = Bv + 3 * (size_t)nb;
    int* Cs = reinterpret_cast<int*>(On + NH);
    size_t off = (size_t)((reinterpret_cast<double*>(Cs) + (nb + 1) / 2) - Ltri);
    off = (off + 1) & ~(size_t)1;
    cplx* Sl = reinterpret_cast<cplx*>(Ltri + off);     // [NC (NC+1)/2] packed lower soc_mat (SOC only)
    double* qd = reinterpret_cast<double*>(Sl + (SOC ? NC * (NC + 1) / 2 : 0)) + (size_t)warp * 2 * NC * DSTRIDE;
    double* qe = qd + NC * DSTRIDE;
    for (int t = tid; t < nb * NT; t += THREAD_CTA) {
        const int sl = t / NT, q = t - sl * NT, b = a.t.thread_order[sl];
        int mu = 0;
        while ((mu + 1) * (mu + 2) / 2 <= q) ++mu;
        const int nu = q - mu * (mu + 1) / 2;
        Ltri[t] = (mu < n) ? a.t.Lt[((size_t)b * n + nu) * n + mu] : 0.0;             // Lt[b][kappa = nu][iota = mu] = T_b[mu][nu]
    }
    for (int t = tid; t < 3 * nb; t += THREAD_CTA) Bv[t] = a.t.bvec[3 * a.t.thread_order[t / 3] + t % 3];
    for (int t = tid; t < nb; t += THREAD_CTA) Cs[t] = a.t.thread_order[nb + t];       // 1: conjugate of the previous slot's phase
    for (int t = tid; t < NH; t += THREAD_CTA) On[t] = t < n ? a.t.onsite[t] : 0.0;
    if (SOC) {
        for (int t = tid; t < NC * (NC + 1) / 2; t += THREAD_CTA) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= t) ++i;
            const int j = t - i * (i + 1) / 2;
            Sl[t] = (i < N) ? a.t.St[(size_t)j * N + i] : cmake(0.0, 0.0);             // St[col * N + row], mirrored lower: row i >= col j
        }
    }
    __syncthreads();

#pragma unroll 1
    for (long long kbase = ((long long)blockIdx.x * (THREAD_CTA / 32) + warp) * 32; kbase < a.nk; kbase += (long long)gridDim.x * THREAD_CTA) {
        const long long kq = min(kbase + lane, a.nk - 1);
        double k0, k1, k2;
        if (a.k) { k0 = a.k[3 * kq]; k1 = a.k[3 * kq + 1]; k2 = a.k[3 * kq + 2]; }
        else kmesh_point(a.first + kq, a.mesh0, a.mesh1, a.mesh2, k0, k1, k2);
        const double c0 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[0]), __dmul_rn(k1, a.t.rec[3])), __dmul_rn(k2, a.t.rec[6]));
        const double c1 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[1]), __dmul_rn(k1, a.t.rec[4])), __dmul_rn(k2, a.t.rec[7]));
        const double c2 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[2]), __dmul_rn(k1, a.t.rec[5])), __dmul_rn(k2, a.t.rec[8]));
        // ---- K1: packed lower triangle of h(k) ----
        double hr[NT], hi[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q) { hr[q] = 0.0; hi[q] = 0.0; }
        double sn = 0.0, cs = 1.0;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            if (Cs[b]) sn = -sn;                                                       // inversion partner of the previous slot: conj
            else {
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn(c0, Bv[3 * b]), __dmul_rn(c1, Bv[3 * b + 1])), __dmul_rn(c2, Bv[3 * b + 2]));
                sincospi(2.0 * dot, &sn, &cs);
            }
            const double* T = Ltri + (size_t)b * NT;
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                const double t = T[q];
                hr[q] = fma(t, cs, hr[q]); hi[q] = fma(t, sn, hi[q]);
            }
        }
        // ---- H(k) = kron(h, I2) + soc_mat | h, lower triangle, real diagonal (zheevd UPLO='L' reading) ----
        cplx A[NC * (NC + 1) / 2];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                cplx v;
                if (SOC) {
                    const int mu = i >> 1, nu = j >> 1;
                    v = ((i & 1) == (j & 1)) ? cmake(hr[tri(mu, nu)], hi[tri(mu, nu)]) : cmake(0.0, 0.0);
                    if (i == j) v.x += On[mu];
                    const cplx s = Sl[tri(i, j)];
                    v = cmake(v.x + s.x, v.y + s.y);
                } else {
                    v = cmake(hr[tri(i, j)], hi[tri(i, j)]);
                    if (i == j) v.x += On[i];
                }
                if (i == j) v.y = 0.0;
                A[tri(i, j)] = v;
            }
        }
        // ---- K2: tridiagonalisation in registers, per-lane QL on (d, e) in shared memory, sort, store ----
        double d[NC], e[NC];
        thread_tridiag<NC>(A, d, e);
        if constexpr (SPLIT) {
            if (kbase + lane < a.nk) {
                double2* out = reinterpret_cast<double2*>(a.scratch + (size_t)(kbase + lane) * (2 * NC));
#pragma unroll
                for (int i = 0; i < NC; i += 2) { out[i / 2] = make_double2(d[i], d[i + 1]); out[NC / 2 + i / 2] = make_double2(e[i], e[i + 1]); }
                if (a.flags) a.flags[a.flag_off + kbase + lane] = 0;
            }
        } else {
#pragma unroll
        for (int i = 0; i < NC; ++i) { qd[i * DSTRIDE + lane] = d[i]; qe[i * DSTRIDE + lane] = e[i]; }
        if (tql_lane_pwk(qd + lane, qe + lane, N)) atomicAdd(fail, 1);
        constexpr int P2 = NC <= 4 ? 4 : (NC <= 8 ? 8 : 16);             // the sorting network needs a power of two
        double ev[P2];
#pragma unroll
        for (int q = 0; q < P2; ++q) ev[q] = (q < N && q < NC) ? qd[q * DSTRIDE + lane] : __longlong_as_double(0x7ff0000000000000LL);
        bitonic_sort<P2>(ev);
        if (kbase + lane < a.nk) {
#pragma unroll
            for (int q = 0; q < NC; ++q)
                if (q < N) evals[(long long)q * ld + col0 + kbase + lane] = ev[q];
            if (a.flags) a.flags[a.flag_off + kbase + lane] = 0;
        }
        }
    }
}

// host side: slot order [nb] followed by conj flags [nb]: every bond is followed by its inversion partner (the bond of
// the table whose vector is the bitwise negative), which then takes the conjugate phase
inline void thread_bond_order(const double* bvec, int nb, std::vector<int>* out) {
    out->assign(2 * (size_t)nb, 0);
    std::vector<char> used(nb, 0);
    int s = 0;
    for (int b = 0; b < nb; ++b) {
        if (used[b]) continue;
        used[b] = 1;
        (*out)[s++] = b;
        for (int c = b + 1; c < nb; ++c) {
            if (used[c]) continue;
            if (bvec[3 * c] == -bvec[3 * b] && bvec[3 * c + 1] == -bvec[3 * b + 1] && bvec[3 * c + 2] == -bvec[3 * b + 2]) {
                used[c] = 1;
                (*out)[nb + s] = 1;
                (*out)[s++] = c;
                break;
            }
        }
    }
}

inline size_t thread_smem_bytes(int NC, bool soc, int nb, bool split) {
    const int NH = soc ? NC / 2 : NC, NT = NH * (NH + 1) / 2;
    size_t dbl = (size_t)nb * NT + 3 * (size_t)nb + NH + (nb + 1) / 2;
    dbl = (dbl + 1) & ~(size_t)1;
    return dbl * 8 + (soc ? (size_t)NC * (NC + 1) / 2 * 16 : 0) + (split ? 0 : (size_t)(THREAD_CTA / 32) * 2 * NC * DSTRIDE * 8);
}
