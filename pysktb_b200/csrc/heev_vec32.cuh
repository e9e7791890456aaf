// K2b for N <= 32 with eigenvectors: one warp per matrix, two kernels (QL with accumulation | back-transformation).
//   in : (d, e) and the Householder reflectors v_j, tau_j kept by the pair-layout tridiagonalisation (heev_pair.cuh)
//   1. implicit-shift QL on the real tridiagonal with the rotations accumulated into Z (lane r owns row r of Z, Z in
//      shared memory column-major with stride 33: both the rotation update (fixed column, lanes = rows) and the
//      later column reads (lanes = eigenvectors) are conflict-free); every lane runs the scalar recurrence
//      redundantly (warp-uniform broadcast loads of d, e);
//   2. ranks by ascending eigenvalue (ties by index), as numpy.linalg.eigh returns them;
//   3. back-transformation U = H_0 H_1 ... H_{N-2} Z with lane n = eigenvector n, the 32 complex components in
//      registers, reflectors broadcast from global memory (L1-resident, 16 KB per matrix), in phases of four
//      reflectors so that register indices stay static;
//   4. eigenvalues [band][k] and eigenvectors [band][k][comp] (rows = vectors: hamiltonian.py:124) written by the
//      owning lane, 512 B contiguous per vector.
// LAPACK semantics: A = Q T Q^H, Q = H_0 ... H_{N-2}, H_j = I - tau_j v_j v_j^H (zhetd2 / zungtr / zsteqr).
// (included by heev_small.cuh inside namespace sktb::small)
#pragma once

constexpr int VEC32_ZSTRIDE = 33;
constexpr int VEC32_WARPS = 4;
constexpr int VEC32_SMEM_DOUBLES = 32 * VEC32_ZSTRIDE + 64;      // Z | d | e   per warp

struct Vec32Args {
    const double* de;        // [nk][2][32]
    const cplx* vstore;      // [nk][PAIR_VSTORE_CPLX]
    long long nk; int N;
    double* evals; cplx* evecs; long long ld, col0;     // evals[band*ld + col0 + k], evecs[(band*ld + col0 + k)*N + comp]
    int* fail;
};

// four reflectors j0+3 .. j0 applied to u (components i > j0 can change; v_j[i] = 0 for i <= j is stored explicitly)
template <int J0>
SKTB_D void vec32_backtransform4(cplx (&u)[32], const cplx* __restrict__ vst, const cplx* __restrict__ taust, const int N) {
#pragma unroll 1
    for (int j = J0 + 3; j >= J0; --j) {
        if (j > N - 2) continue;
        const cplx* v = vst + j * 32;
        cplx dot = cmake(0.0, 0.0), dot2 = cmake(0.0, 0.0);
#pragma unroll
        for (int i = J0 + 1; i < 32; i += 2) {
            cfmac(dot, v[i], u[i]);
            if (i + 1 < 32) cfmac(dot2, v[i + 1], u[i + 1]);
        }
        const cplx t = cmul(taust[j], cadd(dot, dot2));
#pragma unroll
        for (int i = J0 + 1; i < 32; ++i) cfms(u[i], t, v[i]);
    }
}

__device__ __forceinline__ void vec32_all_phases(cplx (&u)[32], const cplx* vst, const cplx* taust, const int N) {
    vec32_backtransform4<28>(u, vst, taust, N);
    vec32_backtransform4<24>(u, vst, taust, N);
    vec32_backtransform4<20>(u, vst, taust, N);
    vec32_backtransform4<16>(u, vst, taust, N);
    vec32_backtransform4<12>(u, vst, taust, N);
    vec32_backtransform4<8>(u, vst, taust, N);
    vec32_backtransform4<4>(u, vst, taust, N);
    vec32_backtransform4<0>(u, vst, taust, N);
}

// ---- kernel 1: QL with accumulation, few registers -> many warps per SM (the scalar recurrence is latency bound).
//      Writes the sorted eigenvalues and Z (column n = eigenvector n of T, [32][32] doubles, 8 KB per k) + ranks.
constexpr int VECQL_WARPS = 4;
template <int UNUSED>
__global__ void __launch_bounds__(VECQL_WARPS * 32, 5) vecql32_kernel(Vec32Args a, double* __restrict__ zstore, int* __restrict__ rankstore) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Z = reinterpret_cast<double*>(smem_raw) + (size_t)warp * VEC32_SMEM_DOUBLES;
    double* d = Z + 32 * VEC32_ZSTRIDE;
    double* e = d + 32;
    const int n = a.N;
    const double eps = 2.220446049250313e-16;
#pragma unroll 1
    for (long long kq = (long long)blockIdx.x * VECQL_WARPS + warp; kq < a.nk; kq += (long long)gridDim.x * VECQL_WARPS) {
        d[lane] = a.de[(size_t)kq * 64 + lane];
        e[lane] = (lane < n - 1) ? a.de[(size_t)kq * 64 + 32 + lane] : 0.0;
#pragma unroll 4
        for (int c = 0; c < 32; ++c) Z[c * VEC32_ZSTRIDE + lane] = (c == lane) ? 1.0 : 0.0;
        __syncwarp();
        int bad = 0;
#pragma unroll 1
        for (int l = 0; l < n; ++l) {
            int iter = 0;
            while (true) {
                int m = l;
                for (; m < n - 1; ++m) {
                    const double dd = fabs(d[m]) + fabs(d[m + 1]);
                    if (fabs(e[m]) <= eps * dd) break;
                }
                if (m == l) break;
                if (++iter > 80) { bad = 1; break; }
                const double el = e[l], dl = d[l];
                double g = (d[l + 1] - dl) / (2.0 * el);
                double r = sqrt(fma(g, g, 1.0));
                g = d[m] - dl + el / (g + (g >= 0.0 ? r : -r));
                double s = 1.0, c = 1.0, p = 0.0;
                bool under = false;
                __syncwarp();
                double dnext = d[m];                   // d[i+1] before this sweep touches it
#pragma unroll 1
                for (int i = m - 1; i >= l; --i) {
                    const double ei = e[i], di = d[i];
                    const double z1 = Z[(i + 1) * VEC32_ZSTRIDE + lane], z0 = Z[i * VEC32_ZSTRIDE + lane];
                    const double f = s * ei, b = c * ei;
                    const double h2 = fma(f, f, g * g);
                    if (h2 == 0.0) { e[i + 1] = 0.0; d[i + 1] = dnext - p; e[m] = 0.0; under = true; break; }
                    const double ir = pair_rsqrt(h2);
                    r = h2 * ir;
                    e[i + 1] = r;
                    s = f * ir; c = g * ir;
                    g = dnext - p;
                    r = fma(di - g, s, 2.0 * c * b);
                    p = s * r;
                    d[i + 1] = g + p;
                    g = fma(c, r, -b);
                    dnext = di;
                    Z[(i + 1) * VEC32_ZSTRIDE + lane] = fma(s, z0, c * z1);
                    Z[i * VEC32_ZSTRIDE + lane] = fma(c, z0, -s * z1);
                }
                __syncwarp();
                if (under) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
                __syncwarp();
            }
        }
        __syncwarp();
        if (bad && lane == 0) atomicAdd(a.fail, 1);
        const double mine = (lane < n) ? d[lane] : 0.0;
        int rank = 0;
        for (int q = 0; q < n; ++q) {
            const double o = d[q];
            rank += (o < mine || (o == mine && q < lane)) ? 1 : 0;
        }
        if (lane < n) a.evals[(long long)rank * a.ld + a.col0 + kq] = mine;
        rankstore[(size_t)kq * 32 + lane] = rank;
        double* zs = zstore + (size_t)kq * 1024;
#pragma unroll 4
        for (int c = 0; c < 32; ++c) zs[c * 32 + lane] = Z[lane * VEC32_ZSTRIDE + c];     // zs[comp][vector]: coalesced both ways
        __syncwarp();
    }
}

// ---- kernel 2: back-transformation, lane = eigenvector (128 registers of components)
template <int UNUSED>
__global__ void __launch_bounds__(VEC32_WARPS * 32, 2) vecbt32_kernel(Vec32Args a, const double* __restrict__ zstore, const int* __restrict__ rankstore) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.N;
#pragma unroll 1
    for (long long kq = (long long)blockIdx.x * VEC32_WARPS + warp; kq < a.nk; kq += (long long)gridDim.x * VEC32_WARPS) {
        const double* zs = zstore + (size_t)kq * 1024;
        cplx u[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] = cmake((lane < n && i < n) ? zs[i * 32 + lane] : 0.0, 0.0);
        const int rank = rankstore[(size_t)kq * 32 + lane];
        const cplx* vst = a.vstore + (size_t)kq * PAIR_VSTORE_CPLX;
        vec32_all_phases(u, vst, vst + 1024, n);
        if (lane < n) {
            cplx* out = a.evecs + (size_t)((long long)rank * a.ld + a.col0 + kq) * n;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < n) out[i] = u[i];
        }
    }
}
