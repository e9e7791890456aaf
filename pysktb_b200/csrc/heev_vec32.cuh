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
constexpr int VEC32_SMEM_DOUBLES = 32 * VEC32_ZSTRIDE + 32 + 128;      // Z | d | e (vecql32) or rotation windows [2][32] x 16 B (vecapply32)   per warp

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
__global__ void __launch_bounds__(VECQL_WARPS * 32, 5) vecql32_kernel(Vec32Args a, double* __restrict__ zstore, int* __restrict__ rankstore,
                                                                       const int* __restrict__ only_flagged) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Z = reinterpret_cast<double*>(smem_raw) + (size_t)warp * VEC32_SMEM_DOUBLES;
    double* d = Z + 32 * VEC32_ZSTRIDE;
    double* e = d + 32;
    const int n = a.N;
    const double eps = 2.220446049250313e-16;
#pragma unroll 1
    for (long long kq = (long long)blockIdx.x * VECQL_WARPS + warp; kq < a.nk; kq += (long long)gridDim.x * VECQL_WARPS) {
        if (only_flagged && !only_flagged[kq]) continue;              // fallback pass after tqlrec32 / vecapply32
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx* vsm = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * PAIR_VSTORE_CPLX;     // this warp's reflectors + tau (16.5 KB)
    const int n = a.N;
#pragma unroll 1
    for (long long kq = (long long)blockIdx.x * VEC32_WARPS + warp; kq < a.nk; kq += (long long)gridDim.x * VEC32_WARPS) {
        const cplx* vst = a.vstore + (size_t)kq * PAIR_VSTORE_CPLX;
        __syncwarp();
#pragma unroll 11
        for (int t = 0; t < PAIR_VSTORE_CPLX / 32; ++t) vsm[t * 32 + lane] = vst[t * 32 + lane];   // coalesced, 33 loads in flight
        const double* zs = zstore + (size_t)kq * 1024;
        cplx u[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] = cmake((lane < n && i < n) ? zs[i * 32 + lane] : 0.0, 0.0);
        const int rank = rankstore[(size_t)kq * 32 + lane];
        __syncwarp();
        vec32_all_phases(u, vsm, vsm + 1024, n);
        if (lane < n) {
            cplx* out = a.evecs + (size_t)((long long)rank * a.ld + a.col0 + kq) * n;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < n) out[i] = u[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// QL with accumulation without repeating the scalar recurrence on 32 lanes (vecql32 spends its FP64 issue slots on
// that):  tqlrec32 runs the eigenvalue-only QL of tql_small (one matrix per LANE) and records every rotation
// (c, s) sweep by sweep in a per-matrix list in global memory;  vecapply32 (one matrix per WARP, lane = row of Z)
// replays the list on Z in shared memory: 1 broadcast load, 1 LDS, 1 STS and 4 FP64 per rotation.
// List format (16-byte entries): sweep header {m, count} followed by count rotations {c, s} for i = m-1, m-2, ...;
// terminator {-1, 0}.  A matrix whose list would overflow VECREC_CAP is flagged and redone by vecql32_kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int VECREC_CAP = 2048;

SKTB_D int tql_lane_rec(double* d, double* e, int n, double2* list, const int cap, int& ovf) {
    const double eps = 2.220446049250313e-16;
#define D_(i) d[(i) * DSTRIDE]
#define E_(i) e[(i) * DSTRIDE]
    E_(n - 1) = 0.0;
    int fail = 0, l = 0, iter = 0, pos = 0;
    bool rec = true;
    int m = tql_scan(d, e, 0, n);
    while (true) {
        if (m == l || iter > 80) {
            if (m != l) fail = 1;
            ++l; iter = 0;
            if (l >= n) break;
            m = tql_scan(d, e, l, n);
            continue;
        }
        ++iter;
        if (rec && pos + (m - l) + 2 > cap) { rec = false; ovf = 1; }
        const int hp = pos;
        if (rec) ++pos;
        double el = E_(l), dl = D_(l);
        double g = (D_(l + 1) - dl) / (2.0 * el);
        double r = sqrt(fma(g, g, 1.0));
        g = D_(m) - dl + el / (g + (g >= 0.0 ? r : -r));
        double s = 1.0, c = 1.0, p = 0.0;
        bool under = false;
        double di1 = D_(m);
        double dnext = 0.0;
        int mb = m;
        int done = 0;
        for (int i = m - 1; i >= l; --i) {
            double ei = E_(i), di = D_(i);
            double f = s * ei, b = c * ei;
            double h2 = fma(f, f, g * g);
            if (h2 == 0.0) { E_(i + 1) = 0.0; D_(i + 1) = di1 - p; E_(m) = 0.0; under = true; break; }
            double ir = rsqrt(h2);
            r = h2 * ir;
            s = f * ir; c = g * ir;
            if (rec) list[pos++] = make_double2(c, s);
            ++done;
            g = di1 - p;
            double rr = fma(di - g, s, 2.0 * c * b);
            p = s * rr;
            double dnew = g + p;
            D_(i + 1) = dnew;
            double an = fabs(dnew);
            if (i + 1 < m) {
                E_(i + 1) = r;
                if (r <= eps * (an + dnext)) mb = i + 1;
            }
            dnext = an;
            g = fma(c, rr, -b);
            di1 = di;
        }
        if (rec) list[hp] = make_double2((double)m, (double)done);
        if (under) { m = tql_scan(d, e, l, n); continue; }
        double dl_new = di1 - p;
        D_(l) = dl_new; E_(l) = g; E_(m) = 0.0;
        if (fabs(g) <= eps * (fabs(dl_new) + dnext)) { ++l; iter = 0; }
        m = mb;
    }
    if (rec) list[pos] = make_double2(-1.0, 0.0);
#undef D_
#undef E_
    return fail;
}

// one tridiagonal matrix per lane (as tql_small_kernel); writes the final (unsorted) d, the rotation list, the flag
struct TqlRecArgs {
    const double* de; long long nk; int N; int* fail;     // (d, e) scratch [nk][2][NC]
};
template <int NC, int CAP>
__global__ void __launch_bounds__(SMALL_WARPS * 32) tqlrec_kernel(TqlRecArgs a, double* __restrict__ dfinal, double2* __restrict__ lists,
                                                                   int* __restrict__ ovf_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* dbuf = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * NC * DSTRIDE;
    double* ebuf = dbuf + NC * DSTRIDE;
    const long long n_batches = (a.nk + 31) / 32;
    const long long gwarp = (long long)blockIdx.x * SMALL_WARPS + warp, nwarps = (long long)gridDim.x * SMALL_WARPS;
#pragma unroll 1
    for (long long batch = gwarp; batch < n_batches; batch += nwarps) {
        const long long kbase = batch * 32;
        const int cnt = (int)min((long long)32, a.nk - kbase);
        const long long kq = kbase + min(lane, cnt - 1);
        {
            const double2* src = reinterpret_cast<const double2*>(a.de + (size_t)kq * (2 * NC));
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
        int ovf = 0;
        double2* list = lists + (size_t)kq * CAP;
        const int bad = tql_lane_rec(dbuf + lane, ebuf + lane, a.N, list, CAP, ovf);
        if (bad) atomicAdd(a.fail, 1);
        if (lane < cnt) {
            ovf_flag[kq] = ovf;
#pragma unroll 4
            for (int q = 0; q < NC; ++q) dfinal[(size_t)kq * NC + q] = dbuf[q * DSTRIDE + lane];
        }
        __syncwarp();
    }
}

// one matrix per warp, lane = row of Z: replay the rotation list, then ranks / eigenvalues / Z^T exactly as vecql32
template <int UNUSED>
__global__ void __launch_bounds__(VECQL_WARPS * 32, 5) vecapply32_kernel(Vec32Args a, const double* __restrict__ dfinal, const double2* __restrict__ lists,
                                                                          const int* __restrict__ ovf_flag, double* __restrict__ zstore,
                                                                          int* __restrict__ rankstore) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Z = reinterpret_cast<double*>(smem_raw) + (size_t)warp * VEC32_SMEM_DOUBLES;
    double* d = Z + 32 * VEC32_ZSTRIDE;
    const int n = a.N;
#pragma unroll 1
    for (long long kq = (long long)blockIdx.x * VECQL_WARPS + warp; kq < a.nk; kq += (long long)gridDim.x * VECQL_WARPS) {
        if (ovf_flag[kq]) continue;                                   // redone by vecql32_kernel
        d[lane] = dfinal[(size_t)kq * 32 + lane];
#pragma unroll 4
        for (int c = 0; c < 32; ++c) Z[c * VEC32_ZSTRIDE + lane] = (c == lane) ? 1.0 : 0.0;
        __syncwarp();
        // A sweep is a header plus at most 31 rotations: one coalesced 512-byte window (entry pos + lane per lane) holds
        // it.  The window of the NEXT sweep is requested as soon as this sweep's header is known and lands in the other
        // half of the double buffer while the rotations below run from shared memory (broadcast reads).
        const double2* list = lists + (size_t)kq * VECREC_CAP;
        double2* win = reinterpret_cast<double2*>(d + 32);            // [2][32] entries behind d
        int pos = 0, cur = 0;
        double2 pre = list[min(pos + lane, VECREC_CAP - 1)];
#pragma unroll 1
        while (true) {
            win[cur * 32 + lane] = pre;
            __syncwarp();
            const double2 hdr = win[cur * 32];
            const int m = (int)hdr.x, count = (int)hdr.y;
            if (m < 0) break;
            pos += 1 + count;
            pre = list[min(pos + lane, VECREC_CAP - 1)];              // prefetch the next window
            double up = Z[m * VEC32_ZSTRIDE + lane];                  // running value of column i+1
#pragma unroll 4
            for (int t = 0; t < count; ++t) {
                const double2 cs = win[cur * 32 + 1 + t];
                const int i = m - 1 - t;
                const double g = Z[i * VEC32_ZSTRIDE + lane];
                Z[(i + 1) * VEC32_ZSTRIDE + lane] = fma(cs.y, g, cs.x * up);
                up = fma(cs.x, g, -cs.y * up);
            }
            if (count > 0) Z[(m - count) * VEC32_ZSTRIDE + lane] = up;
            cur ^= 1;
        }
        __syncwarp();
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
        for (int c = 0; c < 32; ++c) zs[c * 32 + lane] = Z[lane * VEC32_ZSTRIDE + c];
        __syncwarp();
    }
}

