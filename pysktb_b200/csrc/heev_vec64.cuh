// K2b for 32 < N <= 64 with eigenvectors (config C4): the counterpart of heev_vec32.cuh.
//   vecql64_kernel : one warp per matrix, lane r owns rows r and r+32 of Z (shared memory, column-major, stride 65),
//                    implicit-shift QL with accumulation, ranks, Z^T to global ([comp][vector], 32 KB per k);
//   vecbt64_kernel : one 128-thread CTA per matrix, thread = (eigenvector n, row half h), 32 complex components in
//                    registers, reflectors of tridiag_cta64_kernel broadcast from global memory, dot products
//                    completed with one lane-pair shuffle; output [band][k][comp] (hamiltonian.py:124).
// (included by heev_small.cuh inside namespace sktb::small)
#pragma once

constexpr int VEC64_ZSTRIDE = 65;
constexpr int VEC64_SMEM_DOUBLES = 64 * VEC64_ZSTRIDE + 128;     // Z | d | e

struct Vec64Args {
    const double* de;        // [nk][2][64]
    const cplx* vstore;      // [nk][C64_VSTORE_CPLX]
    long long nk; int N;
    double* evals; cplx* evecs; long long ld, col0;
    int* fail;
};

template <int UNUSED>
__global__ void __launch_bounds__(32) vecql64_kernel(Vec64Args a, double* __restrict__ zstore, int* __restrict__ rankstore,
                                                      const int* __restrict__ only_flagged) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    double* Z = reinterpret_cast<double*>(smem_raw);
    double* d = Z + 64 * VEC64_ZSTRIDE;
    double* e = d + 64;
    const int n = a.N;
    const double eps = 2.220446049250313e-16;
#pragma unroll 1
    for (long long kq = blockIdx.x; kq < a.nk; kq += gridDim.x) {
        if (only_flagged && !only_flagged[kq]) continue;
        for (int t = lane; t < 64; t += 32) {
            d[t] = a.de[(size_t)kq * 128 + t];
            e[t] = (t < n - 1) ? a.de[(size_t)kq * 128 + 64 + t] : 0.0;
        }
#pragma unroll 4
        for (int c = 0; c < 64; ++c) { Z[c * VEC64_ZSTRIDE + lane] = (c == lane) ? 1.0 : 0.0; Z[c * VEC64_ZSTRIDE + lane + 32] = (c == lane + 32) ? 1.0 : 0.0; }
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
                double dnext = d[m];
#pragma unroll 1
                for (int i = m - 1; i >= l; --i) {
                    const double ei = e[i], di = d[i];
                    const double za1 = Z[(i + 1) * VEC64_ZSTRIDE + lane], za0 = Z[i * VEC64_ZSTRIDE + lane];
                    const double zb1 = Z[(i + 1) * VEC64_ZSTRIDE + lane + 32], zb0 = Z[i * VEC64_ZSTRIDE + lane + 32];
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
                    Z[(i + 1) * VEC64_ZSTRIDE + lane] = fma(s, za0, c * za1);
                    Z[i * VEC64_ZSTRIDE + lane] = fma(c, za0, -s * za1);
                    Z[(i + 1) * VEC64_ZSTRIDE + lane + 32] = fma(s, zb0, c * zb1);
                    Z[i * VEC64_ZSTRIDE + lane + 32] = fma(c, zb0, -s * zb1);
                }
                __syncwarp();
                if (under) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
                __syncwarp();
            }
        }
        __syncwarp();
        if (bad && lane == 0) atomicAdd(a.fail, 1);
        for (int t = lane; t < 64; t += 32) {
            const double mine = (t < n) ? d[t] : 0.0;
            int rank = 0;
            for (int q = 0; q < n; ++q) {
                const double o = d[q];
                rank += (o < mine || (o == mine && q < t)) ? 1 : 0;
            }
            if (t < n) a.evals[(long long)rank * a.ld + a.col0 + kq] = mine;
            rankstore[(size_t)kq * 64 + t] = rank;
        }
        double* zs = zstore + (size_t)kq * 4096;
#pragma unroll 2
        for (int c = 0; c < 64; ++c) {                              // zs[comp][vector]
            zs[c * 64 + lane] = Z[lane * VEC64_ZSTRIDE + c];
            zs[c * 64 + lane + 32] = Z[(lane + 32) * VEC64_ZSTRIDE + c];
        }
        __syncwarp();
    }
}

// reflectors J0+3 .. J0 applied to the thread's 32 components (rows 32h .. 32h+31); LI0 = first local row that a
// reflector of this phase can touch in the upper half (h = 1) - the lower half runs the same trip count on zeros
template <int J0>
SKTB_D void vec64_backtransform4(cplx (&u)[32], const cplx* __restrict__ vst, const cplx* __restrict__ taust, const int N, const int h) {
    constexpr int LI0 = J0 >= 32 ? J0 + 1 - 32 : 0;
    // reflectors 32.. live entirely in rows 33..63: the row half h = 0 contributes nothing (and, when the trailing
    // block was tridiagonalised by the 32-wide kernel, its rows 0..31 of the store are not written at all)
    const bool idle = J0 >= 32 && h == 0;
#pragma unroll 1
    for (int j = J0 + 3; j >= J0; --j) {
        if (j > N - 2) continue;
        const cplx* v = vst + (size_t)j * 64 + 32 * h;
        cplx dot = cmake(0.0, 0.0), dot2 = cmake(0.0, 0.0);
        if (!idle) {
#pragma unroll
            for (int i = LI0; i < 32; i += 2) {
                cfmac(dot, v[i], u[i]);
                if (i + 1 < 32) cfmac(dot2, v[i + 1], u[i + 1]);
            }
        }
        dot = cadd(dot, dot2);
        dot.x += __shfl_xor_sync(0xffffffffu, dot.x, 1);
        dot.y += __shfl_xor_sync(0xffffffffu, dot.y, 1);
        const cplx t = cmul(taust[j], dot);
        if (!idle) {
#pragma unroll
            for (int i = LI0; i < 32; ++i) cfms(u[i], t, v[i]);
        }
    }
}

template <int UNUSED>
__global__ void __launch_bounds__(128, 2) vecbt64_kernel(Vec64Args a, const double* __restrict__ zstore, const int* __restrict__ rankstore) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* vsm = reinterpret_cast<cplx*>(smem_raw);                      // this matrix's reflectors + tau (65 KB), staged coalesced
    const int t = threadIdx.x, nvec = t >> 1, h = t & 1;
    const int n = a.N;
#pragma unroll 1
    for (long long kq = blockIdx.x; kq < a.nk; kq += gridDim.x) {
        const cplx* vst = a.vstore + (size_t)kq * C64_VSTORE_CPLX;
        __syncthreads();
#pragma unroll 8
        for (int q = t; q < C64_VSTORE_CPLX; q += 128) vsm[q] = vst[q];
        const double* zs = zstore + (size_t)kq * 4096;
        cplx u[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] = cmake((nvec < n && 32 * h + i < n) ? zs[(32 * h + i) * 64 + nvec] : 0.0, 0.0);
        const int rank = rankstore[(size_t)kq * 64 + nvec];
        __syncthreads();
        const cplx* taust = vsm + 4096;
        vec64_backtransform4<60>(u, vsm, taust, n, h); vec64_backtransform4<56>(u, vsm, taust, n, h);
        vec64_backtransform4<52>(u, vsm, taust, n, h); vec64_backtransform4<48>(u, vsm, taust, n, h);
        vec64_backtransform4<44>(u, vsm, taust, n, h); vec64_backtransform4<40>(u, vsm, taust, n, h);
        vec64_backtransform4<36>(u, vsm, taust, n, h); vec64_backtransform4<32>(u, vsm, taust, n, h);
        vec64_backtransform4<28>(u, vsm, taust, n, h); vec64_backtransform4<24>(u, vsm, taust, n, h);
        vec64_backtransform4<20>(u, vsm, taust, n, h); vec64_backtransform4<16>(u, vsm, taust, n, h);
        vec64_backtransform4<12>(u, vsm, taust, n, h); vec64_backtransform4<8>(u, vsm, taust, n, h);
        vec64_backtransform4<4>(u, vsm, taust, n, h); vec64_backtransform4<0>(u, vsm, taust, n, h);
        if (nvec < n) {
            cplx* out = a.evecs + (size_t)((long long)rank * a.ld + a.col0 + kq) * n + 32 * h;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (32 * h + i < n) out[i] = u[i];
        }
    }
}

// Record-and-replay QL for 32 < N <= 64 (see heev_vec32.cuh): tqlrec_kernel<64, VECREC64_CAP> records, this kernel
// replays on Z (one matrix per warp, lane r owns rows r and r+32; a sweep = header + at most 63 rotations = one
// 64-entry window, double-buffered); flagged overflows are redone by vecql64_kernel.
constexpr int VECREC64_CAP = 8192;
constexpr int VECAPPLY64_SMEM_DOUBLES = 64 * VEC64_ZSTRIDE + 64 + 2 * 64 * 2;      // Z | d | windows [2][64] x 16 B

template <int UNUSED>
__global__ void __launch_bounds__(32) vecapply64_kernel(Vec64Args a, const double* __restrict__ dfinal, const double2* __restrict__ lists,
                                                         const int* __restrict__ ovf_flag, double* __restrict__ zstore, int* __restrict__ rankstore) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    double* Z = reinterpret_cast<double*>(smem_raw);
    double* d = Z + 64 * VEC64_ZSTRIDE;
    double2* win = reinterpret_cast<double2*>(d + 64);
    const int n = a.N;
#pragma unroll 1
    for (long long kq = blockIdx.x; kq < a.nk; kq += gridDim.x) {
        if (ovf_flag[kq]) continue;
        d[lane] = dfinal[(size_t)kq * 64 + lane]; d[lane + 32] = dfinal[(size_t)kq * 64 + lane + 32];
#pragma unroll 4
        for (int c = 0; c < 64; ++c) { Z[c * VEC64_ZSTRIDE + lane] = (c == lane) ? 1.0 : 0.0; Z[c * VEC64_ZSTRIDE + lane + 32] = (c == lane + 32) ? 1.0 : 0.0; }
        __syncwarp();
        const double2* list = lists + (size_t)kq * VECREC64_CAP;
        int pos = 0, cur = 0;
        double2 pre0 = list[min(pos + lane, VECREC64_CAP - 1)], pre1 = list[min(pos + lane + 32, VECREC64_CAP - 1)];
#pragma unroll 1
        while (true) {
            win[cur * 64 + lane] = pre0; win[cur * 64 + lane + 32] = pre1;
            __syncwarp();
            const double2 hdr = win[cur * 64];
            const int m = (int)hdr.x, count = (int)hdr.y;
            if (m < 0) break;
            pos += 1 + count;
            pre0 = list[min(pos + lane, VECREC64_CAP - 1)]; pre1 = list[min(pos + lane + 32, VECREC64_CAP - 1)];
            double upa = Z[m * VEC64_ZSTRIDE + lane], upb = Z[m * VEC64_ZSTRIDE + lane + 32];
#pragma unroll 4
            for (int t = 0; t < count; ++t) {
                const double2 cs = win[cur * 64 + 1 + t];
                const int i = m - 1 - t;
                const double ga = Z[i * VEC64_ZSTRIDE + lane], gb = Z[i * VEC64_ZSTRIDE + lane + 32];
                Z[(i + 1) * VEC64_ZSTRIDE + lane] = fma(cs.y, ga, cs.x * upa);
                Z[(i + 1) * VEC64_ZSTRIDE + lane + 32] = fma(cs.y, gb, cs.x * upb);
                upa = fma(cs.x, ga, -cs.y * upa);
                upb = fma(cs.x, gb, -cs.y * upb);
            }
            if (count > 0) { Z[(m - count) * VEC64_ZSTRIDE + lane] = upa; Z[(m - count) * VEC64_ZSTRIDE + lane + 32] = upb; }
            cur ^= 1;
        }
        __syncwarp();
        for (int t = lane; t < 64; t += 32) {
            const double mine = (t < n) ? d[t] : 0.0;
            int rank = 0;
            for (int q = 0; q < n; ++q) {
                const double o = d[q];
                rank += (o < mine || (o == mine && q < t)) ? 1 : 0;
            }
            if (t < n) a.evals[(long long)rank * a.ld + a.col0 + kq] = mine;
            rankstore[(size_t)kq * 64 + t] = rank;
        }
        double* zs = zstore + (size_t)kq * 4096;
#pragma unroll 2
        for (int c = 0; c < 64; ++c) {
            zs[c * 64 + lane] = Z[lane * VEC64_ZSTRIDE + c];
            zs[c * 64 + lane + 32] = Z[(lane + 32) * VEC64_ZSTRIDE + c];
        }
        __syncwarp();
    }
}

