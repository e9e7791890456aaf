// Micro-benchmarks that bound what an FP64 register-resident kernel can reach on B200 (sm_100a):
//   chains     : 8 independent DFMA chains, two operands reused (the sktb_fp64_peak pattern)
//   distinct   : DFMA with three distinct, rotating register operands (no .reuse), 32 accumulators
//   rank1_reg  : B[q] -= v * conj(p[q]) on 32 complex accumulators, p in registers
//   rank1_lds  : same, p[q] broadcast from shared memory (LDS.128 per 4 DFMA)   -> row-per-lane kernels
//   rank1_lds2 : two rows per lane: LDS.128 per 8 DFMA                            -> pair-layout kernels
// Reports DFMA warp-instructions per clock per SM (2.0 = full rate) for several occupancies.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define NACC 32
template <int MODE>
__global__ void __launch_bounds__(128, 2) probe(double* out, const double* in, int iters, long long* clk) {
    extern __shared__ double2 sm[];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) sm[t] = make_double2(in[t & 63] * 1e-3, in[(t + 7) & 63] * 1e-3);
    __syncthreads();
    double2 B[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) B[q] = make_double2(in[q] + threadIdx.x, in[q + 1]);
    double2 v = make_double2(in[40] * 1e-3, in[41] * 1e-3), v2 = make_double2(in[42] * 1e-3, in[43] * 1e-3);
    double zz[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    double2 zacc[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    long long t0 = clock64();
    const double2* sp = sm + (threadIdx.x >> 5) * 8;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { B[q].x = fma(B[q].x, v.x, v.y); B[q].y = fma(B[q].y, v.x, v.y); }
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                const int a = (q + 5) % NACC, b = (q + 11) % NACC;
                B[q].x = fma(B[a].y, B[b].x, B[q].x); B[q].y = fma(B[a].x, B[b].y, B[q].y);
                B[q].x = fma(B[b].y, B[a].x, B[q].x); B[q].y = fma(B[b].x, B[a].y, B[q].y);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                const double2 p = B[(q + 16) % NACC];
                B[q].x = fma(-v.x, p.x, B[q].x); B[q].y = fma(-v.y, p.x, B[q].y);
                B[q].x = fma(-v.y, p.y, B[q].x); B[q].y = fma(v.x, p.y, B[q].y);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                const double2 p = sp[q];
                B[q].x = fma(-v.x, p.x, B[q].x); B[q].y = fma(-v.y, p.x, B[q].y);
                B[q].x = fma(-v.y, p.y, B[q].x); B[q].y = fma(v.x, p.y, B[q].y);
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double2 p = sp[q];
                B[q].x = fma(-v.x, p.x, B[q].x); B[q].y = fma(-v.y, p.x, B[q].y);
                B[q].x = fma(-v.y, p.y, B[q].x); B[q].y = fma(v.x, p.y, B[q].y);
                B[q + 16].x = fma(-v2.x, p.x, B[q + 16].x); B[q + 16].y = fma(-v2.y, p.x, B[q + 16].y);
                B[q + 16].x = fma(-v2.y, p.y, B[q + 16].x); B[q + 16].y = fma(v2.x, p.y, B[q + 16].y);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double2 p = sp[q + 16];
                B[q].x = fma(-v.x, p.x, B[q].x); B[q].y = fma(-v.y, p.x, B[q].y);
                B[q].x = fma(-v.y, p.y, B[q].x); B[q].y = fma(v.x, p.y, B[q].y);
                B[q + 16].x = fma(-v2.x, p.x, B[q + 16].x); B[q + 16].y = fma(-v2.y, p.x, B[q + 16].y);
                B[q + 16].x = fma(-v2.y, p.y, B[q + 16].x); B[q + 16].y = fma(v2.x, p.y, B[q + 16].y);
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double2 p = sp[q];
                const double2 x = sp[q + 32];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double2 w = r ? v2 : v;
                    double2 u = B[q + 16 * r];
                    u.x = fma(-w.x, p.x, u.x); u.y = fma(-w.y, p.x, u.y);
                    u.x = fma(-w.y, p.y, u.x); u.y = fma(w.x, p.y, u.y);
                    zacc[r].x = fma(u.x, x.x, zacc[r].x); zacc[r].x = fma(-u.y, x.y, zacc[r].x);
                    zacc[r].y = fma(u.x, x.y, zacc[r].y); zacc[r].y = fma(u.y, x.x, zacc[r].y);
                    B[q + 16 * r] = u;
                }
            }
        } else if (MODE == 6 || MODE == 7) {
            // split accumulators: zz[r][0..3] = sum u.x x.x, u.y x.y, u.x x.y, u.y x.x
            if (MODE == 6) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const double2 p = sp[q];
                    const double2 x = sp[q + 32];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double2 w = r ? v2 : v;
                        double2 u = B[q + 16 * r];
                        u.x = fma(-w.x, p.x, u.x); u.y = fma(-w.y, p.x, u.y);
                        u.x = fma(-w.y, p.y, u.x); u.y = fma(w.x, p.y, u.y);
                        zz[r][0] = fma(u.x, x.x, zz[r][0]); zz[r][1] = fma(u.y, x.y, zz[r][1]);
                        zz[r][2] = fma(u.x, x.y, zz[r][2]); zz[r][3] = fma(u.y, x.x, zz[r][3]);
                        B[q + 16 * r] = u;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const double2 p = sp[q];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double2 w = r ? v2 : v;
                        double2 u = B[q + 16 * r];
                        u.x = fma(-w.x, p.x, u.x); u.y = fma(-w.y, p.x, u.y);
                        u.x = fma(-w.y, p.y, u.x); u.y = fma(w.x, p.y, u.y);
                        B[q + 16 * r] = u;
                    }
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const double2 x = sp[q + 32];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double2 u = B[q + 16 * r];
                        zz[r][0] = fma(u.x, x.x, zz[r][0]); zz[r][1] = fma(u.y, x.y, zz[r][1]);
                        zz[r][2] = fma(u.x, x.y, zz[r][2]); zz[r][3] = fma(u.y, x.x, zz[r][3]);
                    }
                }
            }
        }
    }
    long long t1 = clock64();
    if (MODE >= 6) { B[0].x += zz[0][0] + zz[0][1] + zz[0][2] + zz[0][3] + zz[1][0] + zz[1][1] + zz[1][2] + zz[1][3]; }
    if (MODE == 5) { B[0].x += zacc[0].x + zacc[0].y + zacc[1].x + zacc[1].y; }
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc += B[q].x + B[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int dfma_per_iter, int ctas_per_sm, int sms, double* out, double* in, long long* clk) {
    const int iters = 2000, threads = 128;
    int grid = sms * ctas_per_sm;
    size_t smem = (size_t)(220 * 1024 / ctas_per_sm) - 2048;   // limits residency to ctas_per_sm
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE>, threads, smem);
    probe<MODE><<<grid, threads, smem>>>(out, in, 10, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, threads, smem>>>(out, in, iters, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double warp_dfma = (double)dfma_per_iter * iters * (threads / 32) * ctas_per_sm;   // per SM
    printf("%-11s ctas/SM=%d (occ %d, %2d warps/SM) : %.3f DFMA/clk/SM (clock64)  %.2f TFLOP/s (events)  err=%s\n", name, ctas_per_sm, occ,
           ctas_per_sm * 4, warp_dfma / (double)h[0], 2.0 * 32 * warp_dfma * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in; long long* clk;
    cudaMalloc(&out, sizeof(double) * 148 * 16 * 128); cudaMalloc(&in, 64 * 8); cudaMalloc(&clk, 8 * 148 * 16);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + 0.01 * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int c : {1, 2}) {
        run<0>("chains", 16 * 8, c, sms, out, in, clk);
        run<1>("distinct", 4 * NACC, c, sms, out, in, clk);
        run<2>("rank1_reg", 4 * NACC, c, sms, out, in, clk);
        run<3>("rank1_lds", 4 * NACC, c, sms, out, in, clk);
        run<4>("rank1_lds2", 8 * NACC, c, sms, out, in, clk);
        run<5>("fused_pair", 16 * 16, c, sms, out, in, clk);
        run<6>("fused_split", 16 * 16, c, sms, out, in, clk);
        run<7>("two_pass", 16 * 16, c, sms, out, in, clk);
    }
    return 0;
}
