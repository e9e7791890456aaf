// Is the FP64 tensor instruction (DMMA, mma.sync.m8n8k4.f64) on B200 a second FP64 pipe next to the DFMA pipe?
//   mode 0 : DFMA only (8 independent chains per thread, operands reused)
//   mode 1 : DMMA m8n8k4 only (8 independent accumulator tiles per warp)
//   mode 2 : both in one warp, equal flops (1 DMMA per 8 DFMA)
//   mode 3 : warps alternate: even warps DFMA, odd warps DMMA
//   mode 4 : DMMA m16n8k8 only (if the assembler takes it for sm_100a)
// Prints TFLOP/s (2 flops per FMA) for several occupancies.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#ifdef HAVE_M16
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
#endif

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, const double* in, int iters) {
    double acc[8], t[8][2];
    const double a = in[threadIdx.x & 31] * 1e-3, b = in[(threadIdx.x + 3) & 31] * 1e-3;
#pragma unroll
    for (int q = 0; q < 8; ++q) { acc[q] = in[q] + threadIdx.x; t[q][0] = in[q + 8]; t[q][1] = in[q + 9]; }
    const bool odd = (threadIdx.x >> 5) & 1;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fma(acc[q], a, b);
            }
        }
        if (MODE == 1 || (MODE == 3 && odd)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dmma884(t[q][0], t[q][1], a, b);
        }
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dmma884(t[q][0], t[q][1], a, b);
        }
#ifdef HAVE_M16
        if (MODE == 4) {
            double aa[4] = {a, b, a, b}, bb[2] = {b, a};
            dmma1688(&t[0][0], aa, bb); dmma1688(&t[2][0], aa, bb); dmma1688(&t[4][0], aa, bb); dmma1688(&t[6][0], aa, bb);
        }
#endif
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += acc[q] + t[q][0] + t[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double flops_per_warp_iter, int ctas_per_sm, int threads, double* out, const double* in, int sms) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * ctas_per_sm, threads>>>(out, in, 100);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<sms * ctas_per_sm, threads>>>(out, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warps = double(sms) * ctas_per_sm * threads / 32;
    printf("%-28s ctas/SM %d x %3d thr : %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ctas_per_sm, threads, best,
           warps * iters * flops_per_warp_iter / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024); cudaMalloc(&in, 64 * sizeof(double));
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-3 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double dfma = 64.0 * 32 * 2, dmma = 8.0 * 256 * 2;   // per warp and iteration
    for (int c = 1; c <= 8; c *= 2) {
        run<0>("DFMA only", dfma, c, 256, out, in, sms);
        run<1>("DMMA m8n8k4 only", dmma, c, 256, out, in, sms);
        run<2>("DFMA+DMMA same warp", dfma + dmma, c, 256, out, in, sms);
        run<3>("DFMA / DMMA alternate warps", (dfma + dmma) / 2, c, 256, out, in, sms);
#ifdef HAVE_M16
        run<4>("DMMA m16n8k8 only", 4.0 * 16 * 8 * 8 * 2, c, 256, out, in, sms);
#endif
    }
    return 0;
}
