// Dependent-issue latencies on B200 (one warp, one CTA): DFMA, DADD, DMMA m8n8k4, SHFL+DADD, DFMA with an LDS operand, rsqrt
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(double* out, long long* clk, const double* in) {
    __shared__ double sm[64];
    sm[threadIdx.x] = in[threadIdx.x & 7]; sm[32 + threadIdx.x] = 1.0;
    __syncthreads();
    double a = in[0], b = in[1] * 1e-9, c = in[2];
    long long t0, t1;
    const int n = 4096;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    t1 = clock64(); clk[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = a + c;
    t1 = clock64(); clk[1] = t1 - t0;
    double d0 = a, d1 = c;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(b), "d"(b));
    t1 = clock64(); clk[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); clk[3] = t1 - t0;
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { a = fma(a, b, sm[idx]); idx = (idx + (a > 1e300)) & 31; }
    t1 = clock64(); clk[4] = t1 - t0;
    double e = in[3];
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); e = fma(e, b, c); }
    t1 = clock64(); clk[5] = t1 - t0;
    double r = in[4];
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r)); r = y + 1.5; }
    t1 = clock64(); clk[6] = t1 - t0;
    out[threadIdx.x] = a + d0 + d1 + e + r;
}
int main() {
    double *out, *in; long long* clk;
    cudaMalloc(&out, 256); cudaMalloc(&in, 64); cudaMalloc(&clk, 64);
    double h[8] = {1.0, 1.0, 0.5, 2.0, 3.0, 1, 1, 1}; cudaMemcpy(in, h, 64, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(out, clk, in); k<<<1, 32>>>(out, clk, in); cudaDeviceSynchronize();
    long long c[8]; cudaMemcpy(c, clk, 64, cudaMemcpyDeviceToHost);
    const char* nm[7] = {"DFMA dependent", "DADD dependent", "DMMA dependent", "SHFL+DADD", "DFMA with LDS operand", "2 DFMA chains (per pair)", "rsqrt.f64 + DADD"};
    for (int i = 0; i < 7; ++i) printf("%-28s %.1f clk\n", nm[i], c[i] / 4096.0);
    return 0;
}
