// Host side of the two-stage tridiagonalisation (heev_twostage.cuh): its own translation unit so that the DMMA kernels
// compile separately from the rest of libsktb.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define SKTB_TWOSTAGE_KERNELS
#include "heev_twostage.cuh"

namespace sktb {
namespace twostage {

bool supported(int N) { return N >= 2 * SB && stage1_smem_bytes(N) <= 227 * 1024; }

int launch_tridiag(cplx* H, int N, long long nm, void* scratch, double* de, cudaStream_t st, char* info, size_t info_len, int stages) {
    if (!supported(N) || nm <= 0) return -1;
    const int NR = (N + 7) / 8 * 8;
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    Stage1Args a1{};
    a1.H = H; a1.N = N; a1.nm = nm; a1.NR = NR;
    a1.mp = (((N - SB + 7) / 8) * 8) | 1;
    a1.AB = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * N * ABW * sizeof(cplx);
    a1.Vg = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * NR * SB * sizeof(cplx);
    a1.Vt = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * NR * SB * sizeof(cplx);
    a1.Wg = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * NR * SB * sizeof(cplx);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem1 = stage1_smem_bytes(N);
    if (cudaFuncSetAttribute(stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess) return -1;
    const unsigned grid1 = (unsigned)std::min<long long>(nm, sms);
    if (stages & 1) stage1_kernel<<<grid1, S1_THREADS, smem1, st>>>(a1);
    Stage2Args a2{};
    a2.AB = a1.AB; a2.N = N; a2.nm = nm; a2.de = de;
    const size_t smem2 = (size_t)S2_WARPS * 3 * SB * sizeof(cplx) + (size_t)N * sizeof(int);
    const unsigned grid2 = (unsigned)std::min<long long>(nm, 2LL * sms);
    if (stages & 2) stage2_kernel<<<grid2, S2_WARPS * 32, smem2, st>>>(a2);
    if (info) snprintf(info, info_len, "two-stage: stage1_kernel (band %d, DMMA) grid=%u x %d thr smem=%zuB + stage2_kernel (bulge chase) grid=%u x %d thr",
                       SB, grid1, S1_THREADS, smem1, grid2, S2_WARPS * 32);
    return ((stages & 1) ? 1 : 0) + ((stages & 2) ? 1 : 0);
}

}  // namespace twostage
}  // namespace sktb
