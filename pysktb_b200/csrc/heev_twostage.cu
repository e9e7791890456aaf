// Host side of the two-stage tridiagonalisation (heev_twostage.cuh): its own translation unit so that the DMMA kernels
// compile separately from the rest of libsktb.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define SKTB_TWOSTAGE_KERNELS
#include "heev_twostage.cuh"

namespace sktb {
namespace twostage {

static int ring_depth_for(int N) { return stage1_smem_bytes(N, 4) <= 227 * 1024 ? 4 : (stage1_smem_bytes(N, 3) <= 227 * 1024 ? 3 : 2); }
bool supported(int N) { return N >= 2 * SB && N - SB <= 512 && stage1_smem_bytes(N, 2) <= 227 * 1024; }   // panel QR: 8 x 64 rows in registers

static int launch_tridiag_impl(cplx* H, int N, long long nm, void* scratch, double* de, cudaStream_t st, char* info, size_t info_len, int stages,
                               cplx* Tstore, cplx* Rstore, long long rstride) {
    if (!supported(N) || nm <= 0) return -1;
    const int NR = (N + 7) / 8 * 8;
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    Stage1Args a1{};
    a1.H = H; a1.N = N; a1.nm = nm; a1.NR = NR;
    a1.mp = stage1_mp(N);
    a1.Tstore = Tstore; a1.npanels = stage1_panels(N);
    a1.AB = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * N * ABW * sizeof(cplx);
    a1.Yp = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * YP_MAX * NR * SB * sizeof(cplx);
    a1.Wg = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * NR * SB * sizeof(cplx);
    a1.Gg = reinterpret_cast<cplx*>(base);                 base += (size_t)nm * S1_WARPS * SB * SB * sizeof(cplx);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int rd = ring_depth_for(N);
    const size_t smem1 = stage1_smem_bytes(N, rd);
    void (*k1)(Stage1Args) = rd == 4 ? stage1_kernel<4> : (rd == 3 ? stage1_kernel<3> : stage1_kernel<2>);
    if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess) return -1;
    const unsigned grid1 = (unsigned)std::min<long long>(nm, sms);
    if (stages & 1) k1<<<grid1, S1_THREADS, smem1, st>>>(a1);
    Stage2Args a2{};
    a2.AB = a1.AB; a2.N = N; a2.nm = nm; a2.de = de;
    a2.Rstore = Rstore; a2.rstride = rstride;
    const size_t smem2 = stage2_smem_bytes(N);
    if (smem2 > 48 * 1024 && cudaFuncSetAttribute(stage2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) return -1;
    int occ2 = 2;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, stage2_kernel, S2_WARPS * 32, smem2);
    const unsigned grid2 = (unsigned)std::min<long long>(nm, (long long)std::max(occ2, 1) * sms);
    if (stages & 2) stage2_kernel<<<grid2, S2_WARPS * 32, smem2, st>>>(a2);
    if (info) snprintf(info, info_len, "two-stage: stage1_kernel (band %d, DMMA, ring %d) grid=%u x %d thr smem=%zuB + stage2_kernel (bulge chase) grid=%u x %d thr",
                       SB, rd, grid1, S1_THREADS, smem1, grid2, S2_WARPS * 32);
    return ((stages & 1) ? 1 : 0) + ((stages & 2) ? 1 : 0);
}
int launch_tridiag(cplx* H, int N, long long nm, void* scratch, double* de, cudaStream_t st, char* info, size_t info_len, int stages) {
    return launch_tridiag_impl(H, N, nm, scratch, de, st, info, info_len, stages, nullptr, nullptr, 0);
}
int launch_tridiag_keep(cplx* H, int N, long long nm, void* scratch, double* de, cplx* Tstore, cplx* Rstore, long long rstride, cudaStream_t st,
                        char* info, size_t info_len) {
    return launch_tridiag_impl(H, N, nm, scratch, de, st, info, info_len, 3, Tstore, Rstore, rstride);
}
int launch_rows(const RowsArgs& a, cudaStream_t st) {
    if (a.nm <= 0 || a.n_rows <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)ROWS_WARPS * a.N * sizeof(cplx);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rows_kernel, ROWS_WARPS * 32, smem);
    const long long groups = (a.nm * a.n_rows + ROWS_WARPS - 1) / ROWS_WARPS;
    rows_kernel<<<(unsigned)std::max<long long>(1, std::min<long long>(groups, (long long)std::max(occ, 1) * sms)), ROWS_WARPS * 32, smem, st>>>(a);
    return 1;
}

int launch_bisect(const double* de, int N, long long nm, double* evals, long long ld, long long col0, cudaStream_t st) {
    if (N > 1024 || nm <= 0) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int T = (N + 31) / 32 * 32;
    const size_t smem = ((size_t)2 * N + 96) * sizeof(double);
    bisect_prod_kernel<<<(unsigned)std::min<long long>(nm, (long long)sms * 8), T, smem, st>>>(de, N, nm, evals, ld, col0);
    return 1;
}

}  // namespace twostage
}  // namespace sktb

#ifdef SKTB_TS_PROFILE
// profile builds only (make TSPROF=1): phase clocks of stage1_kernel, read and reset
extern "C" int sktb_ts_profile(unsigned long long* out32) {
    cudaDeviceSynchronize();
    unsigned long long z[32] = {0};
    cudaMemcpyFromSymbol(out32, sktb::twostage::g_s1_prof, sizeof(z));
    cudaMemcpyToSymbol(sktb::twostage::g_s1_prof, z, sizeof(z));
    return 0;
}
#endif
