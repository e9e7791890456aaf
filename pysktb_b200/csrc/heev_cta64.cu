// Instantiation + launch of the CTA-resident 33..64 solver (own translation unit: compiles in parallel).
#include "heev_small.cuh"

namespace sktb {
namespace small {

int launch_cta64(const cplx* H, int N, long long nm, double* scratch, double* evals, long long ld, long long col0, int* fail, cudaStream_t st,
                 std::string* info) {
    constexpr int NC = 64;
    if (N <= 32 || N > 64) { *info = "launch_cta64: N out of range"; return -1; }
    auto kB = tql_small_kernel<NC, 64>;
    const size_t smemB = (size_t)SMALL_WARPS * 2 * NC * DSTRIDE * 8;
    if (cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(smem=" + std::to_string(smemB) + ") failed"; return -1;
    }
    int dev = 0, sms = 148, occA = 0, occB = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, tridiag_cta64_kernel<0>, C64_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, kB, SMALL_WARPS * 32, smemB);
    if (occA < 1 || occB < 1) { *info = "cta64 kernels do not fit"; return -1; }
    C64Args a{H, N, nm, scratch};
    const unsigned gridA = (unsigned)std::max<long long>(1, std::min<long long>(nm, (long long)sms * occA));
    tridiag_cta64_kernel<0><<<gridA, C64_THREADS, 0, st>>>(a);
    TqlArgs t{scratch, nm, N, evals, ld, col0, fail};
    const unsigned gridB = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * occB));
    kB<<<gridB, SMALL_WARPS * 32, smemB, st>>>(t);
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[256];
    snprintf(buf, sizeof buf, "hk + tridiag_cta64 grid=%u x %d thr %d CTA/SM + tql_small<64> grid=%u smem=%zuB %d CTA/SM; N=%d", gridA, C64_THREADS, occA,
             gridB, smemB, occB, N);
    *info = buf;
    return 2;
}

}  // namespace small
}  // namespace sktb
