// Instantiation + launch of the CTA-resident 33..64 solver (own translation unit: compiles in parallel).
#include "heev_small.cuh"

namespace sktb {
namespace small {

size_t cta64_bytes_per_k() { return (size_t)128 * 8 + (size_t)1024 * 16 + (size_t)256 * 16; }      // (d,e)[2][64] | 32x32 block | 16x16 block

// Eigenvalues for 32 < N <= 64: tridiag_cta64 (steps 0..31, four warps per matrix) -> trailing 32 x 32 block ->
// pair_wide32<LOAD> (steps 32..47, one warp per matrix) -> pair_tail16 (steps 48..62, two matrices per warp) -> tql<64>.
// SKTB_CTA64_SPLIT=0: all 63 steps in tridiag_cta64.
// fill the BUILD fields of the kernel arguments and opt the kernel into its dynamic shared memory
static bool c64_apply_build(C64Args& a, const C64Build* b, void (*kern)(C64Args), std::string* info) {
    a.H = nullptr; a.pv = b->pv; a.k = b->k;
    a.mesh0 = b->mesh ? b->mesh[0] : 1; a.mesh1 = b->mesh ? b->mesh[1] : 1; a.mesh2 = b->mesh ? b->mesh[2] : 1;
    a.first = b->first; a.soc = b->soc; a.Sdense = b->Sdense;
    // static (14 KB) + dynamic shared memory beyond 48 KB needs the opt-in: always ask for it
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes()) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(tridiag_cta64<build>, " + std::to_string(b->smem_bytes()) + ") failed"; return false;
    }
    return true;
}

int launch_cta64(const cplx* H, int N, long long nm, double* scratch, double* evals, long long ld, long long col0, int* fail, cudaStream_t st,
                 std::string* info, int h_ld, long long h_mstride, long long h_off, bool tridiag_only, const C64Build* build) {
    constexpr int NC = 64;
    if (N <= 32 || N > 64) { *info = "launch_cta64: N out of range"; return -1; }
    static const bool split = !(getenv("SKTB_CTA64_SPLIT") && getenv("SKTB_CTA64_SPLIT")[0] == '0');
    void (*kB)(TqlArgs) = tql_mode_env() == 2 ? tql_small_kernel<NC, 64, 2> : tql_small_kernel<NC, 64, 1>;
    const size_t smemB = (size_t)SMALL_WARPS * 2 * NC * DSTRIDE * 8;
    if (cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(smem=" + std::to_string(smemB) + ") failed"; return -1;
    }
    int dev = 0, sms = 148, occA = 0, occB = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, tridiag_cta64_kernel<false, false>, C64_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, kB, SMALL_WARPS * 32, smemB);
    if (occA < 1 || occB < 1) { *info = "cta64 kernels do not fit"; return -1; }
    double* de = scratch;
    cplx* blk32 = reinterpret_cast<cplx*>(scratch + (size_t)nm * 128);
    cplx* blk16 = blk32 + (size_t)nm * 1024;
    static const bool front = !(getenv("SKTB_CTA64_FRONT") && getenv("SKTB_CTA64_FRONT")[0] == '0');
    const int off = front ? 64 - N : 0;                     // front padding: steps over the padding are skipped
    C64Args a{H, N, nm, de, nullptr, split ? blk32 : nullptr, h_ld, h_mstride, h_off, front ? 1 : 0};
    const unsigned gridA = (unsigned)std::max<long long>(1, std::min<long long>(nm, (long long)sms * occA));
    int launches = 2;
    void (*kA)(C64Args) = split ? (build ? tridiag_cta64_kernel<true, true> : tridiag_cta64_kernel<true, false>)
                                : (build ? tridiag_cta64_kernel<false, true> : tridiag_cta64_kernel<false, false>);
    const size_t smemA = build ? build->smem_bytes() : 0;
    if (build && !c64_apply_build(a, build, kA, info)) return -1;
    if (split) {
        kA<<<gridA, C64_THREADS, smemA, st>>>(a);
        SmallArgs w{};
        w.k = nullptr; w.mesh0 = w.mesh1 = w.mesh2 = 1; w.first = 0; w.nk = nm; w.N = front ? 32 : N - 32; w.soc = 0;
        w.scratch = de; w.flags = nullptr; w.flag_off = 0; w.Hsrc = blk32; w.Hld = 32; w.out_stride = 128; w.d_off = 32 - off; w.e_off = 96 - off;
        const size_t smemW = (size_t)PAIR_WARPS * pair_buf_cplx<32>() * 16, smemT = (size_t)PAIR_WARPS * 2 * pair_buf_cplx<16>() * 16;
        const unsigned gridW = (unsigned)std::max<long long>(1, std::min<long long>((nm + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 2));
        pair_wide32_kernel<PAIR_SRC_LOAD, 4, false, false><<<gridW, PAIR_WARPS * 32, smemW, st>>>(w, blk16, nullptr);
        const unsigned gridT = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 1) / 2 + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 8));
        pair_tail16_kernel<8><<<gridT, PAIR_WARPS * 32, smemT, st>>>(blk16, de, nm, 128, 32 - off, 96 - off);
        launches = 4;
    } else {
        kA<<<gridA, C64_THREADS, smemA, st>>>(a);
    }
    TqlArgs t{de, nm, N, evals, ld, col0, fail};
    const unsigned gridB = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * occB));
    if (tridiag_only) --launches;       // the caller merges (d, e) with its own leading steps
    else kB<<<gridB, SMALL_WARPS * 32, smemB, st>>>(t);
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[320];
    snprintf(buf, sizeof buf, "%stridiag_cta64%s%s grid=%u x %d thr %d CTA/SM%s + tql_small<64> grid=%u smem=%zuB %d CTA/SM; N=%d", build ? "" : "hk + ",
             split ? "<split>" : "", build ? "<K0+K1 fused>" : "", gridA, C64_THREADS, occA, split ? " + pair_wide32<load> + pair_tail16" : "", gridB, smemB, occB, N);
    *info = buf;
    return launches;
}

static const size_t smemBT = (size_t)VEC32_WARPS * PAIR_VSTORE_CPLX * 16;      // vecbt32: reflectors of one matrix per warp
static bool vecbt32_smem_ok() {
    static const bool ok = cudaFuncSetAttribute(vecbt32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBT) == cudaSuccess;
    return ok;
}

static int launch_vecql32_family(const Vec32Args& v, long long nm, double* zstore, int* rankstore, int sms, cudaStream_t st) {
    static const bool use_list = !(getenv("SKTB_VEC_LIST") && getenv("SKTB_VEC_LIST")[0] == '0');
    const size_t smemQ = (size_t)VECQL_WARPS * VEC32_SMEM_DOUBLES * 8;
    const unsigned gridQ = (unsigned)std::max<long long>(1, std::min<long long>((nm + VECQL_WARPS - 1) / VECQL_WARPS, (long long)sms * 5));
    if (!use_list) {
        vecql32_kernel<0><<<gridQ, VECQL_WARPS * 32, smemQ, st>>>(v, zstore, rankstore, nullptr);
        return 1;
    }
    double2* lists = reinterpret_cast<double2*>(rankstore + (size_t)nm * 32);
    double* dfinal = reinterpret_cast<double*>(lists + (size_t)nm * VECREC_CAP);
    int* flags = reinterpret_cast<int*>(dfinal + (size_t)nm * 32);
    const size_t smemR = (size_t)SMALL_WARPS * 2 * 32 * DSTRIDE * 8;
    cudaFuncSetAttribute(tqlrec_kernel<32, VECREC_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR);
    const unsigned gridR = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * 3));
    tqlrec_kernel<32, VECREC_CAP><<<gridR, SMALL_WARPS * 32, smemR, st>>>(TqlRecArgs{v.de, nm, v.N, v.fail}, dfinal, lists, flags);
    vecapply32_kernel<0><<<gridQ, VECQL_WARPS * 32, smemQ, st>>>(v, dfinal, lists, flags, zstore, rankstore);
    vecql32_kernel<0><<<gridQ, VECQL_WARPS * 32, smemQ, st>>>(v, zstore, rankstore, flags);
    return 3;
}

size_t vec64_bytes_per_k() { return (size_t)128 * 8 + (size_t)C64_VSTORE_CPLX * 16 + 4096 * 8 + 64 * 4 + (size_t)VECREC64_CAP * 16 + 64 * 8 + 16; }   // (d,e) | reflectors | Z | ranks | rotation list | d final | flag

int launch_vec64(const cplx* H, int N, long long nm, double* evals, cplx* evecs, long long ld, long long col0, int* fail, void* scratch,
                 cudaStream_t st, std::string* info, const C64Build* build) {
    if (N <= 32 || N > 64 || !evecs) { *info = "launch_vec64: bad arguments"; return -1; }
    double* de = reinterpret_cast<double*>(scratch);
    cplx* vstore = reinterpret_cast<cplx*>(de + (size_t)nm * 128);
    double* zstore = reinterpret_cast<double*>(vstore + (size_t)nm * C64_VSTORE_CPLX);
    int* rankstore = reinterpret_cast<int*>(zstore + (size_t)nm * 4096);
    const size_t smemQ = (size_t)VEC64_SMEM_DOUBLES * 8;
    int dev = 0, sms = 148, occA = 0, occQ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, tridiag_cta64_kernel<false, false>, C64_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occQ, vecql64_kernel<0>, 32, smemQ);
    if (occA < 1 || occQ < 1) { *info = "vec64 kernels do not fit"; return -1; }
    static const bool split = !(getenv("SKTB_CTA64_SPLIT") && getenv("SKTB_CTA64_SPLIT")[0] == '0');
    cplx* blk32 = reinterpret_cast<cplx*>(zstore);          // the trailing blocks are consumed before Z is written
    C64Args a{H, N, nm, de, vstore, split ? blk32 : nullptr, 0, 0, 0, 0};
    const unsigned gridA = (unsigned)std::max<long long>(1, std::min<long long>(nm, (long long)sms * occA));
    void (*kA)(C64Args) = split ? (build ? tridiag_cta64_kernel<true, true> : tridiag_cta64_kernel<true, false>)
                                : (build ? tridiag_cta64_kernel<false, true> : tridiag_cta64_kernel<false, false>);
    const size_t smemA = build ? build->smem_bytes() : 0;
    if (build && !c64_apply_build(a, build, kA, info)) return -1;
    if (split) {
        // steps 0..31 in the CTA kernel, the trailing 32 x 32 block (steps 32..62) in the one-warp-per-matrix kernel,
        // its reflectors into rows / columns 32.. of the same [64][64] store
        kA<<<gridA, C64_THREADS, smemA, st>>>(a);
        SmallArgs w{};
        w.k = nullptr; w.mesh0 = w.mesh1 = w.mesh2 = 1; w.first = 0; w.nk = nm; w.N = N - 32; w.soc = 0;
        w.scratch = de; w.flags = nullptr; w.flag_off = 0; w.Hsrc = blk32; w.Hld = 32; w.out_stride = 128; w.d_off = 32; w.e_off = 96;
        w.v_per_k = C64_VSTORE_CPLX; w.v_stride = 64; w.v_off = 32; w.v_tau = 4096;
        const size_t smemW = (size_t)PAIR_WARPS * pair_buf_cplx<32>() * 16;
        const unsigned gridW = (unsigned)std::max<long long>(1, std::min<long long>((nm + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 2));
        pair_wide32_kernel<PAIR_SRC_LOAD, 4, true, true><<<gridW, PAIR_WARPS * 32, smemW, st>>>(w, nullptr, vstore);
    } else {
        kA<<<gridA, C64_THREADS, smemA, st>>>(a);
    }
    Vec64Args v{de, vstore, nm, N, evals, evecs, ld, col0, fail};
    const unsigned gridQ = (unsigned)std::max<long long>(1, std::min<long long>(nm, (long long)sms * occQ));
    static const bool use_list = !(getenv("SKTB_VEC_LIST") && getenv("SKTB_VEC_LIST")[0] == '0');
    int nql = 1;
    if (!use_list) vecql64_kernel<0><<<gridQ, 32, smemQ, st>>>(v, zstore, rankstore, nullptr);
    else {
        double2* lists = reinterpret_cast<double2*>(rankstore + (size_t)nm * 64);
        double* dfinal = reinterpret_cast<double*>(lists + (size_t)nm * VECREC64_CAP);
        int* flags = reinterpret_cast<int*>(dfinal + (size_t)nm * 64);
        const size_t smemR = (size_t)SMALL_WARPS * 2 * 64 * DSTRIDE * 8, smemP = (size_t)VECAPPLY64_SMEM_DOUBLES * 8;
        if (cudaFuncSetAttribute(tqlrec_kernel<64, VECREC64_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR) != cudaSuccess) {
            *info = "cudaFuncSetAttribute(tqlrec64) failed"; return -1;
        }
        const unsigned gridR = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms));
        tqlrec_kernel<64, VECREC64_CAP><<<gridR, SMALL_WARPS * 32, smemR, st>>>(TqlRecArgs{de, nm, N, fail}, dfinal, lists, flags);
        vecapply64_kernel<0><<<gridQ, 32, smemP, st>>>(v, dfinal, lists, flags, zstore, rankstore);
        vecql64_kernel<0><<<gridQ, 32, smemQ, st>>>(v, zstore, rankstore, flags);
        nql = 3;
    }
    const unsigned gridV = (unsigned)std::max<long long>(1, std::min<long long>(nm, (long long)sms * 2));
    const size_t smemV64 = (size_t)C64_VSTORE_CPLX * 16;
    if (cudaFuncSetAttribute(vecbt64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemV64) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(vecbt64) failed"; return -1;
    }
    vecbt64_kernel<0><<<gridV, 128, smemV64, st>>>(v, zstore, rankstore);
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[256];
    snprintf(buf, sizeof buf, "tridiag_cta64<vec%s%s> grid=%u x %d thr %d CTA/SM + vecql64 grid=%u smem=%zuB %d CTA/SM + vecbt64 grid=%u; N=%d",
             split ? ",split + pair_wide32<load,full,vec>" : "", build ? ",K0+K1 fused" : "", gridA, C64_THREADS, occA, gridQ, smemQ, occQ, gridV, N);
    *info = buf;
    return 2 + nql + (split ? 1 : 0);
}

int launch_heev32(const cplx* H, int N, long long nm, double* evals, cplx* evecs, long long ld, long long col0, int* fail, void* scratch,
                  cudaStream_t st, std::string* info) {
    if (N < 1 || N > 32) { *info = "launch_heev32: N out of range"; return -1; }
    SmallArgs a{};
    a.k = nullptr; a.mesh0 = a.mesh1 = a.mesh2 = 1; a.first = 0; a.nk = nm; a.N = N; a.soc = 0;
    a.scratch = reinterpret_cast<double*>(scratch); a.flags = nullptr; a.flag_off = 0; a.Hsrc = H;
    cplx* vstore = reinterpret_cast<cplx*>(reinterpret_cast<unsigned char*>(scratch) + (size_t)nm * 64 * 8);
    double* zstore = reinterpret_cast<double*>(vstore + (size_t)nm * PAIR_VSTORE_CPLX);
    int* rankstore = reinterpret_cast<int*>(zstore + (size_t)nm * 1024);
    const size_t smemW = (size_t)PAIR_WARPS * pair_buf_cplx<32>() * 16;
    const size_t smemB = (size_t)SMALL_WARPS * 2 * 32 * DSTRIDE * 8;
    // steps 0..15 in the one-warp kernel, the trailing 16 x 16 block (parked in the Z region) two matrices per warp
    static const bool split = !(getenv("SKTB_VEC32_SPLIT") && getenv("SKTB_VEC32_SPLIT")[0] == '0');
    void (*kW)(SmallArgs, cplx*, cplx*) = split ? (evecs ? pair_wide32_kernel<PAIR_SRC_LOAD, 4, false, true> : pair_wide32_kernel<PAIR_SRC_LOAD, 4, false, false>)
                                                : (evecs ? pair_wide32_kernel<PAIR_SRC_LOAD, 4, true, true> : pair_wide32_kernel<PAIR_SRC_LOAD, 4, true, false>);
    void (*kB)(TqlArgs) = tql_mode_env() == 2 ? tql_small_kernel<32, 32, 2> : tql_small_kernel<32, 32, 1>;
    if (cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB) != cudaSuccess) { *info = "cudaFuncSetAttribute failed"; return -1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned gridW = (unsigned)std::max<long long>(1, std::min<long long>((nm + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 2));
    cplx* blk16 = reinterpret_cast<cplx*>(zstore);
    kW<<<gridW, PAIR_WARPS * 32, smemW, st>>>(a, split ? blk16 : nullptr, evecs ? vstore : nullptr);
    int launches = 1;
    if (split) {
        const size_t smemT = (size_t)PAIR_WARPS * 2 * pair_buf_cplx<16>() * 16;
        const unsigned gridT = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 1) / 2 + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 8));
        if (evecs) pair_tail16_kernel<4, true><<<gridT, PAIR_WARPS * 32, smemT, st>>>(blk16, a.scratch, nm, 64, 0, 32, vstore);
        else pair_tail16_kernel<4, false><<<gridT, PAIR_WARPS * 32, smemT, st>>>(blk16, a.scratch, nm, 64, 0, 32, nullptr);
        ++launches;
    }
    if (evecs) {
        Vec32Args v{a.scratch, vstore, nm, N, evals, evecs, ld, col0, fail};
        if (!vecbt32_smem_ok()) { *info = "cudaFuncSetAttribute(vecbt32) failed"; return -1; }
        launches += launch_vecql32_family(v, nm, zstore, rankstore, sms, st);
        const unsigned gridV = (unsigned)std::max<long long>(1, std::min<long long>((nm + VEC32_WARPS - 1) / VEC32_WARPS, (long long)sms * 2));
        vecbt32_kernel<0><<<gridV, VEC32_WARPS * 32, smemBT, st>>>(v, zstore, rankstore);
        launches += 1;
    } else {
        TqlArgs t{a.scratch, nm, N, evals, ld, col0, fail};
        const unsigned gridB = (unsigned)std::max<long long>(1, std::min<long long>(((nm + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * 3));
        kB<<<gridB, SMALL_WARPS * 32, smemB, st>>>(t);
        launches += 1;
    }
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[200];
    snprintf(buf, sizeof buf, "pair_wide32<load,%s%s> grid=%u x %d thr + %s; N=%d", split ? "split + pair_tail16" : "full", evecs ? ",vec" : "", gridW, PAIR_WARPS * 32,
             evecs ? "vecql32 + vecbt32" : "tql_small", N);
    *info = buf;
    return launches;
}

size_t vec32_bytes_per_k() { return (size_t)64 * 8 + (size_t)PAIR_VSTORE_CPLX * 16 + 1024 * 8 + 32 * 4 + (size_t)VECREC_CAP * 16 + 32 * 8 + 16; }   // (d,e) | reflectors | Z | ranks | rotation list | d final | flag

int launch_vec32(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N, double* evals,
                 cplx* evecs, long long ld, long long col0, int* flags, long long flag_off, int* fail, void* scratch, cudaStream_t st,
                 std::string* info) {
    if (N < 1 || N > 32) { *info = "launch_vec32: N out of range"; return -1; }
    const int n = tab.n_orb, nb = tab.n_bonds;
    SmallArgs a;
    a.t = tab; a.k = k_d;
    a.mesh0 = nk_mesh ? nk_mesh[0] : 1; a.mesh1 = nk_mesh ? nk_mesh[1] : 1; a.mesh2 = nk_mesh ? nk_mesh[2] : 1;
    a.first = first; a.nk = nk; a.N = N; a.soc = soc;
    a.scratch = reinterpret_cast<double*>(scratch); a.flags = flags; a.flag_off = flag_off;
    cplx* vstore = reinterpret_cast<cplx*>(reinterpret_cast<unsigned char*>(scratch) + (size_t)nk * 64 * 8);
    const size_t smemW = table_doubles(n, nb, N, soc != 0, false) * 8 + (size_t)PAIR_WARPS * (pair_buf_cplx<32>() + nb) * 16;
    const size_t smemQ = (size_t)VECQL_WARPS * VEC32_SMEM_DOUBLES * 8;
    double* zstore = reinterpret_cast<double*>(vstore + (size_t)nk * PAIR_VSTORE_CPLX);
    int* rankstore = reinterpret_cast<int*>(zstore + (size_t)nk * 1024);
    // split like the eigenvalue path: steps 0..15 in the one-warp kernel, the trailing 16 x 16 block (parked in the Z
    // region, which is written later) in the two-matrices-per-warp kernel; both keep their reflectors
    static const bool split = !(getenv("SKTB_VEC32_SPLIT") && getenv("SKTB_VEC32_SPLIT")[0] == '0');
    void (*kW)(SmallArgs, cplx*, cplx*) = split ? (soc ? pair_wide32_kernel<PAIR_SRC_SOC, 4, false, true> : pair_wide32_kernel<PAIR_SRC_NOSOC, 4, false, true>)
                                                : (soc ? pair_wide32_kernel<PAIR_SRC_SOC, 4, true, true> : pair_wide32_kernel<PAIR_SRC_NOSOC, 4, true, true>);
    if (smemW > 48 * 1024 && cudaFuncSetAttribute(kW, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemW) != cudaSuccess) {
        *info = "cudaFuncSetAttribute(smem=" + std::to_string(smemW) + ") failed"; return -1;
    }
    int dev = 0, sms = 148, occW = 0, occQ = 0, occV = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occW, kW, PAIR_WARPS * 32, smemW);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occQ, vecql32_kernel<0>, VECQL_WARPS * 32, smemQ);
    if (!vecbt32_smem_ok()) { *info = "cudaFuncSetAttribute(vecbt32) failed"; return -1; }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occV, vecbt32_kernel<0>, VEC32_WARPS * 32, smemBT);
    if (occW < 1 || occQ < 1 || occV < 1) { *info = "vec32 kernels do not fit (smem " + std::to_string(smemW) + ")"; return -1; }
    const unsigned gridW = (unsigned)std::max<long long>(1, std::min<long long>((nk + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * occW));
    cplx* blk16 = reinterpret_cast<cplx*>(zstore);
    kW<<<gridW, PAIR_WARPS * 32, smemW, st>>>(a, split ? blk16 : nullptr, vstore);
    if (split) {
        const size_t smemT = (size_t)PAIR_WARPS * 2 * pair_buf_cplx<16>() * 16;
        const unsigned gridT = (unsigned)std::max<long long>(1, std::min<long long>(((nk + 1) / 2 + PAIR_WARPS - 1) / PAIR_WARPS, (long long)sms * 8));
        pair_tail16_kernel<4, true><<<gridT, PAIR_WARPS * 32, smemT, st>>>(blk16, a.scratch, nk, 64, 0, 32, vstore);
    }
    Vec32Args v{a.scratch, vstore, nk, N, evals, evecs, ld, col0, fail};
    const unsigned gridQ = (unsigned)std::max<long long>(1, std::min<long long>((nk + VECQL_WARPS - 1) / VECQL_WARPS, (long long)sms * occQ));
    const int nq = launch_vecql32_family(v, nk, zstore, rankstore, sms, st);
    const unsigned gridV = (unsigned)std::max<long long>(1, std::min<long long>((nk + VEC32_WARPS - 1) / VEC32_WARPS, (long long)sms * occV));
    vecbt32_kernel<0><<<gridV, VEC32_WARPS * 32, smemBT, st>>>(v, zstore, rankstore);
    if (cudaGetLastError() != cudaSuccess) { *info = "launch failed"; return -1; }
    char buf[320];
    snprintf(buf, sizeof buf, "pair_wide32<soc=%d,%s,vec> grid=%u x %d thr smem=%zuB %d CTA/SM + vecql32 grid=%u smem=%zuB %d CTA/SM + vecbt32 grid=%u %d CTA/SM; N=%d",
             soc, split ? "split + pair_tail16<vec>" : "full", gridW, PAIR_WARPS * 32, smemW, occW, gridQ, smemQ, occQ, gridV, occV, N);
    *info = buf;
    return 2 + nq + (split ? 1 : 0);
}

}  // namespace small
}  // namespace sktb
