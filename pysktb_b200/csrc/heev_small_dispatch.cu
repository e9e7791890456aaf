// Instantiations + dispatch of the register-resident solver (kept in their own translation units so the
// fully unrolled kernels compile in parallel).  SKTB_SMALL_NC selects which column count this TU provides.
#include "heev_small.cuh"

namespace sktb {
namespace small {

#ifdef SKTB_SMALL_NC
#define SKTB_CAT2(a, b) a##b
#define SKTB_CAT(a, b) SKTB_CAT2(a, b)
int SKTB_CAT(launch_nc_, SKTB_SMALL_NC)(const LaunchCtx& c, bool gate) {
    constexpr int NC = SKTB_SMALL_NC;
    constexpr int L = NC <= 4 ? 4 : (NC <= 8 ? 8 : (NC <= 16 ? 16 : 32));
    return launch_nc<NC, L>(c, gate);
}
#else
int launch_nc_4(const LaunchCtx&, bool);
int launch_nc_8(const LaunchCtx&, bool);
int launch_nc_12(const LaunchCtx&, bool);
int launch_nc_16(const LaunchCtx&, bool);
int launch_nc_24(const LaunchCtx&, bool);
int launch_nc_32(const LaunchCtx&, bool);

// ---------------------------------------------------------------------------------------------------------
// N = 2 (graphene pz, config C2; also one s orbital with spin): closed form, one k-point per thread.
//   A = [[a, conj(b)], [b, d]] read from the lower triangle (LAPACK UPLO='L'):  lambda = m -+ sqrt(delta^2 + |b|^2),
//   m = (a + d)/2, delta = (a - d)/2.  HBM-write bound by construction (16 B out per k); in practice bound by the
//   one sincos per bond.  Bond phases use the same operation order as the register kernels (calc_g arithmetic).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) solve_n2_kernel(SmallArgs a, double* __restrict__ evals, long long ld, long long col0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.t.n_orb, nb = a.t.n_bonds;
    double* Lt = reinterpret_cast<double*>(smem_raw);          // [nb][n][n]
    double* Bv = Lt + (size_t)nb * n * n;                       // [nb][3]
    for (int t = threadIdx.x; t < nb * n * n; t += blockDim.x) Lt[t] = a.t.Lt[t];
    for (int t = threadIdx.x; t < 3 * nb; t += blockDim.x) Bv[t] = a.t.bvec[t];
    __syncthreads();
    const bool soc = a.soc != 0;
    // soc: n = 1, H = h I2 + S;  no soc: n = 2, H = h
    double on0 = a.t.onsite[0], on1 = soc ? on0 : a.t.onsite[1];
    cplx s00 = cmake(0.0, 0.0), s11 = s00, s10 = s00;
    if (soc) { s00 = a.t.St[0]; s10 = a.t.St[1]; s11 = a.t.St[3]; }     // St[col * 2 + row], mirrored lower
    for (long long kq = (long long)blockIdx.x * blockDim.x + threadIdx.x; kq < a.nk; kq += (long long)gridDim.x * blockDim.x) {
        double k0, k1, k2;
        if (a.k) { k0 = a.k[3 * kq]; k1 = a.k[3 * kq + 1]; k2 = a.k[3 * kq + 2]; }
        else kmesh_point(a.first + kq, a.mesh0, a.mesh1, a.mesh2, k0, k1, k2);
        const double c0 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[0]), __dmul_rn(k1, a.t.rec[3])), __dmul_rn(k2, a.t.rec[6]));
        const double c1 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[1]), __dmul_rn(k1, a.t.rec[4])), __dmul_rn(k2, a.t.rec[7]));
        const double c2 = __dadd_rn(__dadd_rn(__dmul_rn(k0, a.t.rec[2]), __dmul_rn(k1, a.t.rec[5])), __dmul_rn(k2, a.t.rec[8]));
        double a00 = 0.0, a11 = 0.0, bx = 0.0, by = 0.0;
        for (int b = 0; b < nb; ++b) {
            const double dot = __dadd_rn(__dadd_rn(__dmul_rn(c0, Bv[3 * b]), __dmul_rn(c1, Bv[3 * b + 1])), __dmul_rn(c2, Bv[3 * b + 2]));
            double sn, cs;
            sincos(__dmul_rn(6.283185307179586, dot), &sn, &cs);
            const double* T = Lt + (size_t)b * n * n;
            if (soc) { a00 = fma(T[0], cs, a00); }
            else {
                a00 = fma(T[0], cs, a00); a11 = fma(T[3], cs, a11);
                bx = fma(T[1], cs, bx); by = fma(T[1], sn, by);          // Lt[kappa=0][iota=1] = T_b[1][0]
            }
        }
        double A00, A11;
        if (soc) { A00 = a00 + on0 + s00.x; A11 = a00 + on0 + s11.x; bx = s10.x; by = s10.y; }
        else { A00 = a00 + on0; A11 = a11 + on1; }
        const double m = 0.5 * (A00 + A11), dl = 0.5 * (A00 - A11);
        const double r = sqrt(fma(dl, dl, fma(bx, bx, by * by)));
        evals[col0 + kq] = m - r;
        evals[ld + col0 + kq] = m + r;
        if (a.flags) a.flags[a.flag_off + kq] = 0;
    }
}

static int launch_n2(const LaunchCtx& c) {
    const int n = c.a.t.n_orb, nb = c.a.t.n_bonds;
    const size_t smem = ((size_t)nb * n * n + 3 * (size_t)nb) * 8;
    if (smem > 48 * 1024) return -2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((c.a.nk + 255) / 256, (long long)sms * 8));
    solve_n2_kernel<<<grid, 256, smem, c.st>>>(c.a, c.evals, c.ld, c.col0);
    if (cudaGetLastError() != cudaSuccess) { *c.info = "launch failed"; return -1; }
    char buf[160];
    snprintf(buf, sizeof buf, "solve_n2 (closed form) grid=%u x 256 thr smem=%zuB; N=2", grid, smem);
    *c.info = buf;
    return 1;
}

// N <= 8 without the gate: one matrix per thread (heev_thread.cuh).  SKTB_NO_THREAD=1: lane-group kernels (A/B, parity tests).
template <int NC, bool SOC>
static int launch_thread_t(const LaunchCtx& c) {
    // SKTB_THREAD_SPLIT=0: one fused kernel (tridiagonalisation + QL + sort in the same thread)
    static const bool split = !(getenv("SKTB_THREAD_SPLIT") && getenv("SKTB_THREAD_SPLIT")[0] == '0');
    constexpr int P2 = NC <= 4 ? 4 : (NC <= 8 ? 8 : 16);
    const size_t smem = thread_smem_bytes(NC, SOC, c.a.t.n_bonds, split);
    void (*kern)(SmallArgs, double*, long long, long long, int*) = split ? solve_thread_kernel<NC, SOC, true> : solve_thread_kernel<NC, SOC, false>;
    void (*kB)(TqlArgs) = tql_small_kernel<NC, P2, 2>;
    const size_t smemB = (size_t)SMALL_WARPS * 2 * NC * DSTRIDE * 8;
    if (smem > 160 * 1024) return -2;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    int dev = 0, sms = 148, occ = 0, occB = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREAD_CTA, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, kB, SMALL_WARPS * 32, smemB);
    if (occ < 1 || occB < 1) return -2;
    const long long per_k = 2LL * NC * 8;
    const long long chunk = split ? std::max<long long>(32, std::min<long long>(c.a.nk, (long long)(c.scratch_bytes / per_k) / 32 * 32)) : c.a.nk;
    int launches = 0;
    unsigned grid = 0, gridB = 0;
    for (long long c0 = 0; c0 < c.a.nk; c0 += chunk) {
        const long long cnt = std::min(chunk, c.a.nk - c0);
        SmallArgs a = c.a;
        a.k = c.a.k ? c.a.k + 3 * c0 : nullptr;
        a.first = c.a.first + c0; a.nk = cnt; a.scratch = c.scratch;
        a.flag_off = c.a.flag_off + c0;
        grid = (unsigned)std::max<long long>(1, std::min<long long>((cnt + THREAD_CTA - 1) / THREAD_CTA, (long long)sms * occ));
        kern<<<grid, THREAD_CTA, smem, c.st>>>(a, c.evals, c.ld, c.col0 + c0, c.fail);
        ++launches;
        if (split) {
            TqlArgs t{c.scratch, cnt, c.a.N, c.evals, c.ld, c.col0 + c0, c.fail};
            gridB = (unsigned)std::max<long long>(1, std::min<long long>(((cnt + 31) / 32 + SMALL_WARPS - 1) / SMALL_WARPS, (long long)sms * occB));
            kB<<<gridB, SMALL_WARPS * 32, smemB, c.st>>>(t);
            ++launches;
        }
        if (cudaGetLastError() != cudaSuccess) { *c.info = "launch failed"; return -1; }
    }
    char buf[260];
    snprintf(buf, sizeof buf, "solve_thread<NC=%d,soc=%d%s> (one matrix per thread, K0+K1+K2a fused) grid=%u x %d thr smem=%zuB %d CTA/SM%s; N=%d", NC, (int)SOC,
             split ? ",split" : "", grid, THREAD_CTA, smem, occ, split ? " + tql_small (per-lane QL)" : "", c.a.N);
    *c.info = buf;
    return launches;
}
static int launch_thread(const LaunchCtx& c) {
    const int N = c.a.N;
    if (c.a.soc) return N <= 4 ? launch_thread_t<4, true>(c) : (N <= 8 ? launch_thread_t<8, true>(c) : launch_thread_t<12, true>(c));
    return N <= 4 ? launch_thread_t<4, false>(c) : (N <= 8 ? launch_thread_t<8, false>(c) : launch_thread_t<12, false>(c));
}

int launch_solve(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N,
                 double* evals, long long ld, long long col0, int* flags, long long flag_off, int* fail, double* scratch, size_t scratch_bytes,
                 cudaStream_t st, std::string* info) {
    LaunchCtx c;
    SmallArgs& a = c.a;
    a.t = tab; a.k = k_d;
    a.mesh0 = nk_mesh ? nk_mesh[0] : 1; a.mesh1 = nk_mesh ? nk_mesh[1] : 1; a.mesh2 = nk_mesh ? nk_mesh[2] : 1;
    a.first = first; a.nk = nk; a.N = N; a.soc = soc;
    a.scratch = scratch; a.flags = flags; a.flag_off = flag_off;
    c.evals = evals; c.ld = ld; c.col0 = col0; c.fail = fail; c.scratch = scratch; c.scratch_bytes = scratch_bytes; c.st = st; c.info = info;
    if (!scratch || scratch_bytes < (size_t)32 * 2 * padded_nc(N) * 8) { *info = "scratch too small"; return -1; }
    const bool gate = flags != nullptr && tab.asym_bound > 0.5e-9;
    if (N == 2 && !gate) {
        static const bool no_n2 = getenv("SKTB_NO_N2") != nullptr && getenv("SKTB_NO_N2")[0] == '1';
        if (!no_n2) { int rc = launch_n2(c); if (rc != -2) return rc; }
    }
    static const int thread_max = getenv("SKTB_THREAD_MAX") ? atoi(getenv("SKTB_THREAD_MAX")) : 12;   // N = 9..12 spills ~1 KB per thread and still runs 3x faster than 16 lanes per matrix (C3: 2.6e8 -> 8.0e8 k/s)
    if (N <= thread_max && N <= 12 && !gate && tab.thread_order) {
        static const bool no_thread = getenv("SKTB_NO_THREAD") != nullptr && getenv("SKTB_NO_THREAD")[0] == '1';
        if (!no_thread) { int rc = launch_thread(c); if (rc != -2) return rc; }
    }
    switch (padded_nc(N)) {
        case 4: return launch_nc_4(c, gate);
        case 8: return launch_nc_8(c, gate);
        case 12: return launch_nc_12(c, gate);
        case 16: return launch_nc_16(c, gate);
        case 24: return launch_nc_24(c, gate);
        default: return launch_nc_32(c, gate);
    }
}
#endif

}  // namespace small
}  // namespace sktb
