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
