// Instantiations + dispatch of the register-resident solver (kept in their own translation units so the
// fully unrolled kernels compile in parallel).  SKTB_SMALL_NC selects which column count this TU provides.
#include "heev_small.cuh"

namespace sktb {
namespace small {

#ifdef SKTB_SMALL_NC
#define SKTB_CAT2(a, b) a##b
#define SKTB_CAT(a, b) SKTB_CAT2(a, b)
int SKTB_CAT(launch_nc_, SKTB_SMALL_NC)(const SmallArgs& a, bool gate, cudaStream_t st, std::string* info) {
    constexpr int NC = SKTB_SMALL_NC;
    constexpr int L = NC <= 4 ? 4 : (NC <= 8 ? 8 : (NC <= 16 ? 16 : 32));
    return launch_nc<NC, L>(a, gate, st, info);
}
#else
int launch_nc_4(const SmallArgs&, bool, cudaStream_t, std::string*);
int launch_nc_8(const SmallArgs&, bool, cudaStream_t, std::string*);
int launch_nc_12(const SmallArgs&, bool, cudaStream_t, std::string*);
int launch_nc_16(const SmallArgs&, bool, cudaStream_t, std::string*);
int launch_nc_24(const SmallArgs&, bool, cudaStream_t, std::string*);
int launch_nc_32(const SmallArgs&, bool, cudaStream_t, std::string*);

int launch_solve(const SmallTables& tab, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc, int N,
                 double* evals, long long ld, long long col0, int* flags, int* fail, cudaStream_t st, std::string* info) {
    SmallArgs a;
    a.t = tab; a.k = k_d;
    a.mesh0 = nk_mesh ? nk_mesh[0] : 1; a.mesh1 = nk_mesh ? nk_mesh[1] : 1; a.mesh2 = nk_mesh ? nk_mesh[2] : 1;
    a.first = first; a.nk = nk; a.N = N; a.soc = soc;
    a.evals = evals; a.ld = ld; a.col0 = col0; a.flags = flags; a.fail = fail;
    const bool gate = flags != nullptr && tab.asym_bound > 0.5e-9;
    if (N <= 4) return launch_nc_4(a, gate, st, info);
    if (N <= 8) return launch_nc_8(a, gate, st, info);
    if (N <= 12) return launch_nc_12(a, gate, st, info);
    if (N <= 16) return launch_nc_16(a, gate, st, info);
    if (N <= 24) return launch_nc_24(a, gate, st, info);
    return launch_nc_32(a, gate, st, info);
}
#endif

}  // namespace small
}  // namespace sktb
