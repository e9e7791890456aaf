// libsktb.so - C ABI (include/sktb.h) over the sm_100a kernels.  Host side: plan construction, workspace,
// chunking, stream pipelining.  No CPU compute path: every numerical entry point launches kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/sktb.h"
#include "common.cuh"
#include "sk_table.cuh"
#include "kernels_generic.cuh"
#include "gf_inverse.cuh"
#include "heev_twostage.cuh"
#include "kernels_big_vec.cuh"
#include "heev_small.cuh"

using namespace sktb;

// ------------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local std::string g_kinfo;
static std::atomic<long long> g_launches{0};

static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define LAUNCHED()                                                                                   \
    do {                                                                                             \
        ++g_launches;                                                                                \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// Every entry point works on the plan's device and restores the caller's current device on return (a process that
// drives several GPUs must not see its current device change under it; torch reads it through cudaGetDevice).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                                                               \
    DeviceGuard _dev_guard(dev);                                                                     \
    if (!_dev_guard.ok) return fail("cudaSetDevice(%d) failed: %s", (int)(dev), cudaGetErrorString(cudaGetLastError()))

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return 1; }
            want = bytes;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct sktb_plan {
    int device = 0;
    int n_atoms = 0, n_orb = 0, n_bonds = 0, n_pairs = 0, soc_nnz = 0;
    PlanView view{};
    std::vector<void*> owned;                 // device allocations of the tables
    std::vector<double> T_host_creation_order;  // hopping blocks in the caller's bond order
    std::vector<int> bond_ni, bond_nj;
    double asym_bound = 0.0;
    double herm_defect = 0.0, herm_defect_soc = 0.0;   // max |T_b[mu,nu] - T_rev(b)[nu,mu]| / max |S - S^dagger|: both 0 <=> H(k) = H(k)^dagger for every k
    const cplx* soc_dense = nullptr;          // [2n][2n] dense soc_mat (32 < 2n <= 64: fused build of tridiag_cta64)
    small::SmallTables small_tab{};           // dense tables for the register-resident path
    bool small_ok = false;
    struct WorkSet { Buf H, RT, K, S; };
    WorkSet ws[3];                            // [0],[1]: the two host-front-end streams; [2]: caller-stream API
    Buf wsE, wsVec, wsFlag, wsPart, wsMisc;
    int* d_any = nullptr;                     // device int: any Hermiticity-gate hit
    int* d_fail = nullptr;                    // device int: QL non-convergence count
    int* h_pinned = nullptr;                  // 4 pinned ints
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int last_launches = 0;
    bool timing_valid = false;
    // host front end
    cudaStream_t hs[2] = {nullptr, nullptr};
    Buf hK[2], hE[2], hV[2], hF[2];
};

template <class T>
static int upload(sktb_plan* p, const T* src, size_t n, const T** dst) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (cudaMalloc(&d, bytes) != cudaSuccess) return fail("cudaMalloc(%zu) failed", bytes);
    if (n && cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D copy failed");
    p->owned.push_back(d);
    *dst = reinterpret_cast<const T*>(d);
    return 0;
}

extern "C" const char* sktb_last_error(void) { return g_err.c_str(); }
extern "C" const char* sktb_version(void) { return "sktb-b200 0.1 (sm_100a)"; }
extern "C" const char* sktb_last_kernel_info(void) { return g_kinfo.c_str(); }
extern "C" long long sktb_launch_count(void) { return g_launches.load(); }
extern "C" const char* sktb_v_names(void) {
    return "V_sss,V_sps,V_pps,V_ppp,V_sds,V_pds,V_pdp,V_dds,V_ddp,V_ddd,V_sfs,V_pfs,V_pfp,V_dfs,V_dfp,V_dfd,"
           "V_ffs,V_ffp,V_ffd,V_fff,V_SSs,V_sSs,V_Sps,V_Sds,V_Sfs";
}

// ------------------------------------------------------------------------------------------------------
// K4 / K0 entry points
// ------------------------------------------------------------------------------------------------------
extern "C" int sktb_sk_table(const double* V_d, const double* lmn_d, double* out_d, int32_t nb, void* stream) {
    if (nb <= 0) return 0;
    sk_table_kernel<<<(nb + 63) / 64, 64, 0, (cudaStream_t)stream>>>(V_d, lmn_d, out_d, nb);
    LAUNCHED();
    return 0;
}

extern "C" int sktb_sk_table_grad(const double* V_d, const double* dV_d, const double* vec_d, double* out_d, int32_t nb, void* stream) {
    if (nb <= 0) return 0;
    if (!V_d || !dV_d || !vec_d || !out_d) return fail("null argument");
    SkDual* work = nullptr;
    CK(cudaMalloc((void**)&work, (size_t)nb * 289 * sizeof(SkDual)));
    sk_table_grad_kernel<<<(nb + 31) / 32, 32, 0, (cudaStream_t)stream>>>(V_d, dV_d, vec_d, out_d, work, nb);
    ++g_launches;
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(work);
    if (e != cudaSuccess) return fail("sk_table_grad kernel failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sktb_kmesh(const int32_t nk[3], int64_t first, int64_t count, double* k_out_d, void* stream) {
    if (count <= 0) return 0;
    if (nk[0] < 0 || nk[1] < 0 || nk[2] < 0) return fail("negative mesh size");
    long long blocks = std::min<long long>((count + 255) / 256, 148 * 32);
    kmesh_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(nk[0], nk[1], nk[2], first, count, k_out_d);
    LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------------
extern "C" int sktb_plan_create(const sktb_model_desc* D, int device, sktb_plan** out) {
    if (!D || !out) return fail("null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: libsktb has no CPU fallback (cudaGetDeviceCount: %s)", cudaGetErrorString(cudaGetLastError()));
    if (device < 0 || device >= ndev) return fail("device %d out of range (have %d)", device, ndev);
    ON_DEVICE(device);
    const int na = D->n_atoms, n = D->n_orb, nb = D->n_bonds;
    if (na <= 0 || n <= 0 || nb < 0) return fail("bad model sizes");
    for (int a = 0; a < na; ++a)
        if (D->atom_orb_start[a + 1] <= D->atom_orb_start[a]) return fail("atom %d has no orbitals", a);
    if (D->atom_orb_start[0] != 0 || D->atom_orb_start[na] != n) return fail("atom_orb_start inconsistent with n_orb");
    for (int i = 0; i < n; ++i)
        if (D->orb_kind[i] < 0 || D->orb_kind[i] >= SKTB_N_ORB_KINDS) return fail("orb_kind[%d] out of range", i);
    for (int b = 0; b < nb; ++b)
        if (D->bond_i[b] < 0 || D->bond_i[b] >= na || D->bond_j[b] < 0 || D->bond_j[b] >= na) return fail("bond %d: atom index out of range", b);
    for (int t = 0; t < D->soc_nnz; ++t)
        if (D->soc_row[t] < 0 || D->soc_row[t] >= 2 * n || D->soc_col[t] < 0 || D->soc_col[t] >= 2 * n) return fail("soc entry %d out of range", t);

    sktb_plan* p = new sktb_plan();
    p->device = device; p->n_atoms = na; p->n_orb = n; p->n_bonds = nb; p->soc_nnz = D->soc_nnz;
    auto bail = [&](int rc) { sktb_plan_destroy(p); return rc; };

    // ---- K4 on the device: full 17x17 tables per bond -----------------------------------------------
    std::vector<double> tab((size_t)nb * 289);
    if (nb) {
        double *dV = nullptr, *dL = nullptr, *dT = nullptr;
        if (cudaMalloc(&dV, (size_t)nb * 25 * 8) != cudaSuccess || cudaMalloc(&dL, (size_t)nb * 3 * 8) != cudaSuccess ||
            cudaMalloc(&dT, (size_t)nb * 289 * 8) != cudaSuccess) return bail(fail("cudaMalloc failed (sk tables)"));
        cudaMemcpy(dV, D->bond_V, (size_t)nb * 25 * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dL, D->bond_lmn, (size_t)nb * 3 * 8, cudaMemcpyHostToDevice);
        int rc = sktb_sk_table(dV, dL, dT, nb, nullptr);
        cudaError_t e = cudaMemcpy(tab.data(), dT, tab.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(dV); cudaFree(dL); cudaFree(dT);
        if (rc) return bail(rc);
        if (e != cudaSuccess) return bail(fail("sk_table kernel failed: %s", cudaGetErrorString(e)));
    }

    // ---- gather orbital sub-blocks, creation order (for H_wo_g) --------------------------------------
    p->bond_ni.resize(nb); p->bond_nj.resize(nb);
    std::vector<long long> off_creation(nb + 1, 0);
    for (int b = 0; b < nb; ++b) {
        int i = D->bond_i[b], j = D->bond_j[b];
        p->bond_ni[b] = D->atom_orb_start[i + 1] - D->atom_orb_start[i];
        p->bond_nj[b] = D->atom_orb_start[j + 1] - D->atom_orb_start[j];
        off_creation[b + 1] = off_creation[b] + (long long)p->bond_ni[b] * p->bond_nj[b];
    }
    p->T_host_creation_order.resize(off_creation[nb]);
    for (int b = 0; b < nb; ++b) {
        int i = D->bond_i[b], j = D->bond_j[b];
        for (int r = 0; r < p->bond_ni[b]; ++r)
            for (int c = 0; c < p->bond_nj[b]; ++c)
                p->T_host_creation_order[off_creation[b] + (long long)r * p->bond_nj[b] + c] =
                    tab[(size_t)b * 289 + D->orb_kind[D->atom_orb_start[i] + r] * 17 + D->orb_kind[D->atom_orb_start[j] + c]];
    }

    // ---- pair-sorted bond arrays ----------------------------------------------------------------------
    std::vector<int> order(nb);
    for (int b = 0; b < nb; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        if (D->bond_i[x] != D->bond_i[y]) return D->bond_i[x] < D->bond_i[y];
        if (D->bond_j[x] != D->bond_j[y]) return D->bond_j[x] < D->bond_j[y];
        return D->bond_image[x] < D->bond_image[y];
    });
    std::vector<int> pair_of((size_t)na * na, -1), pair_begin;
    std::vector<double> bvec((size_t)nb * 3), Tsorted(off_creation[nb]);
    std::vector<long long> boff(nb);
    long long run = 0;
    int npairs = 0;
    for (int s = 0; s < nb; ++s) {
        int b = order[s], i = D->bond_i[b], j = D->bond_j[b];
        if (pair_of[(size_t)i * na + j] < 0) { pair_of[(size_t)i * na + j] = npairs++; pair_begin.push_back(s); }
        for (int c = 0; c < 3; ++c) bvec[3 * s + c] = D->bond_vec[3 * b + c];
        boff[s] = run;
        long long len = (long long)p->bond_ni[b] * p->bond_nj[b];
        std::copy(p->T_host_creation_order.begin() + off_creation[b], p->T_host_creation_order.begin() + off_creation[b] + len, Tsorted.begin() + run);
        run += len;
    }
    pair_begin.push_back(nb);
    p->n_pairs = npairs;
    std::vector<int> orb_atom(n);
    for (int a = 0; a < na; ++a)
        for (int o = D->atom_orb_start[a]; o < D->atom_orb_start[a + 1]; ++o) orb_atom[o] = a;

    // ---- Hermiticity-gate bound -------------------------------------------------------------------------
    // Re(H - H^dagger)[mu,nu] = sum_{b in pair(i,j)} (T_b[mu,nu] - T_rev(b)[nu,mu]) cos(phi_b), rev(b) = the bond
    // (j,i,-d).  A bond b = (i,j,d) and its inversion partner inv(b) = (i,j,-d) of the SAME pair (present for every
    // centrosymmetric neighbour shell: simple cubic, bcc-like two-atom cells, ...) carry equal cos(phi) and are combined
    // before taking magnitudes - this is what makes odd-parity blocks such as s-p or p-f cancel exactly.
    {
        std::map<std::tuple<int, int, long long, long long, long long>, int> by_vec;
        auto key = [&](int i, int j, const double* v, double sgn) {
            return std::make_tuple(i, j, llround(sgn * v[0] * 1e7), llround(sgn * v[1] * 1e7), llround(sgn * v[2] * 1e7));
        };
        for (int b = 0; b < nb; ++b) by_vec[key(D->bond_i[b], D->bond_j[b], D->bond_vec + 3 * b, 1.0)] = b;
        auto find = [&](int i, int j, const double* v, double sgn) {
            auto it = by_vec.find(key(i, j, v, sgn));
            return it == by_vec.end() ? -1 : it->second;
        };
        auto Tb = [&](int b, int r, int c) { return b < 0 ? 0.0 : p->T_host_creation_order[off_creation[b] + (long long)r * p->bond_nj[b] + c]; };
        std::vector<double> acc((size_t)n * n, 0.0);
        for (int b = 0; b < nb; ++b) {
            const int i = D->bond_i[b], j = D->bond_j[b];
            const double* d = D->bond_vec + 3 * b;
            const int ib = find(i, j, d, -1.0);                 // inversion partner in the same pair (may be b itself for d = 0)
            if (ib >= 0 && ib < b) continue;                    // handled together with its partner
            const int rb = find(j, i, d, -1.0);                 // reverse of b
            const int rib = (ib >= 0 && ib != b) ? find(j, i, d, 1.0) : -1;     // reverse of inv(b)
            for (int r = 0; r < p->bond_ni[b]; ++r)
                for (int c = 0; c < p->bond_nj[b]; ++c) {
                    double v = Tb(b, r, c) - Tb(rb, c, r);
                    if (ib >= 0 && ib != b) v += Tb(ib, r, c) - Tb(rib, c, r);
                    acc[(size_t)(D->atom_orb_start[i] + r) * n + D->atom_orb_start[j] + c] += fabs(v);
                }
        }
        double bound = 0.0;
        for (double x : acc) bound = std::max(bound, x);
        std::map<std::pair<int, int>, double> sre;
        for (int t = 0; t < D->soc_nnz; ++t) sre[{D->soc_row[t], D->soc_col[t]}] = D->soc_val[2 * t];
        double sb = 0.0;
        for (auto& kv : sre) {
            auto it = sre.find({kv.first.second, kv.first.first});
            double other = it == sre.end() ? 0.0 : it->second;
            sb = std::max(sb, fabs(kv.second - other));
        }
        p->asym_bound = bound + sb;
        // full (complex) Hermiticity defect: H(k) = H(k)^dagger for every k iff every hopping block equals the transpose of its
        // reverse bond's block and soc_mat is Hermitian.  Decides which Green's-function path reproduces the reference's dense
        // inverse: eigenpairs + Lorentzians (exact when the defect is 0) or one inverse per (k, E) (gf_inverse.cuh).
        double defect = 0.0;
        for (int b = 0; b < nb; ++b) {
            const int rb = find(D->bond_j[b], D->bond_i[b], D->bond_vec + 3 * b, -1.0);
            for (int r = 0; r < p->bond_ni[b]; ++r)
                for (int c = 0; c < p->bond_nj[b]; ++c) defect = std::max(defect, fabs(Tb(b, r, c) - Tb(rb, c, r)));
        }
        p->herm_defect = defect;
        defect = 0.0;
        std::map<std::pair<int, int>, std::pair<double, double>> sc;
        for (int t = 0; t < D->soc_nnz; ++t) sc[{D->soc_row[t], D->soc_col[t]}] = {D->soc_val[2 * t], D->soc_val[2 * t + 1]};
        for (auto& kv : sc) {
            auto it = sc.find({kv.first.second, kv.first.first});
            const double ore = it == sc.end() ? 0.0 : it->second.first, oim = it == sc.end() ? 0.0 : it->second.second;
            defect = std::max(defect, std::max(fabs(kv.second.first - ore), fabs(kv.second.second + oim)));
        }
        p->herm_defect_soc = defect;
    }

    // ---- upload ---------------------------------------------------------------------------------------
    PlanView& V = p->view;
    V.n_atoms = na; V.n_orb = n; V.n_bonds = nb; V.n_pairs = npairs; V.soc_nnz = D->soc_nnz;
    std::vector<cplx> sval(D->soc_nnz);
    for (int t = 0; t < D->soc_nnz; ++t) sval[t] = cmake(D->soc_val[2 * t], D->soc_val[2 * t + 1]);
    int rc = 0;
    rc |= upload(p, D->atom_orb_start, (size_t)na + 1, &V.atom_start);
    rc |= upload(p, orb_atom.data(), (size_t)n, &V.orb_atom);
    rc |= upload(p, pair_of.data(), pair_of.size(), &V.pair_of);
    rc |= upload(p, pair_begin.data(), pair_begin.size(), &V.pair_begin);
    rc |= upload(p, bvec.data(), bvec.size(), &V.bond_vec);
    rc |= upload(p, boff.data(), boff.size(), &V.bond_off);
    rc |= upload(p, Tsorted.data(), Tsorted.size(), &V.T);
    rc |= upload(p, D->onsite, (size_t)n, &V.onsite);
    rc |= upload(p, D->soc_row, (size_t)D->soc_nnz, &V.soc_row);
    rc |= upload(p, D->soc_col, (size_t)D->soc_nnz, &V.soc_col);
    rc |= upload(p, sval.data(), sval.size(), &V.soc_val);
    if (rc) return bail(1);
    for (int c = 0; c < 9; ++c) V.rec[c] = D->rec_lattice[c];

    if (2 * n > 32 && 2 * n <= 64) {          // dense soc_mat for the fused K1 of the CTA-resident 33..64 kernel
        std::vector<cplx> dense((size_t)4 * n * n, cmake(0.0, 0.0));
        for (int t = 0; t < D->soc_nnz; ++t) dense[(size_t)D->soc_row[t] * 2 * n + D->soc_col[t]] = sval[t];
        if (upload(p, dense.data(), dense.size(), &p->soc_dense)) return bail(1);
    }

    // ---- dense tables for the register-resident path (N <= 32) -----------------------------------------
    {
        std::string why;
        p->small_ok = small::build_tables(D, p->T_host_creation_order, off_creation, p->bond_ni, p->bond_nj, &p->small_tab, &p->owned, &why);
        p->small_tab.asym_bound = p->asym_bound;
        if (!p->small_ok && !why.empty()) g_kinfo = "small path unavailable: " + why;
    }

    if (cudaMalloc((void**)&p->d_any, 2 * sizeof(int)) != cudaSuccess) return bail(fail("cudaMalloc failed"));
    p->d_fail = p->d_any + 1;
    cudaMemset(p->d_any, 0, 2 * sizeof(int));
    if (cudaMallocHost((void**)&p->h_pinned, 4 * sizeof(int)) != cudaSuccess) return bail(fail("cudaMallocHost failed"));
    cudaEventCreate(&p->ev0); cudaEventCreate(&p->ev1);
    CK(cudaDeviceSynchronize());
    *out = p;
    return 0;
}

extern "C" void sktb_plan_destroy(sktb_plan* p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    cudaDeviceSynchronize();
    for (void* d : p->owned) cudaFree(d);
    for (auto& w : p->ws) { w.H.release(); w.RT.release(); w.K.release(); w.S.release(); }
    for (Buf* b : {&p->wsE, &p->wsVec, &p->wsFlag, &p->wsPart, &p->wsMisc}) b->release();
    for (int s = 0; s < 2; ++s) {
        p->hK[s].release(); p->hE[s].release(); p->hV[s].release(); p->hF[s].release();
        if (p->hs[s]) cudaStreamDestroy(p->hs[s]);
    }
    if (p->d_any) cudaFree(p->d_any);
    if (p->h_pinned) cudaFreeHost(p->h_pinned);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}

extern "C" int sktb_plan_hoppings(const sktb_plan* p, double* out_host, int64_t capacity) {
    if (!p || !out_host) return fail("null argument");
    if ((int64_t)p->T_host_creation_order.size() > capacity) return fail("capacity %lld < %zu", (long long)capacity, p->T_host_creation_order.size());
    std::copy(p->T_host_creation_order.begin(), p->T_host_creation_order.end(), out_host);
    return 0;
}

extern "C" int sktb_plan_asym_bound(const sktb_plan* p, double* bound) {
    if (!p || !bound) return fail("null argument");
    *bound = p->asym_bound;
    return 0;
}

extern "C" int sktb_plan_herm_defect(const sktb_plan* p, int32_t soc, double* defect) {
    if (!p || !defect) return fail("null argument");
    *defect = soc ? std::max(p->herm_defect, p->herm_defect_soc) : p->herm_defect;
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------
static int launch_hk(sktb_plan* p, const double* k_d, long long nk, int soc, cplx* H, int* flag, long long flag_off, cudaStream_t st) {
    size_t smem = (size_t)std::max(p->n_bonds, 1) * sizeof(cplx);
    if (smem > 200 * 1024) return fail("model has %d bonds: phase table exceeds shared memory", p->n_bonds);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(hk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int N = soc ? 2 * p->n_orb : p->n_orb;
    int threads = N * N >= 1024 ? 256 : (N * N >= 256 ? 128 : 64);
    unsigned grid = (unsigned)std::min<long long>(nk, 148LL * 16);
    hk_kernel<<<grid, threads, smem, st>>>(p->view, k_d, nk, soc, H, flag, flag_off);
    LAUNCHED();
    return 0;
}

static int launch_heev_generic_phase(sktb_plan* p, HeevArgs a, cudaStream_t st, int phase) {
    const int N = a.N;
    static const long long smemA_max = getenv("SKTB_GENERIC_SMEMA_KB") ? atoll(getenv("SKTB_GENERIC_SMEMA_KB")) * 1024 : 72 * 1024;
    const int NP = std::max(32, (N + 31) / 32 * 32);
    int T = std::min(1024, NP);
    a.phase = phase;
    if (phase == 2) T = std::min(512, std::max(64, (a.n_rows + 31) / 32 * 32));
    size_t aux = ((size_t)4 * N + 64) * sizeof(double) + (size_t)3 * N * sizeof(cplx);
    size_t withA = aux + (size_t)N * N * sizeof(cplx);
    bool smemA = (long long)withA <= smemA_max && T <= 384 && phase == 0;   // beyond N = 64 many CTAs per SM over L2 beat one CTA per SM in shared memory
    // panel formulation for the sizes that stream the matrix through HBM/L2 (one column per thread, N <= 1024)
    static const int nb_env = getenv("SKTB_GENERIC_NB") ? atoi(getenv("SKTB_GENERIC_NB")) : 8;
    static const int remap_min = getenv("SKTB_GENERIC_REMAP_MIN") ? atoi(getenv("SKTB_GENERIC_REMAP_MIN")) : 256;   // same-box A/B: +1..2 % from 256 threads up, nothing below
    a.remap_min_threads = remap_min;
    a.NB = 0;
    if (!smemA && N <= 1024 && phase != 2 && (phase == 3 || (nb_env > 0 && (N >= 96 || getenv("SKTB_GENERIC_NB"))))) {
        int NB = nb_env >= 8 ? 8 : (nb_env >= 4 ? 4 : 2);
        if (!getenv("SKTB_GENERIC_NB")) NB = N >= 160 ? 8 : 4;            // measured: 4 wins below N ~ 160, none below N ~ 96
        if (phase == 3) NB = (N >= 160 && a.peel % 8 == 0) ? 8 : 4;       // the peel count is a multiple of NB
        auto panel_bytes = [&](int nb) { return (size_t)2 * nb * N * sizeof(cplx) + (32 * 32 + 32) * sizeof(double) + (size_t)T * sizeof(cplx); };
        while (NB > 2 && aux + panel_bytes(NB) > 200 * 1024) NB /= 2;
        if (aux + panel_bytes(NB) <= 200 * 1024) { a.NB = NB; aux += panel_bytes(NB); }
    }
    size_t smem = smemA ? withA : aux;
    a.fail = p->d_fail;
    unsigned grid = (unsigned)std::min<long long>(a.nm, 148LL * 32);
    // Three builds: A in shared memory (N <= 64, the fallback of the register-resident paths), A in global memory
    // with a free register budget (T <= 384), and the 64-register build that keeps two 512-thread CTAs per SM.
    auto kern = smemA ? heev_generic_kernel<true, 384> : (T <= 384 ? heev_generic_kernel<false, 384> : (a.NB > 0 && T <= 512 ? heev_generic_kernel<false, 512> : heev_generic_kernel<false, 1024>));
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, T, smem, st>>>(a);
    LAUNCHED();
    char buf[256];
    snprintf(buf, sizeof buf, "heev_generic<smemA=%d> N=%d NB=%d threads=%d smem=%zuB grid=%u rows=%d", (int)smemA, N, a.NB, T, smem, grid, a.n_rows);
    if (phase == 2) g_kinfo += std::string(" + QL kernel x") + std::to_string(T) + " thr"; else g_kinfo = buf;
    return 0;
}

// Tracked rows with N > 64: tridiagonalisation (one 1-CTA-per-SM class kernel) and the serial-generator QL (small CTAs,
// many matrices per SM) are two launches; everything else is one.
// Peeled run (64 < N <= PEEL_MAX, eigenvalues only): the generic kernel does the first P steps (P = the multiple of the
// panel width 4 or 8 that leaves a trailing block of 57..64), the register-resident kernels tridiagonalise that block in place, one
// bisection kernel takes the merged (d, e).
static const int PEEL_MAX = getenv("SKTB_PEEL_MAX") ? atoi(getenv("SKTB_PEEL_MAX")) : 320;
static size_t peel_bytes_per_k(int N) { return (size_t)2 * (N - 56) * 8 + small::cta64_bytes_per_k(); }   // P <= N - 57
// Two-stage tridiagonalisation (heev_twostage.cuh: dense -> band 16 on DMMA, band -> tridiagonal by bulge chasing) + bisection:
// eigenvalues only, N >= SKTB_TWOSTAGE_MIN.  The one-stage kernels read the trailing matrix once per reflector and are HBM-bound
// from N ~ 300 up; here the matrix is read twice and written once per 16 reflectors.
static int twostage_min() { static const int v = getenv("SKTB_TWOSTAGE_MIN") ? atoi(getenv("SKTB_TWOSTAGE_MIN")) : 208; return v; }   // measured, matrices/s one-stage / two-stage: N = 192 9.5e4 / 1.01e5 at 592 matrices but 1.17e5 / 1.01e5 at 8192; N = 208 7.0e4 / 8.8e4; N = 256 4.0e4 / 6.1e4
static bool use_twostage(int N) { return N >= twostage_min() && twostage::supported(N); }
// scratch per matrix of the eigenvalues-only paths above N = 64 (HeevArgs::peel_ws): peeled run or two-stage
static size_t evals_ws_bytes_per_k(int N) {
    if (N <= 64) return 0;
    if (use_twostage(N)) return twostage::scratch_bytes_per_matrix(N);
    return N <= PEEL_MAX ? peel_bytes_per_k(N) : 0;
}
static int launch_heev_peel(sktb_plan* p, HeevArgs a, cudaStream_t st) {
    const int N = a.N;
    const int nb = N >= 160 ? 8 : 4;
    const int P = (N - 64 + nb - 1) / nb * nb, Nt = N - P;
    double* head = (double*)a.peel_ws;
    double* c64s = head + (size_t)a.nm * 2 * P;
    a.peel = P; a.peel_de = head;
    int rc = launch_heev_generic_phase(p, a, st, 3);
    if (rc) return rc;
    std::string first = g_kinfo, info;
    int rl = small::launch_cta64(a.H, Nt, a.nm, c64s, nullptr, 0, 0, p->d_fail, st, &info, N, (long long)N * N, (long long)P * N + P, true);
    if (rl < 0) return fail("cta64 launch failed (peeled run): %s", info.c_str());
    g_launches += rl;
    // tridiagonal eigenvalues: per-lane QL (32 matrices per warp, ~40x fewer instructions than bisection, but one long
    // serial chain per batch) once the batch fills the device, bisection (one eigenvalue per thread) for small batches
    static const long long ql_min = getenv("SKTB_PEEL_QL_MIN") ? atoll(getenv("SKTB_PEEL_QL_MIN")) : 2048;
    const size_t per_warp = (size_t)2 * N * small::DSTRIDE * sizeof(double);
    const int warps = (int)std::min<size_t>(small::SMALL_WARPS, (200 * 1024) / per_warp);
    if (a.nm >= ql_min && warps >= 1) {
        const size_t smem = per_warp * warps;
        auto kM = small::tql_mode_env() == 2 ? small::tql_merge_kernel<2> : small::tql_merge_kernel<1>;
        CK(cudaFuncSetAttribute(kM, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long batches = (a.nm + 31) / 32;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((batches + warps - 1) / warps, 148));
        kM<<<grid, warps * 32, smem, st>>>(head, P, c64s, Nt, a.nm, a.evals, a.ld, a.col0, p->d_fail);
        LAUNCHED();
        g_kinfo = "peel " + std::to_string(P) + " steps: " + first + " | " + info + " | tql_merge x" + std::to_string(warps) + " warps";
        return 0;
    }
    const int T = (N + 31) / 32 * 32;
    const size_t smem = ((size_t)2 * N + 96) * sizeof(double);
    bisect_merge_kernel<<<(unsigned)std::min<long long>(a.nm, 148LL * 8), T, smem, st>>>(head, P, c64s, Nt, a.nm, a.evals, a.ld, a.col0);
    LAUNCHED();
    g_kinfo = "peel " + std::to_string(P) + " steps: " + first + " | " + info + " | bisect_merge";
    return 0;
}

// chunk size of the two-stage path: whole waves of both persistent grids (148 CTAs of stage 1, 4 x 148 of the bulge chase)
static long long twostage_chunk(int N, long long n, size_t bytes_per_k) {
    long long c = std::max<long long>(1, std::min<long long>(n, (long long)(((size_t)5 << 30) / bytes_per_k)));
    if (c >= n) return n;
    if (c >= 592) c = c / 592 * 592; else if (c >= 148) c = c / 148 * 148;
    return c;
}
static int launch_heev_twostage(sktb_plan* p, HeevArgs a, cudaStream_t st) {
    const int N = a.N;
    unsigned char* ws = reinterpret_cast<unsigned char*>(a.peel_ws);
    double* de = reinterpret_cast<double*>(ws + (size_t)a.nm * (twostage::scratch_bytes_per_matrix(N) - (size_t)2 * N * sizeof(double)));
    char info[320];
    const int rl = twostage::launch_tridiag(a.H, N, a.nm, ws, de, st, info, sizeof info);
    if (rl < 0) return fail("two-stage tridiagonalisation does not support N = %d", N);
    g_launches += rl;
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail("two-stage launch: %s", cudaGetErrorString(e_)); }
    static const bool old_bisect = getenv("SKTB_TWOSTAGE_BISECT") && getenv("SKTB_TWOSTAGE_BISECT")[0] == '0';   // A/B: quotient-form Sturm count
    if (old_bisect) {
        const int T = (N + 31) / 32 * 32;
        const size_t smem = ((size_t)2 * N + 96) * sizeof(double);
        bisect_merge_kernel<<<(unsigned)std::min<long long>(a.nm, 148LL * 8), T, smem, st>>>(de, N, de, 0, a.nm, a.evals, a.ld, a.col0);
    } else if (twostage::launch_bisect(de, N, a.nm, a.evals, a.ld, a.col0, st) < 0) return fail("bisection launch failed (N = %d)", N);
    LAUNCHED();
    g_kinfo = std::string(info) + (old_bisect ? " | bisect_merge" : " | bisect_prod (product-form Sturm count)");
    return 0;
}

static int launch_heev_generic(sktb_plan* p, HeevArgs a, cudaStream_t st) {
    if (a.n_rows == 0 && !a.vstore && a.peel_ws && use_twostage(a.N)) return launch_heev_twostage(p, a, st);
    if (a.n_rows == 0 && a.peel_ws && a.N > 64 && a.N <= PEEL_MAX) return launch_heev_peel(p, a, st);
    static const bool no_split = getenv("SKTB_GENERIC_SPLIT") != nullptr && getenv("SKTB_GENERIC_SPLIT")[0] == '0';
    if (a.n_rows > 0 && a.N > 64 && a.N <= 1024 && !no_split) {
        int rc = launch_heev_generic_phase(p, a, st, 1);
        return rc ? rc : launch_heev_generic_phase(p, a, st, 2);
    }
    return launch_heev_generic_phase(p, a, st, 0);
}

static const size_t WS_BUDGET = (size_t)3 << 30;   // per-buffer cap for chunked workspaces

// ------------------------------------------------------------------------------------------------------
// Eigenvectors for 64 < N <= 1024 (kernels_big_vec.cuh): panel tridiagonalisation keeping its reflectors + bisection,
// inverse iteration on the tridiagonal (one eigenvector per thread), back-transformation (one eigenvector per warp).
// H: [nm][N][N] workspace (destroyed).  vec_out[band * vs_band + (vcol0 + mat) * vs_k + r]: all components (n_rows == 0)
// or only the components rows_d[r].  scratch: >= nm * bigvec_bytes_per_matrix(N).  SKTB_BIGVEC=0: row tracking as before.
// ------------------------------------------------------------------------------------------------------
static bool bigvec_enabled(int N) {
    static const bool off = getenv("SKTB_BIGVEC") && getenv("SKTB_BIGVEC")[0] == '0';
    return !off && N > 64 && N <= 1024;
}
static const size_t BIGVEC_BUDGET = (size_t)12 << 30;
template <int NQ>
static int launch_vecbt_big(const BigVecArgs& b, cudaStream_t st, int sms) {
    constexpr int BW = 8;
    const size_t smem = vecbt_big_smem(NQ);
    constexpr int EV = NQ == 8 ? 2 : 1;          // eigenvectors per warp (SKTB_VBT_EV=1: one). Measured: N=256 1.73e4 -> 2.02e4 matrices/s; N=512 (234 registers, one CTA per SM) slower
    static const bool ev1 = getenv("SKTB_VBT_EV") && getenv("SKTB_VBT_EV")[0] == '1';
    void (*kern)(BigVecArgs) = b.n_rows > 0 ? rowsq_big_kernel<NQ, BW> : (ev1 ? vecbt_big_kernel<NQ, BW, 1> : vecbt_big_kernel<NQ, BW, EV>);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BW * 32, smem);
    const int per_cta = b.n_rows > 0 ? BW : BW * (ev1 ? 1 : EV);
    const long long work = b.nm * (((b.n_rows > 0 ? b.n_rows : b.N) + per_cta - 1) / per_cta);
    kern<<<(unsigned)std::max<long long>(1, std::min<long long>(work, (long long)sms * std::max(occ, 1))), BW * 32, smem, st>>>(b);
    LAUNCHED();
    return 0;
}
// Selected rows (LDOS / spectral projections) on top of the two-stage reduction: scratch holds the bigvec layout (the reflector region
// takes the chase reflectors and the T factors), ts_scratch the two-stage workspace.  N >= SKTB_TWOSTAGE_MIN only.
static bool bigvec_twostage(int N, int n_rows) {
    static const bool off = getenv("SKTB_TWOSTAGE_ROWS") && getenv("SKTB_TWOSTAGE_ROWS")[0] == '0';
    return !off && n_rows > 0 && use_twostage(N) &&
           ((size_t)(twostage::chase_reflectors(N) + 1) * (twostage::SB + 1) + (size_t)twostage::stage1_panels(N) * twostage::SB * twostage::SB) <= (size_t)N * N;
}
static int launch_bigvec(sktb_plan* p, cplx* H, int N, long long nm, double* evals_d, long long ld, long long col0, cplx* vec_out,
                         long long vs_band, long long vs_k, long long vcol0, int n_rows, const int* rows_d, void* scratch, cudaStream_t st,
                         void* ts_scratch = nullptr) {
    const size_t nn = (size_t)N * N;
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);      // 16-byte types first: every segment stays aligned for odd N
    cplx* vstore = reinterpret_cast<cplx*>(base);                      base += (size_t)nm * nn * 16;
    cplx* taus = reinterpret_cast<cplx*>(base);                        base += (size_t)nm * N * 16;
    double* Z = reinterpret_cast<double*>(base);                       base += (size_t)nm * nn * 8;
    double* lu = reinterpret_cast<double*>(base);                      base += (size_t)nm * 4 * nn * 8;
    double* de = reinterpret_cast<double*>(base);                      base += (size_t)nm * 2 * N * 8;
    unsigned char* pin = base;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ts_scratch && bigvec_twostage(N, n_rows)) {
        const long long rstride = (twostage::chase_reflectors(N) + 1) * (twostage::SB + 1);
        cplx* Rstore = vstore;                                            // [nm][nn] region: chase reflectors, then the T factors
        cplx* Tstore = vstore + (size_t)nm * rstride;
        if ((size_t)nm * rstride + (size_t)nm * twostage::stage1_panels(N) * twostage::SB * twostage::SB > (size_t)nm * nn) return fail("two-stage rows: reflector store too small");
        char info[320];
        const int rl = twostage::launch_tridiag_keep(H, N, nm, ts_scratch, de, Tstore, Rstore, rstride, st, info, sizeof info);
        if (rl < 0) return fail("two-stage tridiagonalisation does not support N = %d", N);
        g_launches += rl;
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail("two-stage launch: %s", cudaGetErrorString(e_)); }
        BigVecArgs b{};
        b.N = N; b.nm = nm; b.de = de; b.evals = evals_d; b.ld = ld; b.col0 = col0;
        b.Z = Z; b.lu = lu; b.pin = pin; b.vstore = vstore; b.taustore = taus;
        b.vec_out = vec_out; b.vs_band = vs_band; b.vs_k = vs_k; b.vcol0 = vcol0; b.n_rows = n_rows; b.rows = rows_d;
        const int T = (N + 31) / 32 * 32;
        stein_kernel<<<(unsigned)std::min<long long>(nm, (long long)sms * 8), T, (size_t)6 * N * 8, st>>>(b);
        LAUNCHED();
        twostage::RowsArgs ra{};
        ra.N = N; ra.nm = nm; ra.H = H; ra.Tstore = Tstore; ra.npanels = twostage::stage1_panels(N); ra.Rstore = Rstore; ra.rstride = rstride;
        ra.Zf = lu; ra.zstride = (long long)4 * N * N;
        ra.vec_out = vec_out; ra.vs_band = vs_band; ra.vs_k = vs_k; ra.vcol0 = vcol0; ra.n_rows = n_rows; ra.rows = rows_d;
        if (twostage::launch_rows(ra, st) < 0) return fail("two-stage rows kernel launch failed");
        LAUNCHED();
        g_kinfo = std::string(info) + " (reflectors kept) + stein (bisection + inverse iteration, 1 eigenpair/thread) + rows_kernel (selected rows of Q1 Q2, then of U)";
        return 0;
    }
    HeevArgs a{};
    a.H = H; a.N = N; a.nm = (int)nm;
    a.evals = evals_d; a.ld = ld; a.col0 = col0;
    a.vstore = vstore; a.taustore = taus; a.de_out = de;
    int rc = launch_heev_generic_phase(p, a, st, 1);                  // tridiagonalisation only: (d, e) and the reflectors are kept
    if (rc) return rc;
    const std::string first = g_kinfo;
    BigVecArgs b{};
    b.N = N; b.nm = nm; b.de = de; b.evals = evals_d; b.ld = ld; b.col0 = col0;
    b.Z = Z; b.lu = lu; b.pin = pin; b.vstore = vstore; b.taustore = taus;
    b.vec_out = vec_out; b.vs_band = vs_band; b.vs_k = vs_k; b.vcol0 = vcol0; b.n_rows = n_rows; b.rows = rows_d;
    const int T = (N + 31) / 32 * 32;
    stein_kernel<<<(unsigned)std::min<long long>(nm, (long long)sms * 8), T, (size_t)6 * N * 8, st>>>(b);
    LAUNCHED();
    if (N <= 128) rc = launch_vecbt_big<4>(b, st, sms);
    else if (N <= 256) rc = launch_vecbt_big<8>(b, st, sms);
    else if (N <= 512) rc = launch_vecbt_big<16>(b, st, sms);
    else rc = launch_vecbt_big<32>(b, st, sms);
    if (rc) return rc;
    g_kinfo = first + " (reflectors kept) + stein (bisection + inverse iteration, 1 eigenpair/thread) + " +
              (n_rows > 0 ? "rowsq_big (selected rows of Q, then rows of U = Q Z)" : "vecbt_big (1 eigenvector/warp)");
    return 0;
}

// 32 < N <= 64: K0 + K1 fused into tridiag_cta64 (H(k) is assembled in shared memory / registers, never in HBM) whenever the
// per-k Hermiticity flags are not needed.  SKTB_CTA64_FUSE=0: materialise H with hk_kernel as before (A/B, parity tests).
static bool c64_fusable(const sktb_plan* p, int N, int soc, bool need_flags) {
    static const bool off = getenv("SKTB_CTA64_FUSE") && getenv("SKTB_CTA64_FUSE")[0] == '0';
    if (off || N <= 32 || N > 64 || need_flags) return false;
    if (soc && !p->soc_dense) return false;
    const size_t n = (size_t)p->n_orb;
    return (n * (n + 1) / 2 + (size_t)p->n_bonds) * sizeof(cplx) <= 96 * 1024;
}
static small::C64Build c64_build(const sktb_plan* p, const double* k, const int32_t* mesh, long long first, int soc) {
    return small::C64Build{p->view, k, mesh, first, soc, soc ? p->soc_dense : nullptr};
}

// (d,e) scratch of the register-resident path: 2*NC doubles per k-point, at most 2M k-points per launch pair.
// The pair-layout pipeline (24 < N <= 32) keeps two chunks in flight and parks a 4 KB trailing block per k-point.
static size_t small_scratch_bytes(int N, long long nk) {
    if (small::padded_nc(N) == 32) {
        static const int chunk_log2 = getenv("SKTB_PAIR_CHUNK_LOG2") ? atoi(getenv("SKTB_PAIR_CHUNK_LOG2")) : 20;
        long long chunk = std::min<long long>(std::max<long long>(nk, 32), 1LL << chunk_log2);
        chunk = (chunk + 31) / 32 * 32;
        return (size_t)chunk * 2 * small::PAIR_BYTES_PER_K;
    }
    long long chunk = std::min<long long>(std::max<long long>(nk, 32), 1LL << 21);
    chunk = (chunk + 31) / 32 * 32;
    return (size_t)chunk * 2 * small::padded_nc(N) * 8;
}

// shared body of sktb_solve / sktb_solve_mesh.  k_d == NULL -> mesh.
static int solve_impl(sktb_plan* p, const double* k_d, const int32_t* nk_mesh, long long first, long long nk, int soc,
                      double* evals_d, double* evecs_d, long long ld, long long k_off, int* flags_d, cudaStream_t st, bool timed,
                      int ws_id = 2) {
    if (!p || !evals_d) return fail("null argument");
    if (nk <= 0) return 0;
    sktb_plan::WorkSet& W = p->ws[ws_id];
    ON_DEVICE(p->device);
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    int launches0 = (int)g_launches.load();
    if (timed) CK(cudaEventRecord(p->ev0, st));

    const bool want_vecs = evecs_d != nullptr;
    if (!want_vecs && p->small_ok && small::supports(N)) {
        size_t sbytes = small_scratch_bytes(N, nk);
        if (W.S.ensure(sbytes)) return fail("out of device memory (tridiagonal scratch %zu B)", sbytes);
        int rc = small::launch_solve(p->small_tab, k_d, nk_mesh, first, nk, soc, N, evals_d, ld, k_off, flags_d, k_off, p->d_fail,
                                     (double*)W.S.p, sbytes, st, &g_kinfo);
        if (rc < 0) return fail("small-path launch failed: %s", g_kinfo.c_str());
        g_launches += rc;
    } else if (want_vecs && p->small_ok && small::supports(N) && !(p->asym_bound > 0.5e-9)) {
        // N <= 32 with eigenvectors: pair-layout tridiagonalisation keeping the reflectors + vec32_kernel
        const size_t per_k = small::vec32_bytes_per_k();
        long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(WS_BUDGET / per_k)));
        if (W.RT.ensure(per_k * chunk)) return fail("out of device memory (reflector workspace %zu B)", per_k * chunk);
        for (long long c0 = 0; c0 < nk; c0 += chunk) {
            long long cnt = std::min(chunk, nk - c0);
            int rc = small::launch_vec32(p->small_tab, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, first + c0, cnt, soc, N, evals_d, (cplx*)evecs_d, ld,
                                         k_off + c0, flags_d, k_off + c0, p->d_fail, W.RT.p, st, &g_kinfo);
            if (rc < 0) return fail("vec32 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
    } else if (N <= 32) {
        // N <= 32 models the fused register path cannot take (bond tables larger than shared memory, or eigenvectors of a
        // model whose Hermiticity gate can fire): K1 materialises H(k), the pair kernels read it in LOAD mode
        const size_t perH = (size_t)N * N * sizeof(cplx), per_k = perH + small::vec32_bytes_per_k() + 24;
        long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(WS_BUDGET / per_k)));
        if (W.H.ensure(perH * chunk) || W.RT.ensure(small::vec32_bytes_per_k() * chunk)) return fail("out of device memory (heev32 workspaces)");
        if (!k_d && W.K.ensure((size_t)chunk * 3 * 8)) return fail("out of device memory (k workspace)");
        for (long long c0 = 0; c0 < nk; c0 += chunk) {
            long long cnt = std::min(chunk, nk - c0);
            const double* kk = k_d ? k_d + 3 * c0 : (const double*)W.K.p;
            if (!k_d) { int rc = sktb_kmesh(nk_mesh, first + c0, cnt, (double*)W.K.p, st); if (rc) return rc; }
            int rc = launch_hk(p, kk, cnt, soc, (cplx*)W.H.p, flags_d, k_off + c0, st);
            if (rc) return rc;
            rc = small::launch_heev32((const cplx*)W.H.p, N, cnt, evals_d, (cplx*)evecs_d, ld, k_off + c0, p->d_fail, W.RT.p, st, &g_kinfo);
            if (rc < 0) return fail("heev32 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
    } else if (want_vecs && N > 32 && N <= 64) {
        // K1 -> tridiag_cta64 keeping the reflectors -> vecql64 -> vecbt64
        const size_t perH = (size_t)N * N * sizeof(cplx), per_k = perH + small::vec64_bytes_per_k() + 24;
        long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(WS_BUDGET / per_k)));
        if ((!c64_fusable(p, N, soc, flags_d && p->asym_bound > 0.5e-9) && W.H.ensure(perH * chunk)) || W.RT.ensure(small::vec64_bytes_per_k() * chunk))
            return fail("out of device memory (vec64 workspaces)");
        if (!k_d && W.K.ensure((size_t)chunk * 3 * 8)) return fail("out of device memory (k workspace)");
        const bool fuse = c64_fusable(p, N, soc, flags_d && p->asym_bound > 0.5e-9);
        for (long long c0 = 0; c0 < nk; c0 += chunk) {
            long long cnt = std::min(chunk, nk - c0);
            int rc;
            if (fuse) {
                if (flags_d) CK(cudaMemsetAsync(flags_d + k_off + c0, 0, (size_t)cnt * sizeof(int), st));
                const small::C64Build b = c64_build(p, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, first + c0, soc);
                rc = small::launch_vec64(nullptr, N, cnt, evals_d, (cplx*)evecs_d, ld, k_off + c0, p->d_fail, W.RT.p, st, &g_kinfo, &b);
            } else {
                const double* kk = k_d ? k_d + 3 * c0 : (const double*)W.K.p;
                if (!k_d) { rc = sktb_kmesh(nk_mesh, first + c0, cnt, (double*)W.K.p, st); if (rc) return rc; }
                rc = launch_hk(p, kk, cnt, soc, (cplx*)W.H.p, flags_d, k_off + c0, st);
                if (rc) return rc;
                rc = small::launch_vec64((const cplx*)W.H.p, N, cnt, evals_d, (cplx*)evecs_d, ld, k_off + c0, p->d_fail, W.RT.p, st, &g_kinfo);
            }
            if (rc < 0) return fail("vec64 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
    } else if (!want_vecs && N > 32 && N <= 64) {
        // K1 materialises H(k) (64 KB per k at N = 64), tridiag_cta64_kernel keeps it in registers, QL per lane
        const bool fuse = c64_fusable(p, N, soc, flags_d && p->asym_bound > 0.5e-9);
        size_t perH = fuse ? 0 : (size_t)N * N * sizeof(cplx);
        long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(WS_BUDGET / (perH + small::cta64_bytes_per_k()))));
        if ((!fuse && W.H.ensure(perH * chunk)) || W.S.ensure((size_t)chunk * small::cta64_bytes_per_k())) return fail("out of device memory (H workspace %zu B)", perH * chunk);
        if (!fuse && !k_d && W.K.ensure((size_t)chunk * 3 * 8)) return fail("out of device memory (k workspace)");
        for (long long c0 = 0; c0 < nk; c0 += chunk) {
            long long cnt = std::min(chunk, nk - c0);
            int rc;
            if (fuse) {
                if (flags_d) CK(cudaMemsetAsync(flags_d + k_off + c0, 0, (size_t)cnt * sizeof(int), st));
                const small::C64Build b = c64_build(p, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, first + c0, soc);
                rc = small::launch_cta64(nullptr, N, cnt, (double*)W.S.p, evals_d, ld, k_off + c0, p->d_fail, st, &g_kinfo, 0, 0, 0, false, &b);
            } else {
                const double* kk = k_d ? k_d + 3 * c0 : (const double*)W.K.p;
                if (!k_d) { rc = sktb_kmesh(nk_mesh, first + c0, cnt, (double*)W.K.p, st); if (rc) return rc; }
                rc = launch_hk(p, kk, cnt, soc, (cplx*)W.H.p, flags_d, k_off + c0, st);
                if (rc) return rc;
                rc = small::launch_cta64((const cplx*)W.H.p, N, cnt, (double*)W.S.p, evals_d, ld, k_off + c0, p->d_fail, st, &g_kinfo);
            }
            if (rc < 0) return fail("cta64 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
    } else {
        size_t perH = (size_t)N * N * sizeof(cplx);
        const bool bigvec = want_vecs && bigvec_enabled(N);
        long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(WS_BUDGET / perH)));
        if (bigvec) chunk = std::max<long long>(1, std::min<long long>(nk, (long long)(BIGVEC_BUDGET / (perH + bigvec_bytes_per_matrix(N)))));
        if (W.H.ensure(perH * chunk)) return fail("out of device memory (H workspace %zu B)", perH * chunk);
        if (want_vecs && W.RT.ensure(bigvec ? bigvec_bytes_per_matrix(N) * chunk : perH * chunk)) return fail("out of device memory (RT workspace)");
        const bool peel = !want_vecs && evals_ws_bytes_per_k(N) > 0;
        if (peel && use_twostage(N)) {
            chunk = twostage_chunk(N, nk, perH);
            if (W.H.ensure(perH * chunk)) return fail("out of device memory (H workspace %zu B)", perH * chunk);
        }
        if (peel && W.S.ensure(evals_ws_bytes_per_k(N) * chunk)) return fail("out of device memory (peel workspace)");
        if (!k_d && W.K.ensure((size_t)chunk * 3 * 8)) return fail("out of device memory (k workspace)");
        for (long long c0 = 0; c0 < nk; c0 += chunk) {
            long long cnt = std::min(chunk, nk - c0);
            const double* kk = k_d ? k_d + 3 * c0 : (const double*)W.K.p;
            if (!k_d) { int rc = sktb_kmesh(nk_mesh, first + c0, cnt, (double*)W.K.p, st); if (rc) return rc; }
            int rc = launch_hk(p, kk, cnt, soc, (cplx*)W.H.p, flags_d, k_off + c0, st);
            if (rc) return rc;
            if (bigvec) {
                rc = launch_bigvec(p, (cplx*)W.H.p, N, cnt, evals_d, ld, k_off + c0, (cplx*)evecs_d, ld * N, N, k_off + c0, 0, nullptr, W.RT.p, st);
                if (rc) return rc;
                continue;
            }
            HeevArgs a{};
            a.H = (cplx*)W.H.p; a.N = N; a.nm = (int)cnt;
            a.evals = evals_d; a.ld = ld; a.col0 = k_off + c0;
            if (peel) a.peel_ws = W.S.p;
            if (want_vecs) {
                a.RT = (cplx*)W.RT.p; a.n_rows = N; a.rows = nullptr;
                a.vec_out = (cplx*)evecs_d; a.vs_band = ld * N; a.vs_k = N; a.vcol0 = k_off + c0;
            }
            rc = launch_heev_generic(p, a, st);
            if (rc) return rc;
        }
    }
    if (timed) { CK(cudaEventRecord(p->ev1, st)); p->timing_valid = true; p->last_launches = (int)g_launches.load() - launches0; }
    return 0;
}

extern "C" int sktb_hk(sktb_plan* p, const double* k_d, int64_t nk, int32_t soc, double* H_d, int32_t* nonherm_d, void* stream) {
    if (!p || !k_d || !H_d) return fail("null argument");
    if (nk <= 0) return 0;
    ON_DEVICE(p->device);
    return launch_hk(p, k_d, nk, soc, (cplx*)H_d, nonherm_d, 0, (cudaStream_t)stream);
}

extern "C" int sktb_solve(sktb_plan* p, const double* k_d, int64_t nk, int32_t soc, double* evals_d, double* evecs_d,
                          int64_t ld_k, int64_t k_off, int32_t* nonherm_d, void* stream) {
    if (!k_d) return fail("null k pointer");
    return solve_impl(p, k_d, nullptr, 0, nk, soc, evals_d, evecs_d, ld_k, k_off, nonherm_d, (cudaStream_t)stream, true);
}

extern "C" int sktb_solve_mesh(sktb_plan* p, const int32_t nk_mesh[3], int64_t first, int64_t count, int32_t soc,
                               double* evals_d, int64_t ld_k, int64_t k_off, int32_t* nonherm_d, void* stream) {
    if (!nk_mesh) return fail("null mesh");
    long long total = (long long)nk_mesh[0] * nk_mesh[1] * nk_mesh[2];
    if (first < 0 || first + count > total) return fail("mesh range [%lld,%lld) outside %lld points", (long long)first, (long long)(first + count), total);
    return solve_impl(p, nullptr, nk_mesh, first, count, soc, evals_d, nullptr, ld_k, k_off, nonherm_d, (cudaStream_t)stream, true);
}

extern "C" int sktb_ql_failures(sktb_plan* p, int32_t* count) {
    if (!p || !count) return fail("null argument");
    ON_DEVICE(p->device);
    CK(cudaMemcpy(count, p->d_fail, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sktb_last_solve_time(sktb_plan* p, double* ms, int32_t* n_launches) {
    if (!p || !p->timing_valid) return fail("no timed solve on this plan");
    CK(cudaEventSynchronize(p->ev1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, p->ev0, p->ev1));
    if (ms) *ms = t;
    if (n_launches) *n_launches = p->last_launches;
    return 0;
}

// Batched Hermitian eigensolve of caller-supplied matrices (Hamiltonian._sol_ham, hamiltonian.py:108-126):
// gate flags on the matrices as given, then K2 on a workspace copy (K2 destroys its input).
__global__ void herm_gate_kernel(const cplx* __restrict__ H, int N, long long nm, int* __restrict__ flag) {
    __shared__ int s_bad;
    for (long long m = blockIdx.x; m < nm; m += gridDim.x) {
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
        const cplx* A = H + (size_t)m * N * N;
        int bad = 0;
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
            int r = e / N, c = e - r * N;
            if (A[e].x - A[(size_t)c * N + r].x > 1.0e-9) bad = 1;
        }
        if (bad) s_bad = 1;
        __syncthreads();
        if (threadIdx.x == 0) flag[m] = s_bad;
        __syncthreads();
    }
}

extern "C" int sktb_heev(sktb_plan* p, const double* H_d, int32_t N, int64_t nm, double* evals_d, double* evecs_d, int32_t* nonherm_d,
                         void* stream) {
    if (!p || !H_d || !evals_d) return fail("null argument");
    if (N <= 0) return fail("bad matrix size");
    if (nm <= 0) return 0;
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    size_t perH = (size_t)N * N * sizeof(cplx);
    if (nonherm_d) {
        herm_gate_kernel<<<(unsigned)std::min<long long>(nm, 148LL * 8), 128, 0, st>>>((const cplx*)H_d, N, nm, nonherm_d);
        LAUNCHED();
    }
    static const bool heev_generic_only = getenv("SKTB_HEEV_GENERIC") != nullptr && getenv("SKTB_HEEV_GENERIC")[0] == '1';
    if (N <= 32 && !heev_generic_only) {                  // register-resident pair kernels, input read in place
        const size_t per_k = small::vec32_bytes_per_k();
        long long chunk = std::max<long long>(1, std::min<long long>(nm, (long long)(WS_BUDGET / per_k)));
        if (p->ws[2].RT.ensure(per_k * chunk)) return fail("out of device memory (heev workspace)");
        for (long long c0 = 0; c0 < nm; c0 += chunk) {
            long long cnt = std::min(chunk, (long long)nm - c0);
            int rc = small::launch_heev32((const cplx*)H_d + (size_t)c0 * N * N, N, cnt, evals_d, (cplx*)evecs_d, nm, c0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
            if (rc < 0) return fail("heev32 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
        return 0;
    }
    if (N <= 64 && evecs_d && !heev_generic_only) {       // CTA-resident kernel keeping the reflectors + vec64
        const size_t per_k = small::vec64_bytes_per_k();
        long long chunk = std::max<long long>(1, std::min<long long>(nm, (long long)(WS_BUDGET / per_k)));
        if (p->ws[2].RT.ensure(per_k * chunk)) return fail("out of device memory (heev workspace)");
        for (long long c0 = 0; c0 < nm; c0 += chunk) {
            long long cnt = std::min(chunk, (long long)nm - c0);
            int rc = small::launch_vec64((const cplx*)H_d + (size_t)c0 * N * N, N, cnt, evals_d, (cplx*)evecs_d, nm, c0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
            if (rc < 0) return fail("vec64 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        }
        return 0;
    }
    if (N <= 64 && !evecs_d && !heev_generic_only) {      // CTA-resident kernel, eigenvalues only
        if (p->ws[2].S.ensure((size_t)nm * small::cta64_bytes_per_k())) return fail("out of device memory (heev workspace)");
        int rc = small::launch_cta64((const cplx*)H_d, N, nm, (double*)p->ws[2].S.p, evals_d, nm, 0, p->d_fail, st, &g_kinfo);
        if (rc < 0) return fail("cta64 launch failed: %s", g_kinfo.c_str());
        g_launches += rc;
        return 0;
    }
    long long chunk = std::max<long long>(1, std::min<long long>(nm, (long long)(WS_BUDGET / perH)));
    const bool bigvec = evecs_d && bigvec_enabled(N) && !heev_generic_only;
    if (bigvec) chunk = std::max<long long>(1, std::min<long long>(nm, (long long)(BIGVEC_BUDGET / (perH + bigvec_bytes_per_matrix(N)))));
    if (p->ws[2].H.ensure(perH * chunk) || (evecs_d && p->ws[2].RT.ensure(bigvec ? bigvec_bytes_per_matrix(N) * chunk : perH * chunk)))
        return fail("out of device memory (heev workspace)");
    const bool peel = !evecs_d && evals_ws_bytes_per_k(N) > 0 && !heev_generic_only;
    if (peel && use_twostage(N)) {
        chunk = twostage_chunk(N, nm, perH);
        if (p->ws[2].H.ensure(perH * chunk)) return fail("out of device memory (heev workspace)");
    }
    if (peel && p->ws[2].S.ensure(evals_ws_bytes_per_k(N) * chunk)) return fail("out of device memory (peel workspace)");
    for (long long c0 = 0; c0 < nm; c0 += chunk) {
        long long cnt = std::min(chunk, (long long)nm - c0);
        CK(cudaMemcpyAsync(p->ws[2].H.p, (const cplx*)H_d + (size_t)c0 * N * N, perH * cnt, cudaMemcpyDeviceToDevice, st));
        if (bigvec) {
            int rc = launch_bigvec(p, (cplx*)p->ws[2].H.p, N, cnt, evals_d, nm, c0, (cplx*)evecs_d, (long long)nm * N, N, c0, 0, nullptr, p->ws[2].RT.p, st);
            if (rc) return rc;
            continue;
        }
        HeevArgs a{};
        a.H = (cplx*)p->ws[2].H.p; a.N = N; a.nm = (int)cnt;
        a.evals = evals_d; a.ld = nm; a.col0 = c0;
        if (peel) a.peel_ws = p->ws[2].S.p;
        if (evecs_d) {
            a.RT = (cplx*)p->ws[2].RT.p; a.n_rows = N; a.rows = nullptr;
            a.vec_out = (cplx*)evecs_d; a.vs_band = (long long)nm * N; a.vs_k = N; a.vcol0 = c0;
        }
        int rc = launch_heev_generic(p, a, st);
        if (rc) return rc;
    }
    return 0;
}

// Diagnostic / test entry of the two-stage tridiagonalisation (heev_twostage.cuh): band_d (NULL or [nm][N][32] complex) receives
// the band matrix after stage 1, AB[j][d] = A[j+d][j], d <= 16; de_d (NULL or [nm][2][N]) the tridiagonal after stage 2.
extern "C" int sktb_heev_twostage(sktb_plan* p, const double* H_d, int32_t N, int64_t nm, double* band_d, double* de_d, void* stream) {
    if (!p || !H_d) return fail("null argument");
    if (!twostage::supported(N)) return fail("two-stage tridiagonalisation: N = %d is not supported (32 <= N, panel must fit shared memory)", N);
    if (nm <= 0) return 0;
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t perH = (size_t)N * N * sizeof(cplx), perS = twostage::scratch_bytes_per_matrix(N);
    const long long chunk = std::max<long long>(1, std::min<long long>(nm, (long long)(WS_BUDGET / perH)));
    if (p->ws[2].H.ensure(perH * chunk) || p->ws[2].S.ensure(perS * chunk)) return fail("out of device memory (two-stage workspace)");
    for (long long c0 = 0; c0 < nm; c0 += chunk) {
        const long long cnt = std::min(chunk, (long long)nm - c0);
        CK(cudaMemcpyAsync(p->ws[2].H.p, (const cplx*)H_d + (size_t)c0 * N * N, perH * cnt, cudaMemcpyDeviceToDevice, st));
        unsigned char* ws = reinterpret_cast<unsigned char*>(p->ws[2].S.p);
        double* de = reinterpret_cast<double*>(ws + (size_t)cnt * (perS - (size_t)2 * N * sizeof(double)));
        char info[320];
        if (twostage::launch_tridiag((cplx*)p->ws[2].H.p, N, cnt, ws, de, st, info, sizeof info, 1) < 0) return fail("two-stage stage 1 launch failed");
        LAUNCHED();
        if (band_d) CK(cudaMemcpyAsync((cplx*)band_d + (size_t)c0 * N * twostage::ABW, twostage::band_of(ws), (size_t)cnt * N * twostage::ABW * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
        if (de_d) {
            if (twostage::launch_tridiag((cplx*)p->ws[2].H.p, N, cnt, ws, de, st, info, sizeof info, 2) < 0) return fail("two-stage stage 2 launch failed");
            LAUNCHED();
            CK(cudaMemcpyAsync(de_d + (size_t)c0 * 2 * N, de, (size_t)cnt * 2 * N * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        g_kinfo = info;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// host-buffer front end: two streams, chunked H2D(k) -> K1+K2 -> D2H(evals[, evecs]) overlapped
// ------------------------------------------------------------------------------------------------------
extern "C" int sktb_solve_host(sktb_plan* p, const double* k_h, int64_t nk, int32_t soc, double* evals_h, double* evecs_h,
                               int32_t* nonherm_any) {
    if (!p || !k_h || !evals_h) return fail("null argument");
    ON_DEVICE(p->device);
    if (nonherm_any) *nonherm_any = 0;
    if (nk <= 0) return 0;
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    for (int s = 0; s < 2; ++s)
        if (!p->hs[s]) CK(cudaStreamCreateWithFlags(&p->hs[s], cudaStreamNonBlocking));
    // chunk: bounded device staging (<= 512 MB of eigenvalues, <= 1 GB of eigenvectors per stream)
    static const long long host_chunk_mb = getenv("SKTB_HOST_CHUNK_MB") ? atoll(getenv("SKTB_HOST_CHUNK_MB")) : 512;
    long long chunk = std::min<long long>(nk, std::max<long long>(1, (host_chunk_mb << 20) / ((long long)N * 8)));
    if (evecs_h) chunk = std::min<long long>(chunk, std::max<long long>(1, ((long long)1 << 30) / ((long long)N * N * 16)));
    if (nk > 4 * 148 && chunk > (nk + 3) / 4) chunk = (nk + 3) / 4;      // at least 4 chunks so the copies overlap
    const bool gate = p->asym_bound > 0.5e-9;
    CK(cudaMemsetAsync(p->d_any, 0, sizeof(int), p->hs[0]));
    CK(cudaStreamSynchronize(p->hs[0]));
    long long c_index = 0;
    for (long long c0 = 0; c0 < nk; c0 += chunk, ++c_index) {
        int s = (int)(c_index & 1);
        cudaStream_t st = p->hs[s];
        long long cnt = std::min(chunk, nk - c0);
        if (p->hK[s].ensure((size_t)chunk * 24) || p->hE[s].ensure((size_t)chunk * N * 8) ||
            (evecs_h && p->hV[s].ensure((size_t)chunk * N * N * 16)) || (gate && p->hF[s].ensure((size_t)chunk * 4)))
            return fail("out of device memory (host front-end staging)");
        CK(cudaMemcpyAsync(p->hK[s].p, k_h + 3 * c0, (size_t)cnt * 24, cudaMemcpyHostToDevice, st));
        int rc = solve_impl(p, (const double*)p->hK[s].p, nullptr, 0, cnt, soc, (double*)p->hE[s].p, evecs_h ? (double*)p->hV[s].p : nullptr,
                            cnt, 0, gate ? (int*)p->hF[s].p : nullptr, st, false, s);
        if (rc) return rc;
        if (gate) { any_flag_kernel<<<64, 256, 0, st>>>((const int*)p->hF[s].p, cnt, p->d_any); LAUNCHED(); }
        const size_t max_pitch = (size_t)1 << 30;     // stay well inside cudaDeviceProp::memPitch
        if ((size_t)nk * 8 <= max_pitch) {
            CK(cudaMemcpy2DAsync(evals_h + c0, (size_t)nk * 8, p->hE[s].p, (size_t)cnt * 8, (size_t)cnt * 8, N, cudaMemcpyDeviceToHost, st));
        } else {
            for (int b = 0; b < N; ++b)
                CK(cudaMemcpyAsync(evals_h + (size_t)b * nk + c0, (const double*)p->hE[s].p + (size_t)b * cnt, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        }
        if (evecs_h) {   // device [N][cnt][N] -> host [N][nk][N]
            if ((size_t)nk * N * 16 <= max_pitch) {
                CK(cudaMemcpy2DAsync(evecs_h + (size_t)2 * c0 * N, (size_t)nk * N * 16, p->hV[s].p, (size_t)cnt * N * 16, (size_t)cnt * N * 16, N,
                                     cudaMemcpyDeviceToHost, st));
            } else {
                for (int b = 0; b < N; ++b)
                    CK(cudaMemcpyAsync(evecs_h + ((size_t)b * nk + c0) * N * 2, (const double*)p->hV[s].p + (size_t)b * cnt * N * 2, (size_t)cnt * N * 16,
                                       cudaMemcpyDeviceToHost, st));
            }
        }
    }
    CK(cudaMemcpyAsync(p->h_pinned, p->d_any, sizeof(int), cudaMemcpyDeviceToHost, p->hs[0]));
    CK(cudaStreamSynchronize(p->hs[1]));
    CK(cudaStreamSynchronize(p->hs[0]));
    // d_any may have been set by the other stream after the copy was queued: re-read once everything is done
    CK(cudaMemcpy(p->h_pinned, p->d_any, sizeof(int), cudaMemcpyDeviceToHost));
    if (nonherm_any) *nonherm_any = p->h_pinned[0];
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// K1+K2+K3: DOS / LDOS, and A(k,E) maps
// ------------------------------------------------------------------------------------------------------
static int tracked_rows(sktb_plan* p, int N, int n_groups, const int32_t* grp_ptr, const int32_t* grp_idx,
                        std::vector<int>* rows, std::vector<int>* grp_rows) {
    rows->clear(); grp_rows->clear();
    int total = n_groups > 0 ? grp_ptr[n_groups] : 0;
    for (int q = 0; q < total; ++q) {
        if (grp_idx[q] < 0 || grp_idx[q] >= N) return fail("group index %d outside the %d x %d Hamiltonian", grp_idx[q], N, N);
        rows->push_back(grp_idx[q]);
    }
    std::sort(rows->begin(), rows->end());
    rows->erase(std::unique(rows->begin(), rows->end()), rows->end());
    for (int q = 0; q < total; ++q)
        grp_rows->push_back((int)(std::lower_bound(rows->begin(), rows->end(), grp_idx[q]) - rows->begin()));
    (void)p;
    return 0;
}

extern "C" int sktb_dos(sktb_plan* p, const double* k_d, const int32_t nk_mesh[3], int64_t k_first, int64_t k_count, int32_t soc,
                        const double* energies_d, int32_t nE, double eta, int32_t n_groups, const int32_t* grp_ptr,
                        const int32_t* grp_idx, double* out_d, int32_t* nonherm_any_h, void* stream) {
    if (!p || !energies_d || !out_d) return fail("null argument");
    if (!k_d && !nk_mesh) return fail("need k points or a mesh");
    if (nE <= 0) return 0;
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    const int ng = n_groups > 0 ? n_groups : 1;
    std::vector<int> rows, grp_rows;
    if (tracked_rows(p, N, n_groups, grp_ptr, grp_idx, &rows, &grp_rows)) return 1;
    int n_rows = (int)rows.size();
    const bool gate = p->asym_bound > 0.5e-9;
    // N <= 32 with projections: full eigenvectors from the register path (vec32), weights read by row index
    const bool use_vec32 = n_rows > 0 && p->small_ok && small::supports(N) && !gate;
    const bool use_vec64 = n_rows > 0 && N > 32 && N <= 64;        // full eigenvectors from tridiag_cta64 + vec64
    if (use_vec32 || use_vec64) { grp_rows.assign(grp_idx, grp_idx + grp_ptr[n_groups]); n_rows = N; }
    CK(cudaMemsetAsync(out_d, 0, (size_t)ng * nE * 8, st));
    CK(cudaMemsetAsync(p->d_any, 0, sizeof(int), st));
    if (k_count <= 0) { if (nonherm_any_h) *nonherm_any_h = 0; return 0; }

    // device copies of the group tables
    int *d_rows = nullptr, *d_gptr = nullptr, *d_grows = nullptr;
    if (n_groups > 0) {
        size_t need = (rows.size() + (size_t)n_groups + 1 + grp_rows.size()) * sizeof(int);
        if (p->wsMisc.ensure(need)) return fail("out of device memory");
        d_rows = (int*)p->wsMisc.p; d_gptr = d_rows + rows.size(); d_grows = d_gptr + n_groups + 1;
        CK(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_gptr, grp_ptr, ((size_t)n_groups + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_grows, grp_rows.data(), grp_rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // host vectors go out of scope at return; keep it simple
    }
    const bool use_small = n_rows == 0 && p->small_ok && small::supports(N);
    const bool use_cta64 = n_rows == 0 && !use_small && N > 32 && N <= 64;     // eigenvalues only, register-resident CTA kernel
    const bool dos_fused = (use_vec64 || use_cta64) && c64_fusable(p, N, soc, gate);
    const bool dos_bigvec = n_rows > 0 && !use_vec32 && !use_vec64 && bigvec_enabled(N);
    size_t per_k = (size_t)N * 8 + ((use_small || use_vec32 || dos_fused) ? 0 : (size_t)N * N * 16) + (size_t)N * n_rows * 32 + 24 + 4 + (use_cta64 ? small::cta64_bytes_per_k() : 0) +
                   (use_vec32 ? small::vec32_bytes_per_k() : 0) + (use_vec64 ? small::vec64_bytes_per_k() : 0) + (dos_bigvec ? bigvec_bytes_per_matrix(N) : 0);
    long long chunk = std::max<long long>(1, std::min<long long>(k_count, (long long)((dos_bigvec ? BIGVEC_BUDGET : WS_BUDGET) / per_k)));
    if (dos_bigvec && bigvec_twostage(N, n_rows) && chunk >= 592 && chunk < k_count) chunk = chunk / 592 * 592;   // whole waves of stage 1 (148) and of the bulge chase (4 x 148)
    const int LOR_GRID = 148 * 4;
    if (p->wsE.ensure((size_t)chunk * N * 8) || (!use_small && !use_vec32 && !dos_fused && p->ws[2].H.ensure((size_t)chunk * N * N * 16)) ||
        (n_rows && (p->ws[2].RT.ensure(use_vec32 ? (size_t)chunk * small::vec32_bytes_per_k()
                                                 : (use_vec64 ? (size_t)chunk * small::vec64_bytes_per_k()
                                                              : (dos_bigvec ? (size_t)chunk * bigvec_bytes_per_matrix(N) : (size_t)chunk * N * n_rows * 16))) ||
                    p->wsVec.ensure((size_t)chunk * N * n_rows * 16))) ||
        (!k_d && !use_small && !use_vec32 && p->ws[2].K.ensure((size_t)chunk * 24)) || (gate && p->wsFlag.ensure((size_t)chunk * 4)) ||
        p->wsPart.ensure((size_t)LOR_GRID * ng * nE * 8))
        return fail("out of device memory (dos workspaces)");

    for (long long c0 = 0; c0 < k_count; c0 += chunk) {
        long long cnt = std::min(chunk, (long long)k_count - c0);
        int* flags = gate ? (int*)p->wsFlag.p : nullptr;
        if (use_vec32) {
            int rc = small::launch_vec32(p->small_tab, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, k_first + c0, cnt, soc, N, (double*)p->wsE.p,
                                         (cplx*)p->wsVec.p, cnt, 0, flags, 0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
            if (rc < 0) return fail("vec32 launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        } else if (use_small) {
            size_t sbytes = small_scratch_bytes(N, cnt);
            if (p->ws[2].S.ensure(sbytes)) return fail("out of device memory (tridiagonal scratch)");
            int rc = small::launch_solve(p->small_tab, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, k_first + c0, cnt, soc, N, (double*)p->wsE.p, cnt, 0,
                                         flags, 0, p->d_fail, (double*)p->ws[2].S.p, sbytes, st, &g_kinfo);
            if (rc < 0) return fail("small-path launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        } else if ((use_vec64 || use_cta64) && c64_fusable(p, N, soc, gate)) {
            const small::C64Build b = c64_build(p, k_d ? k_d + 3 * c0 : nullptr, nk_mesh, k_first + c0, soc);
            int rl;
            if (use_vec64) rl = small::launch_vec64(nullptr, N, cnt, (double*)p->wsE.p, (cplx*)p->wsVec.p, cnt, 0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo, &b);
            else {
                if (p->ws[2].S.ensure((size_t)cnt * small::cta64_bytes_per_k())) return fail("out of device memory (tridiagonal scratch)");
                rl = small::launch_cta64(nullptr, N, cnt, (double*)p->ws[2].S.p, (double*)p->wsE.p, cnt, 0, p->d_fail, st, &g_kinfo, 0, 0, 0, false, &b);
            }
            if (rl < 0) return fail("fused cta64 launch failed: %s", g_kinfo.c_str());
            g_launches += rl;
        } else {
            const double* kk = k_d ? k_d + 3 * c0 : (const double*)p->ws[2].K.p;
            if (!k_d) { int rc = sktb_kmesh(nk_mesh, k_first + c0, cnt, (double*)p->ws[2].K.p, st); if (rc) return rc; }
            int rc = launch_hk(p, kk, cnt, soc, (cplx*)p->ws[2].H.p, flags, 0, st);
            if (rc) return rc;
            HeevArgs a{};
            a.H = (cplx*)p->ws[2].H.p; a.N = N; a.nm = (int)cnt;
            a.evals = (double*)p->wsE.p; a.ld = cnt; a.col0 = 0;
            if (n_rows) {
                a.RT = (cplx*)p->ws[2].RT.p; a.n_rows = n_rows; a.rows = d_rows;
                a.vec_out = (cplx*)p->wsVec.p; a.vs_band = cnt * n_rows; a.vs_k = n_rows; a.vcol0 = 0;
            }
            if (dos_bigvec) {
                void* ts_ws = nullptr;
                if (bigvec_twostage(N, n_rows)) {
                    if (p->ws[2].S.ensure((size_t)cnt * twostage::scratch_bytes_per_matrix(N))) return fail("out of device memory (two-stage workspace)");
                    ts_ws = p->ws[2].S.p;
                }
                rc = launch_bigvec(p, (cplx*)p->ws[2].H.p, N, cnt, (double*)p->wsE.p, cnt, 0, (cplx*)p->wsVec.p, cnt * n_rows, n_rows, 0, n_rows, d_rows,
                                   p->ws[2].RT.p, st, ts_ws);
                if (rc) return rc;
            } else if (use_vec64) {
                int rl = small::launch_vec64((const cplx*)p->ws[2].H.p, N, cnt, (double*)p->wsE.p, (cplx*)p->wsVec.p, cnt, 0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
                if (rl < 0) return fail("vec64 launch failed: %s", g_kinfo.c_str());
                g_launches += rl;
            } else if (use_cta64) {
                if (p->ws[2].S.ensure((size_t)cnt * small::cta64_bytes_per_k())) return fail("out of device memory (tridiagonal scratch)");
                int rl = small::launch_cta64((const cplx*)p->ws[2].H.p, N, cnt, (double*)p->ws[2].S.p, (double*)p->wsE.p, cnt, 0, p->d_fail, st, &g_kinfo);
                if (rl < 0) return fail("cta64 launch failed: %s", g_kinfo.c_str());
                g_launches += rl;
            } else {
                if (!n_rows && evals_ws_bytes_per_k(N) > 0) {
                    if (p->ws[2].S.ensure((size_t)cnt * evals_ws_bytes_per_k(N))) return fail("out of device memory (peel workspace)");
                    a.peel_ws = p->ws[2].S.p;
                }
                rc = launch_heev_generic(p, a, st);
                if (rc) return rc;
            }
        }
        if (gate) { any_flag_kernel<<<64, 256, 0, st>>>(flags, cnt, p->d_any); LAUNCHED(); }
        long long tiles = (cnt * N + SKTB_LOR_TILE - 1) / SKTB_LOR_TILE;
        int gx = (int)std::min<long long>(tiles, LOR_GRID);
        const int ept = nE <= 128 ? 1 : (nE <= 256 ? 2 : 4);
        for (int e_base = 0; e_base < nE; e_base += 128 * ept) {
            auto kl = ept == 1 ? lorentz_partial_kernel<1> : (ept == 2 ? lorentz_partial_kernel<2> : lorentz_partial_kernel<4>);
            kl<<<dim3(gx, ng), 128, 0, st>>>((const double*)p->wsE.p, cnt, 0, cnt, N, (const cplx*)p->wsVec.p, n_rows, d_gptr, d_grows, n_groups,
                                             energies_d, nE, e_base, eta, (double*)p->wsPart.p);
            LAUNCHED();
        }
        lorentz_reduce_kernel<<<(ng * nE + 127) / 128, 128, 0, st>>>((const double*)p->wsPart.p, gx, ng * nE, out_d, 1);
        LAUNCHED();
    }
    CK(cudaMemcpyAsync(p->h_pinned, p->d_any, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nonherm_any_h) *nonherm_any_h = p->h_pinned[0];
    return 0;
}

extern "C" int sktb_spectral(sktb_plan* p, const double* k_d, int64_t nk, int32_t soc, const double* energies_d, int32_t nE, double eta,
                             int32_t n_idx, const int32_t* idx, double* out_d, int32_t* nonherm_any_h, void* stream) {
    if (!p || !k_d || !energies_d || !out_d) return fail("null argument");
    if (nk <= 0 || nE <= 0) return 0;
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    std::vector<int> rows, grp_rows;
    int gp[2] = {0, n_idx};
    if (tracked_rows(p, N, n_idx > 0 ? 1 : 0, gp, idx, &rows, &grp_rows)) return 1;
    const int n_rows = (int)rows.size();
    // duplicated indices count twice in the reference's loop (greens.py:462-465): weight rows by multiplicity
    // by tracking each occurrence separately when duplicates exist
    if ((int)grp_rows.size() != n_rows) { rows.assign(idx, idx + n_idx); }
    const int nr = (int)rows.size();
    int* d_rows = nullptr;
    if (nr) {
        if (p->wsMisc.ensure(rows.size() * sizeof(int))) return fail("out of device memory");
        d_rows = (int*)p->wsMisc.p;
        CK(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    const bool gate = p->asym_bound > 0.5e-9;
    CK(cudaMemsetAsync(p->d_any, 0, sizeof(int), st));
    // solver choice: register-resident paths (full eigenvectors when a projection is asked for), else generic row tracking
    const bool small_path = p->small_ok && small::supports(N) && !gate;
    const bool mid_path = N > 32 && N <= 64;
    const bool full_vec = nr > 0 && (small_path || mid_path);
    const int vstride = full_vec ? N : nr;
    const bool sp_bigvec = nr > 0 && !small_path && !mid_path && bigvec_enabled(N);
    size_t per_k = (size_t)N * 8 + ((small_path) ? 0 : (size_t)N * N * 16) + (size_t)N * vstride * 32 + 4 +
                   (small_path ? (nr ? small::vec32_bytes_per_k() : 1024) : 0) + (mid_path ? small::vec64_bytes_per_k() : 0) + (sp_bigvec ? bigvec_bytes_per_matrix(N) : 0);
    long long chunk = std::max<long long>(1, std::min<long long>(nk, (long long)((sp_bigvec ? BIGVEC_BUDGET : WS_BUDGET) / per_k)));
    if (sp_bigvec && bigvec_twostage(N, nr) && chunk >= 592 && chunk < nk) chunk = chunk / 592 * 592;
    if (p->wsE.ensure((size_t)chunk * N * 8) || (!small_path && p->ws[2].H.ensure((size_t)chunk * N * N * 16)) ||
        (nr && p->wsVec.ensure((size_t)chunk * N * vstride * 16)) ||
        (nr && p->ws[2].RT.ensure(small_path ? (size_t)chunk * small::vec32_bytes_per_k()
                                             : (mid_path ? (size_t)chunk * small::vec64_bytes_per_k()
                                                         : (sp_bigvec ? (size_t)chunk * bigvec_bytes_per_matrix(N) : (size_t)chunk * N * nr * 16)))) ||
        (gate && p->wsFlag.ensure((size_t)chunk * 4)))
        return fail("out of device memory (spectral workspaces)");
    for (long long c0 = 0; c0 < nk; c0 += chunk) {
        long long cnt = std::min(chunk, (long long)nk - c0);
        int* flags = gate ? (int*)p->wsFlag.p : nullptr;
        if (small_path) {
            int rc;
            if (nr) rc = small::launch_vec32(p->small_tab, k_d + 3 * c0, nullptr, 0, cnt, soc, N, (double*)p->wsE.p, (cplx*)p->wsVec.p, cnt, 0, flags, 0,
                                             p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
            else {
                size_t sbytes = small_scratch_bytes(N, cnt);
                if (p->ws[2].S.ensure(sbytes)) return fail("out of device memory (tridiagonal scratch)");
                rc = small::launch_solve(p->small_tab, k_d + 3 * c0, nullptr, 0, cnt, soc, N, (double*)p->wsE.p, cnt, 0, flags, 0, p->d_fail,
                                         (double*)p->ws[2].S.p, sbytes, st, &g_kinfo);
            }
            if (rc < 0) return fail("small-path launch failed: %s", g_kinfo.c_str());
            g_launches += rc;
        } else if (mid_path && c64_fusable(p, N, soc, gate)) {
            const small::C64Build b = c64_build(p, k_d + 3 * c0, nullptr, 0, soc);
            int rl;
            if (nr) rl = small::launch_vec64(nullptr, N, cnt, (double*)p->wsE.p, (cplx*)p->wsVec.p, cnt, 0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo, &b);
            else {
                if (p->ws[2].S.ensure((size_t)cnt * small::cta64_bytes_per_k())) return fail("out of device memory (tridiagonal scratch)");
                rl = small::launch_cta64(nullptr, N, cnt, (double*)p->ws[2].S.p, (double*)p->wsE.p, cnt, 0, p->d_fail, st, &g_kinfo, 0, 0, 0, false, &b);
            }
            if (rl < 0) return fail("fused mid-path launch failed: %s", g_kinfo.c_str());
            g_launches += rl;
        } else {
            int rc = launch_hk(p, k_d + 3 * c0, cnt, soc, (cplx*)p->ws[2].H.p, flags, 0, st);
            if (rc) return rc;
            if (mid_path) {
                int rl;
                if (nr) rl = small::launch_vec64((const cplx*)p->ws[2].H.p, N, cnt, (double*)p->wsE.p, (cplx*)p->wsVec.p, cnt, 0, p->d_fail, p->ws[2].RT.p, st, &g_kinfo);
                else {
                    if (p->ws[2].S.ensure((size_t)cnt * small::cta64_bytes_per_k())) return fail("out of device memory (tridiagonal scratch)");
                    rl = small::launch_cta64((const cplx*)p->ws[2].H.p, N, cnt, (double*)p->ws[2].S.p, (double*)p->wsE.p, cnt, 0, p->d_fail, st, &g_kinfo);
                }
                if (rl < 0) return fail("mid-path launch failed: %s", g_kinfo.c_str());
                g_launches += rl;
            } else {
                HeevArgs a{};
                a.H = (cplx*)p->ws[2].H.p; a.N = N; a.nm = (int)cnt;
                a.evals = (double*)p->wsE.p; a.ld = cnt; a.col0 = 0;
                if (sp_bigvec) {
                    void* ts_ws = nullptr;
                    if (bigvec_twostage(N, nr)) {
                        if (p->ws[2].S.ensure((size_t)cnt * twostage::scratch_bytes_per_matrix(N))) return fail("out of device memory (two-stage workspace)");
                        ts_ws = p->ws[2].S.p;
                    }
                    rc = launch_bigvec(p, (cplx*)p->ws[2].H.p, N, cnt, (double*)p->wsE.p, cnt, 0, (cplx*)p->wsVec.p, cnt * nr, nr, 0, nr, d_rows, p->ws[2].RT.p, st, ts_ws);
                    if (rc) return rc;
                } else {
                if (nr) {
                    a.RT = (cplx*)p->ws[2].RT.p; a.n_rows = nr; a.rows = d_rows;
                    a.vec_out = (cplx*)p->wsVec.p; a.vs_band = cnt * nr; a.vs_k = nr; a.vcol0 = 0;
                } else if (evals_ws_bytes_per_k(N) > 0) {
                    if (p->ws[2].S.ensure((size_t)cnt * evals_ws_bytes_per_k(N))) return fail("out of device memory (peel workspace)");
                    a.peel_ws = p->ws[2].S.p;
                }
                rc = launch_heev_generic(p, a, st);
                if (rc) return rc;
                }
            }
        }
        if (gate) { any_flag_kernel<<<64, 256, 0, st>>>(flags, cnt, p->d_any); LAUNCHED(); }
        unsigned grid = (unsigned)std::min<long long>(cnt, 148LL * 8);
        spectral_map_kernel<<<grid, 128, (size_t)2 * N * 8, st>>>((const double*)p->wsE.p, cnt, 0, cnt, N, (const cplx*)p->wsVec.p, vstride,
                                                                full_vec ? d_rows : nullptr, nr, energies_d, nE, eta, out_d, c0);
        LAUNCHED();
    }
    CK(cudaMemcpyAsync(p->h_pinned, p->d_any, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nonherm_any_h) *nonherm_any_h = p->h_pinned[0];
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// K1+K5: the reference's dense-inverse Green's function for models whose H(k) is not Hermitian (gf_inverse.cuh)
// ------------------------------------------------------------------------------------------------------
static int launch_gf(sktb_plan* p, GfArgs a, cudaStream_t st) {
    const int N = a.N;
    if (N > 2048) return fail("dense-inverse Green's function: N = %d > 2048", N);
    static const bool no_warp = getenv("SKTB_GF_WARP") && atoi(getenv("SKTB_GF_WARP")) == 0;
    static const bool no_tile = getenv("SKTB_GF_TILE") && atoi(getenv("SKTB_GF_TILE")) == 0;
    static const bool tile1 = getenv("SKTB_GF_TILE1") && atoi(getenv("SKTB_GF_TILE1")) == 1;
    if (N <= 64 && (N > 32 || tile1) && !a.Gout && a.part && !no_tile) {   // 8 x 4 register tiles per lane, four warps per matrix (one for N <= 32: knob)
        const long long items = a.nk * a.nE;
        const int mpc = N <= 32 ? 4 : 1;
        const unsigned grid = (unsigned)std::min<long long>((items + mpc - 1) / mpc, 148LL * SKTB_GF_TILE_MINB);
        if (N <= 32) gf_inverse_tile_kernel<1><<<grid, 128, 0, st>>>(a);
        else gf_inverse_tile_kernel<4><<<grid, 128, 0, st>>>(a);
        LAUNCHED();
        char info[200];
        snprintf(info, sizeof info, "gf_inverse_tile<%d> (dense inverse per (k,E), N=%d, 8x4 register tile per lane, %s) grid=%u x 128 thr 2 CTA/SM", N <= 32 ? 1 : 4,
                 N, N <= 32 ? "one warp per matrix" : "four warps per matrix", grid);
        g_kinfo = info;
        return 0;
    }
    if (N <= 32 && !a.Gout && a.part && !no_warp) {        // one matrix per warp, rows in registers
        const long long items = a.nk * a.nE;
        const unsigned grid = (unsigned)std::min<long long>((items + GFW_WARPS - 1) / GFW_WARPS, 148LL * 3);
        gf_inverse_warp_kernel<<<grid, GFW_WARPS * 32, 0, st>>>(a);
        LAUNCHED();
        char info[200];
        snprintf(info, sizeof info, "gf_inverse_warp (dense inverse per (k,E), N=%d, one matrix per warp in registers) grid=%u x %d thr 3 CTA/SM", N, grid,
                 GFW_WARPS * 32);
        g_kinfo = info;
        return 0;
    }
    if (N > 32 && N <= 96 && !a.Gout && a.part && !no_warp) {   // one matrix per CTA, distributed over its registers
        const long long items = a.nk * a.nE;
        const unsigned grid = (unsigned)std::min<long long>(items, 148LL * 2);
        if (N <= 64) gf_inverse_cta_kernel<8, 2><<<grid, GF_THREADS, 0, st>>>(a);
        else gf_inverse_cta_kernel<12, 3><<<grid, GF_THREADS, 0, st>>>(a);
        LAUNCHED();
        char info[200];
        snprintf(info, sizeof info, "gf_inverse_cta<%s> (dense inverse per (k,E), N=%d, one matrix per CTA in registers) grid=%u x %d thr", N <= 64 ? "8,2" : "12,3", N,
                 grid, GF_THREADS);
        g_kinfo = info;
        return 0;
    }
    const bool in_smem = gf_smem_bytes(N, true) <= 227 * 1024;
    const size_t smem = gf_smem_bytes(N, in_smem);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gf_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every call
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gf_inverse_kernel, GF_THREADS, smem));
    per_sm = std::max(per_sm, 1);
    const long long items = a.nk * a.nE;
    const unsigned grid = (unsigned)std::min<long long>(items, 148LL * per_sm);
    if (!in_smem) {
        if (p->ws[2].RT.ensure((size_t)grid * N * (N + 1) * sizeof(cplx))) return fail("out of device memory (inverse workspace)");
        a.ws = (cplx*)p->ws[2].RT.p;
    }
    gf_inverse_kernel<<<grid, GF_THREADS, smem, st>>>(a);
    LAUNCHED();
    char info[200];
    snprintf(info, sizeof info, "gf_inverse (dense inverse per (k,E), N=%d, %s) grid=%u x %d thr smem=%zuB %d CTA/SM", N,
             in_smem ? "matrix in shared memory" : "matrix in a global workspace", grid, GF_THREADS, smem, per_sm);
    g_kinfo = info;
    return 0;
}

extern "C" int sktb_gf_dos(sktb_plan* p, const double* k_d, const int32_t nk_mesh[3], int64_t k_first, int64_t k_count, int32_t soc,
                           const double* energies_d, int32_t nE, double eta, int32_t n_groups, const int32_t* grp_ptr,
                           const int32_t* grp_idx, int32_t per_k, double* out_d, void* stream) {
    if (!p || !energies_d || !out_d) return fail("null argument");
    if (!k_d && !nk_mesh) return fail("need k points or a mesh");
    if (nE <= 0) return 0;
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    const int ng = n_groups > 0 ? n_groups : 1;
    const int total = n_groups > 0 ? grp_ptr[n_groups] : 0;
    for (int q = 0; q < total; ++q)
        if (grp_idx[q] < 0 || grp_idx[q] >= N) return fail("group index %d outside the %d x %d Hamiltonian", grp_idx[q], N, N);
    if (!per_k) CK(cudaMemsetAsync(out_d, 0, (size_t)ng * nE * 8, st));
    if (k_count <= 0) return 0;
    int *d_gptr = nullptr, *d_grows = nullptr;
    if (n_groups > 0) {
        if (p->wsMisc.ensure(((size_t)n_groups + 1 + std::max(total, 1)) * sizeof(int))) return fail("out of device memory");
        d_gptr = (int*)p->wsMisc.p; d_grows = d_gptr + n_groups + 1;
        CK(cudaMemcpyAsync(d_gptr, grp_ptr, ((size_t)n_groups + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        if (total) CK(cudaMemcpyAsync(d_grows, grp_idx, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    const size_t per_kpt = (size_t)N * N * 16 + 24 + (per_k ? 0 : (size_t)ng * nE * 8);
    const long long chunk = std::max<long long>(1, std::min<long long>(k_count, (long long)(((size_t)1 << 30) / per_kpt)));
    if (p->ws[2].H.ensure((size_t)chunk * N * N * 16) || (!k_d && p->ws[2].K.ensure((size_t)chunk * 24)) ||
        (!per_k && p->wsPart.ensure((size_t)chunk * ng * nE * 8)))
        return fail("out of device memory (Green's function workspaces)");
    for (long long c0 = 0; c0 < k_count; c0 += chunk) {
        const long long cnt = std::min(chunk, (long long)k_count - c0);
        const double* kk = k_d ? k_d + 3 * c0 : (const double*)p->ws[2].K.p;
        if (!k_d) { int rc = sktb_kmesh(nk_mesh, k_first + c0, cnt, (double*)p->ws[2].K.p, st); if (rc) return rc; }
        int rc = launch_hk(p, kk, cnt, soc, (cplx*)p->ws[2].H.p, nullptr, 0, st);
        if (rc) return rc;
        GfArgs a{};
        a.H = (const cplx*)p->ws[2].H.p; a.N = N; a.nk = cnt; a.E = energies_d; a.nE = nE; a.eta = eta;
        a.n_groups = n_groups > 0 ? n_groups : 0; a.gptr = d_gptr; a.grows = d_grows;
        a.part = per_k ? out_d + (size_t)c0 * ng * nE : (double*)p->wsPart.p;
        rc = launch_gf(p, a, st);
        if (rc) return rc;
        if (!per_k) {
            lorentz_reduce_kernel<<<(ng * nE + 127) / 128, 128, 0, st>>>((const double*)p->wsPart.p, (int)cnt, ng * nE, out_d, 1);
            LAUNCHED();
        }
    }
    return 0;
}

extern "C" int sktb_gf_inverse(sktb_plan* p, const double* H_d, int32_t N, int64_t nm, const double* energies_d, int32_t nE, double eta,
                               double* G_d, double* trace_d, void* stream) {
    if (!p || !H_d || !energies_d || (!G_d && !trace_d)) return fail("null argument");
    if (N <= 0 || nm <= 0 || nE <= 0) return 0;
    ON_DEVICE(p->device);
    GfArgs a{};
    a.H = reinterpret_cast<const cplx*>(H_d); a.N = N; a.nk = nm; a.E = energies_d; a.nE = nE; a.eta = eta;
    a.Gout = reinterpret_cast<cplx*>(G_d); a.part = trace_d;
    return launch_gf(p, a, (cudaStream_t)stream);
}

extern "C" int sktb_gf_retarded(sktb_plan* p, const double k_h[3], double E, double eta, int32_t soc, double* G_d, void* stream) {
    if (!p || !k_h || !G_d) return fail("null argument");
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = soc ? 2 * p->n_orb : p->n_orb;
    if (p->ws[2].H.ensure((size_t)N * N * 16) || p->ws[2].K.ensure(64)) return fail("out of device memory");
    double* kE = (double*)p->ws[2].K.p;                     // k (3 doubles) | E
    const double host[4] = {k_h[0], k_h[1], k_h[2], E};
    CK(cudaMemcpyAsync(kE, host, sizeof host, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                          // `host` leaves scope at return
    int rc = launch_hk(p, kE, 1, soc, (cplx*)p->ws[2].H.p, nullptr, 0, st);
    if (rc) return rc;
    return sktb_gf_inverse(p, (const double*)p->ws[2].H.p, N, 1, kE + 3, 1, eta, G_d, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------------
// Gaussian get_dos, band energy
// ------------------------------------------------------------------------------------------------------
extern "C" int sktb_gauss_dos(sktb_plan* p, const double* evals_d, int64_t n, const double* energies_d, int32_t nE, double w, double* out_d,
                              void* stream) {
    if (!p || !energies_d || !out_d) return fail("null argument");
    if (nE <= 0) return 0;
    if (!(w > 0.0)) return fail("gaussian width must be positive");
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(out_d, 0, (size_t)nE * 8, st));
    if (n <= 0) return 0;
    if (!evals_d) return fail("null argument");
    const int LOR_GRID = 148 * 4;
    const int gx = (int)std::min<long long>((n + SKTB_LOR_TILE - 1) / SKTB_LOR_TILE, LOR_GRID);
    if (p->wsPart.ensure((size_t)LOR_GRID * nE * 8)) return fail("out of device memory (dos partials)");
    const int ept = nE <= 128 ? 1 : (nE <= 256 ? 2 : 4);
    for (int e_base = 0; e_base < nE; e_base += 128 * ept) {
        auto kg = ept == 1 ? gauss_partial_kernel<1> : (ept == 2 ? gauss_partial_kernel<2> : gauss_partial_kernel<4>);
        kg<<<gx, 128, 0, st>>>(evals_d, n, energies_d, nE, e_base, w, (double*)p->wsPart.p);
        LAUNCHED();
    }
    lorentz_reduce_kernel<<<(nE + 127) / 128, 128, 0, st>>>((const double*)p->wsPart.p, gx, nE, out_d, 0);
    LAUNCHED();
    return 0;
}

extern "C" int sktb_band_energy(sktb_plan* p, const double* evals_d, int64_t ld, int64_t nk, int32_t n_occ, double* out_d, void* stream) {
    if (!p || !out_d) return fail("null argument");
    ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(out_d, 0, 8, st));
    if (nk <= 0 || n_occ <= 0) return 0;
    if (!evals_d) return fail("null argument");
    const int grid = (int)std::min<long long>((nk + 255) / 256, 148 * 4);
    if (p->wsPart.ensure((size_t)grid * 8)) return fail("out of device memory");
    band_energy_partial_kernel<<<grid, 256, 0, st>>>(evals_d, ld, nk, n_occ, (double*)p->wsPart.p);
    LAUNCHED();
    band_energy_reduce_kernel<<<1, 32, 0, st>>>((const double*)p->wsPart.p, grid, out_d);
    LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// measurement helper
// ------------------------------------------------------------------------------------------------------
extern "C" int sktb_fp64_peak(int device, double* tflops, double* ms_out) {
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CK(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16, threads = 512, blocks = prop.multiProcessorCount * 4;
    fp64_peak_kernel<<<blocks, threads>>>(d, 1024, 1.0);           // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float t; cudaEventElapsedTime(&t, e0, e1);
        best = std::min(best, t);
    }
    g_launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}
