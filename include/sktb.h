/* sktb.h - C ABI of libsktb.so: the B200-native k-point hot path of pysktb.
 *
 * The reference (pure Python, /root/reference/pysktb) has no FFI; its only swappable operator is the Numba
 * phase loop `get_gmat_jit` (jit_modules.py:11-23) and everything around it is numpy.  This ABI is what a
 * Python shim binds with ctypes to replace, per entry point, the reference call sites cited below.
 *
 * Conventions: all `*_d` pointers are DEVICE pointers on the plan's device, caller-allocated; `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream); every call is stream-ordered and returns
 * immediately unless stated; return value 0 = ok, non-zero = error with text in sktb_last_error().
 * Complex numbers are interleaved (re, im) doubles.  A plan is bound to one device; calls on one plan are
 * not re-entrant.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef SKTB_H
#define SKTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKTB_N_V 25        /* number of V_* Slater-Koster parameters, order of sktb_v_names() */
#define SKTB_N_ORB_KINDS 17 /* s px py pz dxy dyz dxz dx2-y2 dz2 fz3 fxz2 fyz2 fz(x2-y2) fxyz fx(x2-3y2) fy(3x2-y2) S */

typedef struct sktb_plan sktb_plan;

/* k-independent model tables, HOST pointers (copied by sktb_plan_create).
 * Replaces what Hamiltonian.__init__ / calc_ham_wo_k store in H_wo_g, soc_mat, bond_mat, dist_mat_vec
 * (hamiltonian.py:64-79, 425-488). Bonds are directed (image, atom_i, atom_j) in the reference's visiting
 * order (atom_i outer, atom_j, image inner; hamiltonian.py:447-454).                                       */
typedef struct {
    int32_t n_atoms;
    int32_t n_orb;                  /* orbitals without spin (Hamiltonian.n_orbitals)                        */
    const int32_t* atom_orb_start;  /* [n_atoms+1] first orbital of each atom (System.all_iter order)        */
    const int32_t* orb_kind;        /* [n_orb] 0..16 index into the canonical orbital list (atom.py:21-39)   */
    const double* onsite;           /* [n_orb] on-site energy of each orbital (system.py:466-521)            */
    int32_t n_bonds;
    const int32_t* bond_image;      /* [n_bonds] periodic-image index (itertools.product order)              */
    const int32_t* bond_i;          /* [n_bonds] row atom    (atom_1 of calc_ham_wo_k)                       */
    const int32_t* bond_j;          /* [n_bonds] column atom (atom_2)                                        */
    const double* bond_vec;         /* [n_bonds][3] Cartesian displacement dist_mat_vec[image,i,j]           */
    const double* bond_lmn;         /* [n_bonds][3] direction cosines (structure.py:168-174)                 */
    const double* bond_V;           /* [n_bonds][25] V_* at the bond length (system.py:173-213)              */
    const double* rec_lattice;      /* [3][3] inv(lattice).T, used as k_frac @ rec (hamiltonian.py:279-280)  */
    int32_t soc_nnz;                /* COO of soc_mat (2n x 2n) in the reference's index convention          */
    const int32_t* soc_row;
    const int32_t* soc_col;
    const double* soc_val;          /* [soc_nnz][2] (re, im)                                                 */
} sktb_model_desc;

const char* sktb_last_error(void);
const char* sktb_version(void);
const char* sktb_v_names(void); /* comma separated V_* names in bond_V column order */

/* Uploads the tables, runs K4 (Slater-Koster tables, _params.py:36-513) on the device and packs the per-bond
 * hopping blocks.  Synchronous.                                                                              */
int sktb_plan_create(const sktb_model_desc* desc, int device, sktb_plan** out);
void sktb_plan_destroy(sktb_plan* plan);
/* Download the per-bond hopping blocks (host pointer): block b is [n_i][n_j] doubles at offset
 * sum_{b'<b} n_i(b') n_j(b'), bonds in the order given at creation.  Gives Hamiltonian.H_wo_g.  Synchronous. */
int sktb_plan_hoppings(const sktb_plan* plan, double* out_host, int64_t capacity);
/* Upper bound on max Re(H - H^dagger) over all k (sum over bonds of |T_b - T_rev(b)^T| + SOC asymmetry);
 * 0 means the Hermiticity gate of _sol_ham (hamiltonian.py:112) can never fire.                              */
int sktb_plan_asym_bound(const sktb_plan* plan, double* bound);

/* K4: get_hop_int for nb bonds (_params.py:36-513). V_d [nb][25], lmn_d [nb][3], out_d [nb][17][17].         */
int sktb_sk_table(const double* V_d, const double* lmn_d, double* out_d, int32_t nb, void* stream);

/* K4 with its Cartesian gradient, for Hellmann-Feynman forces (Forces.get_forces, forces.py:43-72; replaces the reference's
 * _compute_dH_pair, forces.py:242-318, which differentiates get_hop_int numerically with eps = 1e-6): forward-mode
 * differentiation of the same table code.  V_d [nb][25] values and dV_d [nb][25] derivatives dV/dd at the bond length,
 * vec_d [nb][3] Cartesian bond vectors; out_d [nb][4][17][17]: table | d/dx | d/dy | d/dz.  Synchronous.                  */
int sktb_sk_table_grad(const double* V_d, const double* dV_d, const double* vec_d, double* out_d, int32_t nb, void* stream);

/* K0: GreensFunction._generate_kmesh (greens.py:100-111): flat index i*n1*n2 + j*n2 + l ->
 * (i/max(n0,1), j/max(n1,1), l/max(n2,1)), IEEE division (bit-exact). k_out_d [count][3].                    */
int sktb_kmesh(const int32_t nk[3], int64_t first, int64_t count, double* k_out_d, void* stream);

/* K1 (materialising): Hamiltonian.get_ham (hamiltonian.py:98-106, calc_g :275-317) for nk k-points.
 * k_d [nk][3] fractional; H_d [nk][N][N] row-major, N = soc ? 2*n_orb : n_orb.
 * nonherm_d (may be NULL) [nk] int32: 1 where max Re(H - H^dagger) > 1e-9 (the gate of hamiltonian.py:112).  */
int sktb_hk(sktb_plan* plan, const double* k_d, int64_t nk, int32_t soc, double* H_d, int32_t* nonherm_d,
            void* stream);

/* K1+K2: solve_kpath's per-k body (hamiltonian.py:47-56, 98-126, 223-231): H(k), Hermitian eigensolve on the
 * lower triangle (LAPACK zheevd UPLO='L' semantics), ascending order.
 * evals_d [N][ld_k], k fastest (the reference's (N, nk) layout); this call writes columns k_off .. k_off+nk-1.
 * evecs_d NULL or [N][ld_k][N]: row `band` is the eigenvector (the reference's eig.T, hamiltonian.py:124).
 * nonherm_d NULL or [ld_k] int32 flags as in sktb_hk.                                                        */
int sktb_solve(sktb_plan* plan, const double* k_d, int64_t nk, int32_t soc, double* evals_d, double* evecs_d,
               int64_t ld_k, int64_t k_off, int32_t* nonherm_d, void* stream);

/* Same with k generated on the device from a mesh (K0 fused; never materialises the k list):
 * mesh points [first, first+count) of nk_mesh -> columns k_off.. of evals_d.                                 */
int sktb_solve_mesh(sktb_plan* plan, const int32_t nk_mesh[3], int64_t first, int64_t count, int32_t soc,
                    double* evals_d, int64_t ld_k, int64_t k_off, int32_t* nonherm_d, void* stream);

/* Batched Hermitian eigensolve of caller-supplied matrices: Hamiltonian._sol_ham (hamiltonian.py:108-126).
 * H_d [nm][N][N] row-major complex (only the lower triangle and the real diagonal enter the eigensolve, as with
 * numpy's default UPLO='L'); evals_d [N][nm]; evecs_d NULL or [N][nm][N]; nonherm_d NULL or [nm] gate flags.
 * Entries are expected at energy scale: no LAPACK-style rescaling is done, 1e-140 < |H_ij| < 1e140 (or exactly 0). */
int sktb_heev(sktb_plan* plan, const double* H_d, int32_t N, int64_t nm, double* evals_d, double* evecs_d,
              int32_t* nonherm_d, void* stream);

/* Diagnostic entry of the two-stage tridiagonalisation that sktb_heev / sktb_solve* use for eigenvalues of large matrices
 * (N >= 208 by default; csrc/heev_twostage.cuh): stage 1 reduces H to a Hermitian band of width 16 (blocked Householder,
 * GEMMs on the FP64 tensor instruction), stage 2 chases the band to tridiagonal form.  Same input convention as sktb_heev.
 * band_d NULL or [nm][N][32] complex: band after stage 1, band[j][d] = A[j+d][j] (d <= 16, rest zero);
 * de_d NULL or [nm][2][N]: diagonal and sub-diagonal of the tridiagonal matrix.  Used by the parity tests to check each
 * stage against LAPACK (the eigenvalues of the band and of the tridiagonal equal those of H).                          */
int sktb_heev_twostage(sktb_plan* plan, const double* H_d, int32_t N, int64_t nm, double* band_d, double* de_d, void* stream);

/* Host-buffer front end of sktb_solve: pinned staging, H2D of k / compute / D2H of results overlapped in
 * chunks on internal streams.  k_h [nk][3], evals_h [N][nk], evecs_h NULL or [N][nk][N] are HOST pointers.
 * *nonherm_any = 1 if any k tripped the Hermiticity gate.  Synchronous.                                      */
int sktb_solve_host(sktb_plan* plan, const double* k_h, int64_t nk, int32_t soc, double* evals_h, double* evecs_h,
                    int32_t* nonherm_any);

/* K1+K2+K3: GreensFunction.dos / ldos / ldos_by_orbital / SurfaceGreensFunction.edge_dos
 * (greens.py:113-336, 431-523) through the closed form  -Im G_ii / pi = sum_n |U_in|^2 (eta/pi)/((E-e_n)^2+eta^2).
 * k-points: k_d [k_count][3] if non-NULL, else mesh points [k_first, k_first+k_count) of nk_mesh.
 * Groups: n_groups == 0 -> total DOS, out_d [nE]. Otherwise group g sums |U_in|^2 over the Hamiltonian row
 * indices grp_idx[grp_ptr[g] .. grp_ptr[g+1]) (HOST arrays; the caller encodes the reference's idx /
 * idx+n_orbitals convention), out_d [n_groups][nE].  Sums are UN-normalised (caller divides by n_kpts and
 * all-reduces across ranks); out_d is overwritten.  energies_d [nE] device.  nonherm_any_h: host int32.      */
int sktb_dos(sktb_plan* plan, const double* k_d, const int32_t nk_mesh[3], int64_t k_first, int64_t k_count,
             int32_t soc, const double* energies_d, int32_t nE, double eta, int32_t n_groups,
             const int32_t* grp_ptr, const int32_t* grp_idx, double* out_d, int32_t* nonherm_any_h, void* stream);

/* Spectral maps A(k,E) = sum_{n} w_n(k) L(E - e_n(k)) per k (greens.py:338-376 spectral_kpath, :525-557
 * edge_spectral_kpath): k_d [nk][3]; one group (n_idx == 0 -> trace). out_d [nk][nE].                        */
int sktb_spectral(sktb_plan* plan, const double* k_d, int64_t nk, int32_t soc, const double* energies_d, int32_t nE,
                  double eta, int32_t n_idx, const int32_t* idx, double* out_d, int32_t* nonherm_any_h,
                  void* stream);

/* max over bonds of |T_b[mu,nu] - T_rev(b)[nu,mu]| and (soc != 0) over soc_mat of |S - S^dagger|: 0 (to rounding) <=> the H(k)
 * that get_ham(k, l_soc=soc) returns is Hermitian for every k.  Unlike sktb_plan_asym_bound (the real-part test of hamiltonian.py:112) this sees the purely imaginary
 * asymmetry of the f-orbital blocks; the Python layer uses it to choose between sktb_dos / sktb_spectral
 * (eigenpairs, exact for Hermitian H) and sktb_gf_dos (dense inverse, what the reference computes in every case).      */
int sktb_plan_herm_defect(const sktb_plan* plan, int32_t soc, double* defect);

/* K1+K5: GreensFunction.dos / ldos / ldos_by_orbital / spectral_kpath, SurfaceGreensFunction.* (greens.py:113-376, 431-557)
 * exactly as the reference computes them: G = inv((E + i eta) I - H(k)) of the FULL matrix get_ham returns (greens.py:70,
 * :455), one complex inverse per (k, E) (csrc/gf_inverse.cuh).  8 N^3 flops per (k, E) instead of one eigensolve per k: meant
 * for models whose H(k) is not Hermitian, where the eigenpair route is not what the reference returns.
 * k-points, groups and energies as in sktb_dos (n_groups == 0 -> trace; indices count with multiplicity).
 * per_k == 0: out_d [max(n_groups,1)][nE] = sum over the k-points of -(1/pi) sum_{i in group} Im G_ii, UN-normalised;
 * per_k == 1: out_d [k_count][max(n_groups,1)][nE], one map per k-point (spectral_kpath / edge_spectral_kpath).        */
int sktb_gf_dos(sktb_plan* plan, const double* k_d, const int32_t nk_mesh[3], int64_t k_first, int64_t k_count, int32_t soc,
                const double* energies_d, int32_t nE, double eta, int32_t n_groups, const int32_t* grp_ptr,
                const int32_t* grp_idx, int32_t per_k, double* out_d, void* stream);

/* Batched dense inverse of caller-supplied matrices (the counterpart of sktb_heev for the Green's function, greens.py:70):
 * H_d [nm][N][N] row-major complex (every entry is read - no triangle reading), energies_d [nE] device;
 * G_d NULL or [nm][nE][N][N] = inv((E_e + i eta) I - H_m), Gauss-Jordan with partial (row) pivoting;
 * trace_d NULL or [nm][nE] = -(1/pi) Im Tr G.  With G_d == NULL the register-resident kernels run (one matrix per warp for
 * N <= 32, one per CTA with the matrix spread over its registers for N <= 96), else the shared-memory kernel.              */
int sktb_gf_inverse(sktb_plan* plan, const double* H_d, int32_t N, int64_t nm, const double* energies_d, int32_t nE, double eta,
                    double* G_d, double* trace_d, void* stream);

/* GreensFunction.retarded (greens.py:46-71): G_d [N][N] row-major complex = inv((E + i eta) I - H(k)), k_h HOST [3].   */
int sktb_gf_retarded(sktb_plan* plan, const double k_h[3], double E, double eta, int32_t soc, double* G_d, void* stream);

/* Hamiltonian.get_dos (hamiltonian.py:233-253): Gaussian-broadened DOS of a set of eigenvalues,
 *   out[e] = sum_i exp(-(E_e - eps_i)^2 / (2 w^2)) / (pi w sqrt(2))      (the reference's own prefactor),
 * evals_d [n] device (any layout, flat), energies_d [nE], out_d [nE] overwritten.  Deterministic (fixed-order sums). */
int sktb_gauss_dos(sktb_plan* plan, const double* evals_d, int64_t n, const double* energies_d, int32_t nE, double w,
                   double* out_d, void* stream);

/* Band energy of Forces.get_total_energy (forces.py:381-395): sum over the k-points of the n_occ lowest
 * eigenvalues, evals_d [N][ld] ascending per k (the layout sktb_solve / sktb_solve_mesh write), columns
 * [0, nk).  out_d: one double, UN-normalised (caller multiplies by weight = 1/n_kpts and the spin factor).   */
int sktb_band_energy(sktb_plan* plan, const double* evals_d, int64_t ld, int64_t nk, int32_t n_occ, double* out_d,
                     void* stream);

/* Measurement helpers (bench.py): sustained FP64 FMA throughput of the device in TFLOP/s (DFMA microbenchmark,
 * `ms` = duration of the timed launch), and the last solve's kernel configuration as text.                   */
int sktb_fp64_peak(int device, double* tflops, double* ms);
const char* sktb_last_kernel_info(void);
/* CUDA-event time in ms of the most recent K1+K2 launch sequence issued by sktb_solve / sktb_solve_mesh on the
 * plan (blocks until that work has finished), and how many kernels it launched.                             */
int sktb_last_solve_time(sktb_plan* plan, double* ms, int32_t* n_launches);

/* Number of kernels this library has launched in the process so far (bench.py's gpu_launches).             */
long long sktb_launch_count(void);
/* Number of matrices whose tridiagonal QL hit its iteration limit since plan creation (should stay 0).      */
int sktb_ql_failures(sktb_plan* plan, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* SKTB_H */
