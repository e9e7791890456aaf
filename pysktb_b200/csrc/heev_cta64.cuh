// K2a for 32 < N <= 64 (config C4: two spdf atoms with SOC): CTA-resident Householder tridiagonalisation, eigenvalues
// only.  The pair layout of heev_pair.cuh generalised to four column classes: one matrix per CTA of 128 threads,
//
//   thread t = r + 32 h,  r in [0,32): slot A = row r, slot B = row 63 - r;  h = warp = column class
//   register window  B[slot][m]  holds column  base + 4 m + h  (base = 4 floor(j/4)) of the mirrored-lower matrix
//
// so the 64 x 64 complex matrix (64 KB) lives in 128 x 128 registers, every warp owns ALL row pairs (norm and
// dot-product reductions stay inside a warp) and a quarter of the columns (the mat-vec partial sums of the four
// warps meet in shared memory).  One code body per phase serves the four sub-steps of a register cycle: the
// register that holds the next pivot column / the unit entry of the next reflector is chosen by a warp-uniform
// select between registers 0 and 1, and the window is rotated by one register every fourth step.  Three CTA
// barriers per step (pivot column published | mat-vec partials | reflector published).  H(k) comes from K1
// (hk_kernel, row-major c128 in the workspace; LAPACK UPLO='L' semantics are applied while loading), (d, e) go to
// the scratch consumed by tql_small_kernel<64>.  FP64 CUDA cores; see heev_pair.cuh for the update formulas.
// (included by heev_small.cuh inside namespace sktb::small)
#pragma once

constexpr int C64_THREADS = 128;
constexpr int C64_SMEM_CPLX = 4 * 128 + 2 + 4 * 64 + 32 + 32;   // V | P | X | spare | alpha,pad | ypart[4][64] | D[64] | E[64]

struct C64Smem {
    cplx *V, *P, *X, *alpha, *yp;
    double *D, *E;
    cplx* vst = nullptr;     // global [64][64] reflector store of this matrix (eigenvector path) or NULL
    cplx* taust = nullptr;   // global [64]
};
constexpr int C64_VSTORE_CPLX = 64 * 64 + 64;
SKTB_D C64Smem c64_smem(cplx* base) {
    C64Smem s;
    s.V = base; s.P = base + 128; s.X = base + 256; s.alpha = base + 512; s.yp = base + 514;
    s.D = reinterpret_cast<double*>(base + 514 + 256); s.E = s.D + 64;
    return s;
}

struct C64Lane {
    int r, h, iA, iB;
};

SKTB_D double c64_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// publish pivot column jn (owner class only): X masked to rows > jn+1, alpha = A[jn+1][jn], d[jn]
template <int NS>
SKTB_D void c64_publish(const cplx (&x)[2], const int jn, const bool owner, const C64Lane& L, const C64Smem& S) {
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        const int i = s ? L.iB : L.iA;
        const cplx xv = (i > jn + 1) ? x[s] : cmake(0.0, 0.0);
        if (owner) S.X[i] = xv;
        if (owner && i == jn + 1) *S.alpha = x[s];
        if (owner && i == jn) S.D[jn] = x[s].x;
    }
}

// after barrier 1: own-row entries of the pivot column, its norm, the reflector scalars
template <int NS>
SKTB_D void c64_reflector(const C64Lane& L, const C64Smem& S, cplx (&xo)[2], double& beta, cplx& tau, cplx& scale) {
    double xnp = 0.0;
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        xo[s] = S.X[s ? L.iB : L.iA];
        xnp = fma(xo[s].x, xo[s].x, xnp); xnp = fma(xo[s].y, xo[s].y, xnp);
    }
    const cplx alpha = *S.alpha;
    const double xn = c64_wsum(xnp);
    pair_larfg(alpha, xn, beta, tau, scale);
}

// mat-vec partials -> (barrier 2) -> y, p, v, g -> publish v (warp 0), p (warp 1), e (warp 2) -> (barrier 3)
template <int NS>
SKTB_D void c64_finish(PairCarry& c, const cplx (&ypart)[2], const cplx (&xo)[2], const double beta, const cplx tau, const cplx scale,
                       const int jn, const C64Lane& L, const C64Smem& S) {
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) S.yp[L.h * 64 + (s ? L.iB : L.iA)] = ypart[s];
    __syncthreads();
    c.tau2 = fma(tau.x, tau.x, tau.y * tau.y);
    double g = 0.0;
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        const int i = s ? L.iB : L.iA;
        const cplx y0 = S.yp[i], y1 = S.yp[64 + i], y2 = S.yp[128 + i], y3 = S.yp[192 + i];
        cplx y = cmake((y0.x + y1.x) + (y2.x + y3.x), (y0.y + y1.y) + (y2.y + y3.y));
        if (i <= jn) y = cmake(0.0, 0.0);
        c.p[s] = cmul(tau, y);
        c.v[s] = (i == jn + 1) ? cmake(1.0, 0.0) : cmul(xo[s], scale);
        g = fma(y.x, c.v[s].x, g); g = fma(y.y, c.v[s].y, g);
    }
    c.g = g;
    if (L.h < 2) {
        cplx* dst = L.h ? S.P : S.V;
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) dst[s ? L.iB : L.iA] = L.h ? c.p[s] : c.v[s];
    }
    if (L.h == 2 && L.r == 0) S.E[jn] = beta;
    if (S.vst && jn < 63) {          // eigenvector path: keep reflector jn and tau (warp 3 is otherwise idle here)
        if (L.h == 3) {
#pragma unroll
            for (int s = 2 - NS; s < 2; ++s) S.vst[(size_t)jn * 64 + (s ? L.iB : L.iA)] = c.v[s];
            if (NS == 1) S.vst[(size_t)jn * 64 + L.iA] = cmake(0.0, 0.0);      // retired rows: every entry of the store is defined
            if (L.r == 0) S.taust[jn] = tau;
        }
    }
    __syncthreads();
}

// one Householder step j (runtime), HALF live registers per slot, NS live slots
template <int HALF, int NS>
SKTB_D void c64_step(cplx (&B)[2][16], PairCarry& c, const int j, const C64Lane& L, const C64Smem& S) {
    const int base = j & ~3, jn = j + 1;
    const cplx* Pl = S.P + base + L.h;
    const cplx* Vl = S.V + base + L.h;
    const cplx* Xl = S.X + base + L.h;
    // ---- p-half of the update on the live window, one butterfly stage of the w' reduction per slice of it (source
    // order matters to ptxas: written first, the reduction's DADDs stall the in-order warp ahead of the DFMA stream) ----
    constexpr int CH = (HALF + 5) / 6;
    double G = c.g;
#pragma unroll
    for (int m = 0; m < HALF; ++m) {
        const cplx pq = Pl[4 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) pair_upd1(B[s][m], c.v[s], pq);
        if ((m + 1) % CH == 0 && (m + 1) / CH <= 5) G += __shfl_xor_sync(0xffffffffu, G, 32 >> ((m + 1) / CH));
    }
#pragma unroll
    for (int st = HALF / CH + 1; st <= 5; ++st) G += __shfl_xor_sync(0xffffffffu, G, 32 >> st);
    const double coef = -c.tau2 * G;
    cplx wp[2], x[2];
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) wp[s] = cmake(fma(coef, c.v[s].x, c.p[s].x), fma(coef, c.v[s].y, c.p[s].y));
    // ---- v-half on registers 0 and 1 (one of them holds the next pivot column), publish it ----
    {
        const cplx v0 = Vl[0], v1 = Vl[4];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) { pair_upd1(B[s][0], wp[s], v0); pair_upd1(B[s][1], wp[s], v1); }
    }
    const bool me1 = (jn - base) >= 4;            // next pivot column sits in register 1 (class 0) after a full cycle
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) x[s] = me1 ? B[s][1] : B[s][0];
    c64_publish<NS>(x, jn, L.h == (jn & 3), L, S);
    __syncthreads();
    cplx xo[2];
    double beta; cplx tau, scale;
    c64_reflector<NS>(L, S, xo, beta, tau, scale);
    // ---- v-half on the remaining registers fused with the look-ahead mat-vec ----
    cplx z[2];
    {
        const cplx x0 = Xl[0], x1 = Xl[4];        // masked: zero for columns <= jn + 1
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) {
            z[s] = cmake(0.0, 0.0);
            cfma(z[s], B[s][0], x0); cfma(z[s], B[s][1], x1);
        }
    }
#pragma unroll
    for (int m = 2; m < HALF; ++m) {
        const cplx vq = Vl[4 * m];
        const cplx xq = Xl[4 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) {
            pair_upd1(B[s][m], wp[s], vq);
            cfma(z[s], B[s][m], xq);
        }
    }
    // column jn+1 (where v' = 1): class (jn+1)&3, register 0 or 1
    const bool ma1 = (jn + 1 - base) >= 4;
    const bool a_here = L.h == ((jn + 1) & 3);
    cplx ypart[2];
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        const cplx acol = ma1 ? B[s][1] : B[s][0];
        cplx yp = cmake(a_here ? acol.x : 0.0, a_here ? acol.y : 0.0);      // the column where v' = 1 enters as the addend of the product
        cfma(yp, scale, z[s]);
        ypart[s] = yp;
    }
    c64_finish<NS>(c, ypart, xo, beta, tau, scale, jn, L, S);
}

// SPLIT: stop after the two-slot phases (steps 0..31); the trailing 32 x 32 block goes to the N <= 32 kernels.
// J0 (a multiple of 8, CTA-uniform): phases whose steps lie before J0 are skipped - the caller loaded the register
// window as it would stand at step J0 (front-padded matrices: those steps would only walk over zero padding).
template <int HALF, bool SPLIT>
SKTB_D void c64_phases(cplx (&B)[2][16], PairCarry& c, int& j, const int J0, const C64Lane& L, const C64Smem& S) {
    constexpr int NS = HALF > 8 ? 2 : 1;
    constexpr int JSTART = 4 * (16 - HALF);
    if (J0 <= JSTART) {
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
#pragma unroll 1
            for (int q = 0; q < 4; ++q, ++j)
                if (j <= 62) c64_step<HALF, NS>(B, c, j, L, S);       // step 63 does not exist (warp-uniform guard)
#pragma unroll
            for (int s = 2 - NS; s < 2; ++s) {
#pragma unroll
                for (int m = 0; m + 1 < HALF; ++m) B[s][m] = B[s][m + 1];
                B[s][HALF - 1] = cmake(0.0, 0.0);
            }
        }
    }
    if constexpr (HALF > 2 && !(SPLIT && HALF - 2 <= 8)) c64_phases<HALF - 2, SPLIT>(B, c, j, J0, L, S);
}

struct C64Args {
    const cplx* H;          // [nm][N][N] row-major (K1 output)
    int N; long long nm;
    double* scratch;        // [nm][2][64]  (d | e)
    cplx* vstore = nullptr; // [nm][C64_VSTORE_CPLX] reflectors + tau (eigenvector path) or NULL
    cplx* blocks = nullptr; // SPLIT: [nm][32][32] trailing block after step 31 (row-major, full rows)
    int ld = 0; long long mstride = 0, hoff = 0;   // input layout: element (r, c) of matrix km at H[km * mstride + hoff + r * ld + c]
                            // (0 -> dense [N][N]); a peeled run passes the trailing block of a larger matrix in place
    int front = 0;          // 1: N < 64 is padded at the FRONT of the 64-wide frame (rows / columns < 64 - N are zero), so
                            // that the steps over the padding can be skipped: the run starts at step 8 * ((64 - N) / 8)
    // BUILD mode (K0 + K1 fused: H(k) never exists in memory): H == NULL, the CTA assembles the lower triangle of the
    // orbital-space h(k) in shared memory from the plan's bond tables and every thread picks its window entries from it
    PlanView pv{};
    const double* k = nullptr;      // [nm][3] fractional k-points, or NULL -> mesh points first.. of (mesh0, mesh1, mesh2)
    int mesh0 = 1, mesh1 = 1, mesh2 = 1;
    long long first = 0;
    int soc = 0;
    const cplx* Sdense = nullptr;   // soc: dense soc_mat [N][N] row-major (only the lower triangle is read)
};

// host-side description of a fused build (NULL -> the caller materialised H with K1)
struct C64Build {
    PlanView pv;
    const double* k; const int32_t* mesh; long long first;
    int soc;
    const cplx* Sdense;
    size_t smem_bytes() const { const size_t n = (size_t)pv.n_orb; return (n * (n + 1) / 2 + (size_t)pv.n_bonds) * sizeof(cplx); }
};

// K0 + K1 in the CTA: bond phases (calc_g arithmetic), then the packed lower triangle hs[mu (mu+1)/2 + nu], nu <= mu, of
//   h_mu,nu(k) = sum_{bonds b of pair (atom(mu), atom(nu))} T_b[mu,nu] phase_b + delta onsite        (hamiltonian.py:98-103)
// with the accumulation order of hk_kernel (pair-sorted bonds, image ascending) - the two paths agree bit for bit.
SKTB_D void c64_build_h(const C64Args& a, const long long km, cplx* hs, cplx* phs) {
    const PlanView& P = a.pv;
    const int n = P.n_orb, t = threadIdx.x;
    double k0, k1, k2;
    if (a.k) { k0 = a.k[3 * km]; k1 = a.k[3 * km + 1]; k2 = a.k[3 * km + 2]; }
    else kmesh_point(a.first + km, a.mesh0, a.mesh1, a.mesh2, k0, k1, k2);
    for (int b = t; b < P.n_bonds; b += C64_THREADS) phs[b] = bond_phase(P.rec, k0, k1, k2, P.bond_vec + 3 * b);
    __syncthreads();
    for (int e = t; e < n * n; e += C64_THREADS) {
        const int mu = e / n, nu = e - mu * n;
        if (nu > mu) continue;
        cplx v = cmake(0.0, 0.0);
        const int ai = P.orb_atom[mu], aj = P.orb_atom[nu];
        const int pr = P.pair_of[ai * P.n_atoms + aj];
        if (pr >= 0) {
            const int ni_off = mu - P.atom_start[ai], nj_off = nu - P.atom_start[aj];
            const int nj = P.atom_start[aj + 1] - P.atom_start[aj];
            for (int b = P.pair_begin[pr]; b < P.pair_begin[pr + 1]; ++b) {
                const double tv = P.T[P.bond_off[b] + (long long)ni_off * nj + nj_off];
                const cplx ph = phs[b];
                v.x = fma(tv, ph.x, v.x); v.y = fma(tv, ph.y, v.y);
            }
        }
        if (mu == nu) v.x += P.onsite[mu];
        hs[mu * (mu + 1) / 2 + nu] = v;
    }
    __syncthreads();
}

// prologue (pivot column J0: class 0, register 0 of a window whose base is J0; registers beyond the live window and
// X beyond row 63 are zero, so the loop runs over all 16 registers) + phases
template <bool SPLIT>
SKTB_D void c64_run(cplx (&B)[2][16], PairCarry& c, const int J0, const C64Lane& L, const C64Smem& S) {
    {
        cplx x[2] = {B[0][0], B[1][0]};
        c64_publish<2>(x, J0, L.h == 0, L, S);
        __syncthreads();
        cplx xo[2];
        double beta; cplx tau, scale;
        c64_reflector<2>(L, S, xo, beta, tau, scale);
        const cplx* Xl = S.X + J0 + L.h;
        cplx ypart[2];
        ypart[0] = ypart[1] = cmake(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const cplx xq = Xl[4 * m];
            cfma(ypart[0], B[0][m], xq); cfma(ypart[1], B[1][m], xq);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            cplx yp = cmake((L.h == 1) ? B[s][0].x : 0.0, (L.h == 1) ? B[s][0].y : 0.0);     // column J0+1 = class 1, register 0
            cfma(yp, scale, ypart[s]);
            ypart[s] = yp;
        }
        c64_finish<2>(c, ypart, xo, beta, tau, scale, J0, L, S);
    }
    int j = J0;
    c64_phases<16, SPLIT>(B, c, j, J0, L, S);
}

template <bool SPLIT, bool BUILD = false>
__global__ void __launch_bounds__(C64_THREADS, 2) tridiag_cta64_kernel(C64Args a) {
    __shared__ __align__(16) cplx smem[C64_SMEM_CPLX];
    extern __shared__ __align__(16) unsigned char c64_dyn[];       // BUILD: hs (packed lower triangle of h(k)) | bond phases
    const int t = threadIdx.x;
    C64Lane L;
    L.r = t & 31; L.h = t >> 5; L.iA = L.r; L.iB = 63 - L.r;
    C64Smem S = c64_smem(smem);
    for (int q = t; q < C64_SMEM_CPLX; q += C64_THREADS) smem[q] = cmake(0.0, 0.0);
    __syncthreads();
    const int N = a.N;
    const int off = a.front ? 64 - N : 0;                   // first real row / column of the frame
    const int J0 = off & ~7;                                // first step that is run (a multiple of 8: a phase boundary)
#pragma unroll 1
    for (long long km = blockIdx.x; km < a.nm; km += gridDim.x) {
        // ---- load rows iA, iB x columns J0 + 4m + h with zheevd(UPLO='L') semantics: upper triangle = conj(lower) ----
        const int ldh = a.ld ? a.ld : N;
        const cplx* H = a.H + (size_t)km * (a.mstride ? a.mstride : (long long)N * N) + a.hoff;
        if (a.vstore) { S.vst = a.vstore + (size_t)km * C64_VSTORE_CPLX; S.taust = S.vst + 4096; }
        cplx B[2][16];
        if constexpr (BUILD) {
            // H(k) = kron(h, I2) + soc_mat (soc) | h, read with zheevd(UPLO='L') semantics from the lower triangle in
            // shared memory; the soc entries come from the dense table (L1 / L2 resident, shared by every k-point)
            cplx* hs = reinterpret_cast<cplx*>(c64_dyn);
            const int n = a.pv.n_orb;
            c64_build_h(a, km, hs, hs + (size_t)n * (n + 1) / 2);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = (s ? L.iB : L.iA) - off;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int col = J0 + 4 * m + L.h - off;
                    const bool ok = i >= 0 && col >= 0 && i < N && col < N;
                    const int hi = ok ? max(i, col) : 0, lo = ok ? min(i, col) : 0;
                    cplx e;
                    if (a.soc) {
                        const int mu = hi >> 1, nu = lo >> 1;
                        e = hs[mu * (mu + 1) / 2 + nu];
                        if ((hi ^ lo) & 1) e = cmake(0.0, 0.0);            // kron(h, I2): opposite spins do not couple
                        const cplx sv = a.Sdense[(size_t)hi * N + lo];
                        e = cmake(e.x + sv.x, e.y + sv.y);
                    } else {
                        e = hs[hi * (hi + 1) / 2 + lo];
                    }
                    if (col > i) e.y = -e.y;
                    if (col == i) e.y = 0.0;
                    if (!ok) e = cmake(0.0, 0.0);
                    B[s][m] = e;
                }
            }
        } else {
        // branch-free: the address of the lower-triangle element is clamped, all 32 loads are independent and in flight
        // together; conjugation / real diagonal / zero padding are applied to the loaded value
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = (s ? L.iB : L.iA) - off;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int col = J0 + 4 * m + L.h - off;
                const bool ok = i >= 0 && col >= 0 && i < N && col < N;
                const int hi = ok ? max(i, col) : 0, lo = ok ? min(i, col) : 0;
                cplx e = H[(size_t)hi * ldh + lo];
                if (col > i) e.y = -e.y;
                if (col == i) e.y = 0.0;
                if (!ok) e = cmake(0.0, 0.0);
                B[s][m] = e;
            }
        }
        }
        PairCarry c;
        c.eval = 0.0;
        c64_run<SPLIT>(B, c, J0, L, S);
        if (t >= off && t < 64) {
            double* out = a.scratch + (size_t)km * 128;
            if (!SPLIT || t < 32) { out[t - off] = S.D[t]; out[64 + t - off] = S.E[t]; }   // SPLIT: d, e of steps < 32; the rest comes from the 32-wide kernels
        }
        if (SPLIT) {    // slot B = rows 32..63; registers 0..7 hold columns 32 + 4m + h (base = 32 after 32 steps)
            cplx* blk = a.blocks + (size_t)km * 1024 + (size_t)(31 - L.r) * 32 + L.h;
#pragma unroll
            for (int m = 0; m < 8; ++m) blk[4 * m] = B[1][m];
        }
        __syncthreads();
    }
}
