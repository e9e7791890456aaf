// K0+K1+K2a for 24 < N <= 32 (the north-star 32-orbital model): "pair layout" Householder tridiagonalisation,
// split into a WIDE kernel (steps 0..15 on the 32x32 matrix, 2 warps per scheduler, register-heavy, FP64-pipe
// bound) and a TAIL kernel (steps 16..30 on the trailing 16x16 block, two matrices per warp, latency bound, few
// registers).  The launch code (heev_small.cuh) pipelines chunks of k-points over two streams so that the tail and
// QL kernels of chunk c co-reside on the SMs with the wide kernel of chunk c+1 and fill its issue bubbles.
//
// Why a second layout.  The row-per-lane kernels (heev_small.cuh) read every reflector entry v_q, w_q as a
// warp-uniform LDS.128 (2 shared-memory wavefronts) that feeds only 4 DFMA per lane: per Householder step the
// shared-memory data pipe needs as many SM cycles (6 W) as the FP64 pipe, and ncu shows both ~55 % busy with the
// warps waiting on each other's LDS.  Here every lane owns TWO rows and HALF of the columns, so each broadcast
// load feeds 8 DFMA (LN = lanes per matrix = matrix size, 32 or 16):
//
//   lane-in-matrix l = 2 r + h,  r in [0, LN/2): slot A = row r, slot B = row LN-1-r,  h = column parity
//   register window  B[slot][m]  holds column  base + 2 m + h  of the mirrored-lower Hermitian matrix
//
//   * pairing row r with row LN-1-r retires slot A in ALL lanes after LN/2 steps: the second half of a
//     reduction runs single-slot code (half the instructions) without re-packing;
//   * the window is rotated by one register every second step (the odd step writes its results one register
//     down, so the rotation costs no moves) and shrinks by S/2 registers per phase of S steps;
//   * look-ahead: while step j's rank-2 update streams over the window, the new pivot column (updated first and
//     published through shared memory) is multiplied in, so y' = A_new v' of step j+1 is complete when the
//     update ends;  the norm reduction and the branch-free scalar zlarfg chain hide under the same loop;
//   * the rank-2 update is applied as  A -= v p^H + w' v^H  with  w' = p - |tau|^2 Re(y^H v) v  (identical to
//     LAPACK's  v w^H + w v^H, w = p - tau/2 (p^H v) v):  the p-half needs no reduction, so the one remaining
//     warp reduction (a single real number) hides under it.
//
// Per step and window width W the warp issues 12 W DFMA (two slots) against ~3 W + 36 shared-memory wavefronts
// (was 6 W + 34).  Semantics are unchanged: LAPACK zhetd2(UPLO='L') on the matrix numpy's eigvalsh reads
// (hamiltonian.py:119), output (d, e) into the scratch consumed by tql_small_kernel.  The Hermiticity-gate
// variant stays on the row-per-lane kernels.  tools/proto_pair_layout.py is the numpy model of this file.
// (included by heev_small.cuh inside namespace sktb::small, after the shared build/phase helpers)
#pragma once

constexpr int PAIR_WARPS = 4;
constexpr int PAIR_VSTORE_CPLX = 32 * 32 + 32;      // reflectors [j][i] + tau[j] per k-point (eigenvector path)
#ifndef SKTB_PAIR_S
#define SKTB_PAIR_S 4
#endif
#ifndef SKTB_PAIR_EARLY
#define SKTB_PAIR_EARLY 0      // > 0: columns of the p-half that run before the new pivot column is published (see pair_step)
#endif

// One buffer set per matrix: reflector broadcast buffers of 2 LN entries whose upper half stays zero (window
// columns beyond the matrix read v = p = x = 0).
struct PairBuf {
    cplx *V0, *V1, *P, *X, *alpha;   // V double-buffered by step parity; X, alpha, D: one copy per column parity (branch-free publish)
    double* D;                       // [2][LN] diagonal: entry j lives in copy j & 1
    cplx* vst = nullptr;             // global [32][32] reflector store of this matrix (eigenvector path) or NULL
    cplx* taust = nullptr;           // global [32] tau
    int voff = 0;                    // row / reflector offset of this (sub)matrix inside the store
    int vstride = 32;                // row stride of the store (32, or 64 when the block is the tail of a 64-wide matrix)
};
template <int LN> __host__ __device__ constexpr int pair_buf_cplx() { return 11 * LN + 2; }   // V0 | V1 | P | X[2] | alpha[2] | D[2]

template <int LN>
SKTB_D PairBuf pair_buf(cplx* base) {
    PairBuf b;
    b.V0 = base; b.V1 = base + 2 * LN; b.P = base + 4 * LN; b.X = base + 6 * LN; b.alpha = base + 10 * LN;
    b.D = reinterpret_cast<double*>(base + 10 * LN + 2);
    return b;
}

struct PairLane {
    int l, r, h, iA, iB;             // lane in matrix, row pair, column parity, rows of slot A / B
};
template <int LN>
SKTB_D PairLane pair_lane(const int lane) {
    PairLane L;
    L.l = lane & (LN - 1); L.r = L.l >> 1; L.h = L.l & 1; L.iA = L.r; L.iB = LN - 1 - L.r;
    return L;
}

SKTB_D void pair_upd1(cplx& t, const cplx v, const cplx pq) {
    t.x = fma(-v.x, pq.x, t.x); t.y = fma(-v.y, pq.x, t.y);
    t.x = fma(-v.y, pq.y, t.x); t.y = fma(v.x, pq.y, t.y);
}

// 1/sqrt(a), 1/a for normal positive a (eV-scale squared norms): hardware seed + one cubic / two quadratic
// Newton steps, no slow-path branch (the CUDA library versions branch on denormals / infinities).
SKTB_D double pair_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a * y, y, 1.0);                 // 1 - a y^2
    const double c = fma(e, 0.375, 0.5);
    return fma(c, y * e, y);                              // y (1 + e/2 + 3 e^2 / 8)
}
SKTB_D double pair_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);                                     // cubic step: seed error e <= ~2^-19 -> e^3 below the last rounding
#if !SKTB_RCP3
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
#endif
    return r;
}

// sum over the LN/2 row pairs (lanes of equal column parity within the matrix)
template <int LN>
SKTB_D double pair_rsum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (LN == 32) v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// x with its sign bit flipped when `flip`, and x or +0.0: integer operations on the bit pattern (the FP64 pipe is the bound of these
// kernels; a DMUL by +-1 / by a 0-1 mask or a DADD from zero would take an FP64 issue slot each)
SKTB_D double pair_flip(double x, bool flip) {
    return __hiloint2double(__double2hiint(x) ^ (flip ? (int)0x80000000 : 0), __double2loint(x));
}
SKTB_D double pair_keep(double x, bool keep) {
    const int m = keep ? -1 : 0;
    return __hiloint2double(__double2hiint(x) & m, __double2loint(x) & m);
}

// zlarfg, branch-free (straight-line code, so that ptxas can schedule this scalar chain under the update loop)
SKTB_D void pair_larfg(const cplx alpha, const double xn, double& beta, cplx& tau, cplx& scale) {
    // The reciprocal square root / reciprocal run on every path (inputs replaced by 1, results multiplied by a 0/1
    // mask when the reflector is trivial): a result that is only *selected* lets ptxas branch around the chain, and
    // that branch would end the basic block in the middle of the update loop this chain is meant to hide under.
    const bool trivial = (xn == 0.0) && (alpha.y == 0.0);
    const double live = trivial ? 0.0 : 1.0;
    const double n2 = fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xn));
#if SKTB_RCP3
    const double inrm = pair_keep(pair_rsqrt(trivial ? 1.0 : n2), !trivial);
#else
    const double inrm = pair_rsqrt(trivial ? 1.0 : n2) * live;
#endif
#if SKTB_RCP3
    // FP64-issue diet (the scalar chain is 20-40 % of the FP64 instructions of the tail steps): the sign of beta is applied by
    // selecting between x and -x (a sign-bit operation, not a DMUL by +-1), and dx = alpha.x - beta reuses beta - alpha.x
    const bool pos = alpha.x >= 0.0;
    const double nrm = n2 * inrm;
    beta = trivial ? alpha.x : pair_flip(nrm, pos);
    const double ib = pair_flip(inrm, pos);
    const double t = beta - alpha.x, dy = alpha.y;
    tau = cmake(t * ib, -alpha.y * ib);
    const double den = fma(t, t, dy * dy);
    const double id = pair_keep(pair_rcp(trivial ? 1.0 : den), !trivial);
    scale = cmake(-t * id, -dy * id);
    (void)live;
#else
    const double sgn = alpha.x >= 0.0 ? -1.0 : 1.0;
    beta = trivial ? alpha.x : sgn * (n2 * inrm);
    const double ib = sgn * inrm;
    tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
    const double dx = alpha.x - beta, dy = alpha.y;
    const double den = fma(dx, dx, dy * dy);
    const double id = pair_rcp(trivial ? 1.0 : den) * live;
    scale = cmake(dx * id, -dy * id);
#endif
}

// State carried from one step to the next: the reflector v (own rows), p = tau A v (own rows), the lane's share g
// of Re(y^H v), |tau|^2 and this lane's entry of e.  w' is finished at the top of the next step, under the p-half
// of the update.
struct PairCarry {
    cplx v[2], p[2];
    double g, tau2, eval;
};

// publish pivot column jn: X (masked to rows > jn+1), alpha = A[jn+1][jn], d[jn] = A[jn][jn].  Branch-free: the lanes of
// BOTH column parities store - each into its own parity's copy of X / alpha / D - and the readers take the copy of the
// parity that owns column jn (the other copy receives that parity's register of the same window slot: never read).
template <int LN, int NS>
SKTB_D void pair_publish(const cplx (&x)[2], const int jn, const PairLane& L, const PairBuf& S) {
    cplx* Xw = S.X + L.h * (2 * LN);
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        const int i = s ? L.iB : L.iA;
        const cplx xv = (i > jn + 1) ? x[s] : cmake(0.0, 0.0);
        Xw[i] = xv;
        if (i == jn + 1) S.alpha[L.h] = x[s];
        if (i == jn) S.D[L.h * LN + jn] = x[s].x;
    }
}

// reflector of step jn from the published column + y = A v partial sums z -> new carry; publishes v (parity 0) / p (parity 1)
template <int LN, int NS, bool VEC>
SKTB_D void pair_finish(PairCarry& c, const cplx (&z)[2], const cplx (&acol)[2], const bool a_here, const cplx (&xo)[2], const cplx alpha,
                        const double xn, const int jn, const PairLane& L, const PairBuf& S, cplx* Vnext) {
    double beta; cplx tau, scale;
    pair_larfg(alpha, xn, beta, tau, scale);
    c.eval = (L.l == jn) ? beta : c.eval;
    c.tau2 = fma(tau.x, tau.x, tau.y * tau.y);
    cplx y[2];
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
#if SKTB_RCP3
        cplx yp = cmake(a_here ? acol[s].x : 0.0, a_here ? acol[s].y : 0.0);     // the column where v' = 1 enters as the addend of the product
        cfma(yp, scale, z[s]);
#else
        cplx yp = cmul(scale, z[s]);
        yp.x += a_here ? acol[s].x : 0.0; yp.y += a_here ? acol[s].y : 0.0;
#endif
        y[s] = yp;
    }
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        y[s].x += __shfl_xor_sync(0xffffffffu, y[s].x, 1);
        y[s].y += __shfl_xor_sync(0xffffffffu, y[s].y, 1);
    }
    double g = 0.0;
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        const int i = s ? L.iB : L.iA;
        if (i <= jn) y[s] = cmake(0.0, 0.0);
        c.p[s] = cmul(tau, y[s]);
        c.v[s] = (i == jn + 1) ? cmake(1.0, 0.0) : cmul(xo[s], scale);
        g = fma(y[s].x, c.v[s].x, g); g = fma(y[s].y, c.v[s].y, g);
    }
    c.g = g;
    cplx* dst = L.h ? S.P : Vnext;
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) dst[s ? L.iB : L.iA] = L.h ? c.p[s] : c.v[s];
    if (VEC && jn < LN - 1) {      // eigenvector path: keep reflector jn (rows of the live slots) and tau; the no-op
                                     // step LN-1 at the end of the last phase must not write (it would land in tau[])
        if (L.h == 0) {
#pragma unroll
            for (int s = 2 - NS; s < 2; ++s) S.vst[(size_t)(jn + S.voff) * S.vstride + S.voff + (s ? L.iB : L.iA)] = c.v[s];
        }
        if (L.l == 0) S.taust[jn + S.voff] = tau;
    }
    __syncwarp();
}

// One Householder step j.  EVEN: j == base (pivot column in register 0 of parity-0 lanes, next pivot in register 0
// of parity-1 lanes); ODD: j == base + 1 (register 0 dead, next pivot in register 1 of parity-0 lanes) and every
// result is written one register down (the window rotation).  HALF = live registers per slot, NS = live slots.
template <int LN, int HALF, int NS, bool EVEN, bool VEC>
SKTB_D void pair_step(cplx (&B)[2][LN / 2], PairCarry& c, const int j, const PairLane& L, const PairBuf& S) {
    constexpr int ME = EVEN ? 0 : 1;          // register of the next pivot column
    constexpr int SH = EVEN ? 0 : 1;          // destination shift
    const int base = EVEN ? j : j - 1;
    constexpr int PH = EVEN ? 1 : 0;          // parity of the lanes that own the next pivot column jn
    const cplx* Pl = S.P + base + L.h;
    const cplx* Vl = (EVEN ? S.V0 : S.V1) + base + L.h;
    const cplx* Xs = S.X + PH * (2 * LN);
    const cplx* Xl = Xs + base + L.h;
    const int jn = j + 1;

    // ---- reduction for w': one butterfly stage per slice of the p-half of the update, in SOURCE order (ptxas keeps the
    // rough order of independent work: with the reduction written first its DADDs - each waiting ~30 cycles for a shuffle
    // - were issued ahead of the DFMA stream and stalled the in-order warp at the top of every step) ----
    constexpr int NST = LN == 32 ? 4 : 3;     // xor 2, 4, 8 (, 16): lanes of equal column parity within the matrix
#if SKTB_PAIR_EARLY
    // EARLY PUBLISH: only the first K1 columns of the p-half run before the new pivot column is finished and published (the butterfly
    // gets one stage per column), so that the STS -> LDS round trip of the publish, the norm reduction and the zlarfg chain all hide
    // under the REST of the p-half instead of opening a bubble between the p-half and the v-half.
    constexpr int K1 = (HALF - ME) < SKTB_PAIR_EARLY ? (HALF - ME) : SKTB_PAIR_EARLY;
    double G = c.g;
#pragma unroll
    for (int m = ME; m < ME + K1; ++m) {
        const cplx pq = Pl[2 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) pair_upd1(B[s][m], c.v[s], pq);
        const int st = m - ME + 1;
        if (st <= NST) G += __shfl_xor_sync(0xffffffffu, G, 1 << st);
    }
#pragma unroll
    for (int st = K1 + 1; st <= NST; ++st) G += __shfl_xor_sync(0xffffffffu, G, 1 << st);
#else
    constexpr int CH = (HALF - ME + NST) / (NST + 1) > 0 ? (HALF - ME + NST) / (NST + 1) : 1;
    double G = c.g;
#pragma unroll
    for (int m = ME; m < HALF; ++m) {
        const cplx pq = Pl[2 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) pair_upd1(B[s][m], c.v[s], pq);
        const int done = m - ME + 1;
        if (done % CH == 0 && done / CH <= NST) G += __shfl_xor_sync(0xffffffffu, G, 1 << (done / CH));
    }
#pragma unroll
    for (int st = (HALF - ME) / CH + 1; st <= NST; ++st) G += __shfl_xor_sync(0xffffffffu, G, 1 << st);   // stages the loop was too short for
#endif
    const double coef = -c.tau2 * G;
    cplx wp[2], x[2];
    {
        const cplx vq = Vl[2 * ME];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) {
            wp[s] = cmake(fma(coef, c.v[s].x, c.p[s].x), fma(coef, c.v[s].y, c.p[s].y));
            pair_upd1(B[s][ME], wp[s], vq);
            x[s] = B[s][ME];
        }
    }
    pair_publish<LN, NS>(x, jn, L, S);
    __syncwarp();
    cplx xo[2];
    double xnp = 0.0;
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) {
        xo[s] = Xs[s ? L.iB : L.iA];
        xnp = fma(xo[s].x, xo[s].x, xnp); xnp = fma(xo[s].y, xo[s].y, xnp);
    }
    const cplx alpha = S.alpha[PH];
    const double xn = pair_rsum<LN>(xnp);
#if SKTB_PAIR_EARLY
#pragma unroll
    for (int m = ME + K1; m < HALF; ++m) {     // rest of the p-half
        const cplx pq = Pl[2 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) pair_upd1(B[s][m], c.v[s], pq);
    }
#endif
    // ---- v-half on the remaining registers fused with the look-ahead mat-vec ----
    cplx z[2], acol[2];
    z[0] = z[1] = cmake(0.0, 0.0);
    if (!EVEN) {
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) B[s][0] = B[s][1];
    }
#pragma unroll
    for (int m = ME + 1; m < HALF; ++m) {
        const cplx vq = Vl[2 * m];
        const cplx xq = Xl[2 * m];
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) {
            cplx u = B[s][m];
            pair_upd1(u, wp[s], vq);
            cfma(z[s], u, xq);
            B[s][m - SH] = u;
        }
    }
    if (!EVEN) {
#pragma unroll
        for (int s = 2 - NS; s < 2; ++s) B[s][HALF - 1] = cmake(0.0, 0.0);
    }
    // column jn+1 (where v' = 1): EVEN -> register 1 of parity-0 lanes; ODD -> old register 1 (now 0) of parity-1 lanes
#pragma unroll
    for (int s = 2 - NS; s < 2; ++s) acol[s] = EVEN ? B[s][1] : B[s][0];
    pair_finish<LN, NS, VEC>(c, z, acol, L.h == (EVEN ? 0 : 1), xo, alpha, xn, jn, L, S, EVEN ? S.V1 : S.V0);
}

// pivot column 0 (register 0 of parity-0 lanes) and the plain mat-vec for reflector 0
template <int LN, bool VEC>
SKTB_D void pair_prologue(cplx (&B)[2][LN / 2], PairCarry& c, const PairLane& L, const PairBuf& S) {
    cplx x[2] = {B[0][0], B[1][0]};
    pair_publish<LN, 2>(x, 0, L, S);           // column 0: parity-0 copy
    __syncwarp();
    cplx xo[2];
    double xnp = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        xo[s] = S.X[s ? L.iB : L.iA];
        xnp = fma(xo[s].x, xo[s].x, xnp); xnp = fma(xo[s].y, xo[s].y, xnp);
    }
    const cplx alpha = S.alpha[0];
    const double xn = pair_rsum<LN>(xnp);
    cplx z[2], acol[2];
    z[0] = z[1] = cmake(0.0, 0.0);
    const cplx* Xl = S.X + L.h;
#pragma unroll
    for (int m = 0; m < LN / 2; ++m) {
        const cplx xq = Xl[2 * m];
        cfma(z[0], B[0][m], xq); cfma(z[1], B[1][m], xq);
    }
    acol[0] = B[0][0]; acol[1] = B[1][0];                 // column 1 = register 0 of parity-1 lanes
    c.eval = 0.0;
    pair_finish<LN, 2, VEC>(c, z, acol, L.h == 1, xo, alpha, xn, 0, L, S, S.V0);
}

// phases of S steps: window HALF registers, then HALF - S/2, ... down to (excluding) HEND; slot A is dead once
// HALF <= LN/4 (step >= LN/2)
template <int LN, int HALF, int HEND, int S, bool VEC>
SKTB_D void pair_phases(cplx (&B)[2][LN / 2], PairCarry& c, int& j, const PairLane& L, const PairBuf& S_) {
    constexpr int NS = HALF > LN / 4 ? 2 : 1;
#pragma unroll 1
    for (int it = 0; it < S / 2; ++it) {
        pair_step<LN, HALF, NS, true, VEC>(B, c, j, L, S_);
        pair_step<LN, HALF, NS, false, VEC>(B, c, j + 1, L, S_);      // (the step LN-1 of the last phase is a harmless no-op)
        j += 2;
    }
    if constexpr (HALF - S / 2 > HEND) pair_phases<LN, HALF - S / 2, HEND, S, VEC>(B, c, j, L, S_);
}

// K1 in the pair layout: rows iA, iB x columns 2m + h of the mirrored-lower H(k).  Branch-free: table offsets of
// out-of-range orbitals are clamped to 0 and the garbage they accumulate is masked when the row is assembled.
template <bool SOC>
SKTB_D void pair_build(cplx (&B)[2][16], const SmallArgs& a, const SmemView& s, const cplx* phs, const PairLane& L) {
    const int n = a.t.n_orb, nb = a.t.n_bonds, N = a.N;
    if (SOC) {
        // kron(h, I2): row 2 mu + spin couples to column 2 nu + spin only; this lane's columns have spin h, and
        // exactly one of its two rows has that spin
        const bool selA = ((L.r & 1) == L.h);
        const int isel = selA ? L.iA : L.iB;
        const int mu = isel >> 1;
        const bool ok = isel < N && mu < n;
        const int mu_c = ok ? mu : 0;
        int off[16];
#pragma unroll
        for (int nu = 0; nu < 16; ++nu) off[nu] = nu < n ? nu * n : 0;
        double re[16], im[16];
#pragma unroll
        for (int nu = 0; nu < 16; ++nu) { re[nu] = 0.0; im[nu] = 0.0; }
#pragma unroll 2
        for (int b = 0; b < nb; ++b) {
            const cplx ph = phs[b];
            const double* Lb = s.Lt + (size_t)b * n * n + mu_c;
#pragma unroll
            for (int nu = 0; nu < 16; ++nu) {
                const double t = Lb[off[nu]];
                re[nu] = fma(t, ph.x, re[nu]); im[nu] = fma(t, ph.y, im[nu]);
            }
        }
        const double ons = s.On[mu_c];
        const int rA = min(L.iA, N - 1), rB = min(L.iB, N - 1);
#pragma unroll
        for (int nu = 0; nu < 16; ++nu) {
            const bool live = ok && nu < n;
            double r_ = live ? re[nu] : 0.0, i_ = live ? im[nu] : 0.0;
            if (nu == mu) { r_ += ok ? ons : 0.0; i_ = 0.0; }
            else if (nu > mu) i_ = -i_;
            const int col = 2 * nu + L.h;
            const int cc = min(col, N - 1);
            cplx sA = s.St[(size_t)cc * N + rA], sB = s.St[(size_t)cc * N + rB];
            if (!(col < N && L.iA < N)) sA = cmake(0.0, 0.0);
            if (!(col < N && L.iB < N)) sB = cmake(0.0, 0.0);
            B[0][nu] = cmake((selA ? r_ : 0.0) + sA.x, (selA ? i_ : 0.0) + sA.y);
            B[1][nu] = cmake((selA ? 0.0 : r_) + sB.x, (selA ? 0.0 : i_) + sB.y);
        }
    } else {
        int off[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) off[m] = (2 * m + L.h) < n ? (2 * m + L.h) * n : 0;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int mu = sl ? L.iB : L.iA;
            const bool ok = mu < N && mu < n;
            const int mu_c = ok ? mu : 0;
            double re[16], im[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) { re[m] = 0.0; im[m] = 0.0; }
#pragma unroll 2
            for (int b = 0; b < nb; ++b) {
                const cplx ph = phs[b];
                const double* Lb = s.Lt + (size_t)b * n * n + mu_c;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const double t = Lb[off[m]];
                    re[m] = fma(t, ph.x, re[m]); im[m] = fma(t, ph.y, im[m]);
                }
            }
            const double ons = s.On[mu_c];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int nu = 2 * m + L.h;
                const bool live = ok && nu < n;
                double r_ = live ? re[m] : 0.0, i_ = live ? im[m] : 0.0;
                if (nu == mu) { r_ += ok ? ons : 0.0; i_ = 0.0; }
                else if (nu > mu) i_ = -i_;
                B[sl][m] = cmake(r_, i_);
            }
        }
    }
}

// WIDE kernel: K0 + K1 + steps 0..15 of the 32x32 tridiagonalisation.  Writes d[0..16], e[0..16] into the (d,e)
// scratch and the trailing 16x16 block (updated through step 15, column-major [q][i], 4 KB) into `blocks`.
// FULL: run all 31 steps here (no tail kernel).
// rows iA, iB x columns 2m + h of a caller-supplied matrix, zheevd(UPLO='L') reading (upper triangle = conj(lower))
SKTB_D void pair_load(cplx (&B)[2][16], const cplx* __restrict__ H, const int N, const int ldh, const PairLane& L) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = s ? L.iB : L.iA;
#pragma unroll
        for (int m = 0; m < 16; ++m) {                    // branch-free (clamped address): 32 independent loads
            const int col = 2 * m + L.h;
            const bool ok = i < N && col < N;
            const int hi = ok ? max(i, col) : 0, lo = ok ? min(i, col) : 0;
            cplx e = H[(size_t)hi * ldh + lo];
            if (col > i) e.y = -e.y;
            if (col == i) e.y = 0.0;
            if (!ok) e = cmake(0.0, 0.0);
            B[s][m] = e;
        }
    }
}

constexpr int PAIR_SRC_NOSOC = 0, PAIR_SRC_SOC = 1, PAIR_SRC_LOAD = 2;

template <int SRC, int S, bool FULL, bool VEC = false>
__global__ void __launch_bounds__(PAIR_WARPS * 32, 2) pair_wide32_kernel(SmallArgs a, cplx* __restrict__ blocks, cplx* __restrict__ vstore) {
    constexpr int NC = 32;
    constexpr bool SOC = SRC == PAIR_SRC_SOC, LOAD = SRC == PAIR_SRC_LOAD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = LOAD ? 0 : a.t.n_bonds;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SmemView s{};
    if (LOAD) s.warp_base = reinterpret_cast<cplx*>(smem_raw);
    else s = carve_and_load<SOC, false>(smem_raw, a);
    const size_t per_warp_cplx = (size_t)pair_buf_cplx<32>() + (size_t)nb;
    cplx* myw = s.warp_base + per_warp_cplx * warp;
    cplx* phs = myw + pair_buf_cplx<32>();
    for (int t = lane; t < pair_buf_cplx<32>(); t += 32) myw[t] = cmake(0.0, 0.0);
    __syncthreads();
    PairBuf S_ = pair_buf<32>(myw);
    const PairLane L = pair_lane<32>(lane);

#pragma unroll 1
    for (long long kcta = (long long)blockIdx.x * PAIR_WARPS; kcta < a.nk; kcta += (long long)gridDim.x * PAIR_WARPS) {
#if SKTB_SMALL_SYNC
        __syncthreads();
#endif
        long long kq = kcta + warp;
        const bool k_ok = kq < a.nk;
        if (!k_ok) kq = a.nk - 1;
        if (VEC) {
            S_.vst = vstore + (size_t)kq * (a.v_per_k ? a.v_per_k : PAIR_VSTORE_CPLX);
            S_.taust = S_.vst + (a.v_tau ? a.v_tau : 1024);
            S_.vstride = a.v_stride ? a.v_stride : 32; S_.voff = a.v_off;
        }
        cplx B[2][16];
        if (LOAD) pair_load(B, a.Hsrc + (size_t)kq * (a.Hld ? a.Hld * a.Hld : a.N * a.N), a.N, a.Hld ? a.Hld : a.N, L);
        else {
            bond_phases<32>(a, s, kq, lane, phs);
            __syncwarp();
            pair_build<SOC>(B, a, s, phs, L);
        }
        if (a.flags && lane == 0 && k_ok) a.flags[a.flag_off + kq] = 0;
        PairCarry c;
        pair_prologue<32, VEC>(B, c, L, S_);
        int j = 0;
        pair_phases<32, 16, FULL ? 0 : 8, S, VEC>(B, c, j, L, S_);
        if (k_ok) {
            double* out = a.scratch + (size_t)kq * (a.out_stride ? a.out_stride : 2 * NC);
            out[a.d_off + lane] = S_.D[(lane & 1) * 32 + lane]; out[(a.out_stride ? a.e_off : NC) + lane] = c.eval;
            if (!FULL) {
                cplx* blk = blocks + (size_t)kq * 256 + (15 - L.r);
#pragma unroll
                for (int m = 0; m < 8; ++m) blk[(2 * m + L.h) * 16] = B[1][m];
            }
        }
        __syncwarp();
    }
}

// TAIL kernel: tridiagonalisation of the trailing 16x16 blocks (steps 16..30), two matrices per warp, d[16..31] and
// e[16..30] into the (d,e) scratch.  Few registers: co-resides with the wide kernel of the next chunk.
// STOP8 (eigenvalue path): run only the two-slot phases (steps 16..23) and hand the trailing 8 x 8 block (updated through
// step 23, column-major [q][i], 1 KB - written over the head of this matrix's own 4 KB input block) to tail8_thread_kernel:
// the last 7 steps are all fixed per-step cost in a lane-group layout, and none in a one-matrix-per-thread layout.
template <int S, bool VEC = false, bool STOP8 = false>
__global__ void __launch_bounds__(PAIR_WARPS * 32, 4) pair_tail16_kernel(const cplx* __restrict__ blocks, double* __restrict__ scratch, long long nk,
                                                                          int out_stride, int d_off, int e_off, cplx* __restrict__ vstore = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 4;
    cplx* myw = reinterpret_cast<cplx*>(smem_raw) + (size_t)(2 * warp + g) * pair_buf_cplx<16>();
    for (int t = lane & 15; t < pair_buf_cplx<16>(); t += 16) myw[t] = cmake(0.0, 0.0);
    __syncwarp();
    PairBuf S_ = pair_buf<16>(myw);
    S_.voff = 16;                 // eigenvector path: reflectors 16..30 / rows 16..31 of the [32][32] store
    const PairLane L = pair_lane<16>(lane);
    const long long n_pairs = (nk + 1) / 2;
#pragma unroll 1
    for (long long pcta = (long long)blockIdx.x * PAIR_WARPS; pcta < n_pairs; pcta += (long long)gridDim.x * PAIR_WARPS) {
#if SKTB_SMALL_SYNC
        __syncthreads();          // keep the CTA's warps on the same stretch of the instruction stream (L1.5 I-cache is 32 KB)
#endif
        const long long pr = pcta + warp;
        long long kq = 2 * pr + g;
        const bool k_ok = kq < nk;
        if (!k_ok) kq = nk - 1;
        const cplx* blk = blocks + (size_t)kq * 256;
        cplx B[2][8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            B[0][m] = blk[(2 * m + L.h) * 16 + L.iA];
            B[1][m] = blk[(2 * m + L.h) * 16 + L.iB];
        }
        if (VEC) { S_.vst = vstore + (size_t)kq * PAIR_VSTORE_CPLX; S_.taust = S_.vst + 1024; }
        PairCarry c;
        pair_prologue<16, VEC>(B, c, L, S_);
        int j = 0;
        pair_phases<16, 8, STOP8 ? 4 : 0, S, VEC>(B, c, j, L, S_);
        if (k_ok) {
            double* out = scratch + (size_t)kq * out_stride;
            out[d_off + 16 + L.l] = S_.D[(L.l & 1) * 16 + L.l]; out[e_off + 16 + L.l] = c.eval;      // STOP8: entries 8.. are rewritten by tail8
            if (STOP8 && L.r < 8) {       // slot B = rows 15 - r = 8..15; registers 0..3 hold columns 8 + 2m + h (base = 8 after 8 steps)
                cplx* b8 = const_cast<cplx*>(blk) + (7 - L.r);
#pragma unroll
                for (int m = 0; m < 4; ++m) b8[(2 * m + L.h) * 8] = B[1][m];
            }
        }
        __syncwarp();
    }
}

// Steps 24..30 of the 32 x 32 reduction: one trailing 8 x 8 block per THREAD (thread_tridiag<8>, heev_thread.cuh): lower
// triangle of the block pair_tail16<STOP8> left, d[24..31] and e[24..30] into the (d, e) record.
template <int UNUSED>
__global__ void __launch_bounds__(128) tail8_thread_kernel(const cplx* __restrict__ blocks, double* __restrict__ scratch, long long nk, int out_stride,
                                                           int d_off, int e_off) {
    for (long long kq = (long long)blockIdx.x * blockDim.x + threadIdx.x; kq < nk; kq += (long long)gridDim.x * blockDim.x) {
        const cplx* blk = blocks + (size_t)kq * 256;
        cplx A[36];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int i = j; i < 8; ++i) A[i * (i + 1) / 2 + j] = blk[j * 8 + i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) A[i * (i + 1) / 2 + i].y = 0.0;
        double d[8], e[8];
        thread_tridiag<8>(A, d, e);
        double* out = scratch + (size_t)kq * out_stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) { out[d_off + 24 + i] = d[i]; out[e_off + 24 + i] = e[i]; }
    }
}

