// Register-resident Cholesky of the small dense reduced camera systems (LocalBundleAdjustment's 6 Kv <= 152 unknowns,
// CeresOptimizer.cc:508-524; the nested-dissection nodes of GlobalBundleAdjustemnt, band_cr.cuh).  Included by ba.cu after
// dmma_884 / kSolveThreads.
//
// Why a third version.  packed_cholesky (v2) keeps the trailing matrix in shared memory: every rank-6 update reads and
// writes the whole remaining triangle through shared memory (a warp-wide fp64 access is two wavefronts) and the pivot chain
// is one thread factoring a 6 x 6 block (~150 dependent fp64 operations): 2900 cycles per panel, 46 us per 114 unknowns.
// Here the trailing matrix never leaves the register file:
//   * the matrix (rows 0..n-1, padded with identity to a multiple of 8, then ONE tile row that carries the rhs) is cut
//     into 8 x 8 tiles in the accumulator layout of mma.m8n8k4.f64; warp i holds tile row i (2 x 16 doubles per lane);
//   * panel k: the diagonal tile is factored WITH SHUFFLES by the warp that owns it (8 steps of: broadcast the pivot, one
//     reciprocal, one FMA — the elimination runs on unscaled columns; the inverse of the Cholesky block is accumulated
//     beside it from the identity, off the critical chain) and its inverse M published; every tile (i, k) below becomes
//     L_ik = A_ik M' by two DMMA (its accumulator turned into an A fragment by four shuffles) and goes to shared memory as
//     the panel; every remaining tile takes C_ij -= L_ik L_jk' by two DMMA straight into its registers — while the warp of
//     row k + 1, whose last tile is its diagonal tile, is already factoring it.  Two named barriers per panel (panel buffer
//     and M are double buffered), no shared-memory read-modify-write at all.  Measured on B200 (tools/micro/lat.cu): DFMA 8,
//     double shuffle 26, DMMA 26 cycles of latency, one DMMA per 16 cycles and scheduler, __syncthreads of 512 threads 45.
// The result has the format the substitutions expect: packed lower triangle in shared memory, y in row n, the 24 x 24
// diagonal blocks inverted (invert_diag24_r8 merges the 8 x 8 inverses: 8 -> 24 in two dependent products).
#pragma once

namespace cmos {

constexpr int kCholRegMaxT = 16;        // tile rows incl. the rhs row (one warp each): n <= 120
__host__ __device__ inline int chol_reg_tiles(int n) { return ((n + 7) >> 3) + 1; }
__host__ __device__ inline bool chol_reg_supported(int n) { return n >= 8 && chol_reg_tiles(n) <= kCholRegMaxT; }
// row pitch of the column-major panel buffer: 4 mod 16 doubles, so the four k-columns of a fragment load fall in distinct banks
__host__ __device__ inline int chol_reg_pitch(int n) { return ((8 * chol_reg_tiles(n) + 11) / 16) * 16 + 4; }
constexpr int kCholMPitch = 12;         // pitch of the 8 x 8 inverse in shared memory (bank-conflict-free fragment loads)
// doubles of scratch behind the packed triangle: two panel buffers, two inverses, one hand-over tile per warp (and T1..T3 of
// invert_diag24_r8 reuse it)
__host__ __device__ inline size_t chol_reg_scratch(int n) { return (size_t)16 * chol_reg_pitch(n) + (2 + kCholRegMaxT) * 8 * kCholMPitch; }

__device__ __forceinline__ double shfl_d(const double v, const int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1 / d to fp64 accuracy without the IEEE slow path: MUFU seed (2^-23) + two Newton steps
__device__ __forceinline__ double rcp_newton(const double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}

// One warp factors a symmetric 8 x 8 tile held in the mma accumulator layout (lane (g = lane / 4, q = lane % 4) holds
// C[g][2q], C[g][2q + 1]) and returns M = inv(chol(C)) in the same layout.  false: a pivot was not positive / finite.
// Straight-line code, no branches: the elimination runs unscaled (C[g][c] -= C[g][j] C[j][c] / d_j — the loop-carried
// chain is one double shuffle, one reciprocal (MUFU seed + two Newton steps) and one FMA per pivot: measured 26 + 52 + 8
// cycles on B200), the inverse is accumulated beside it from the identity with the same multipliers (W = inv(L_unit)),
// and the square roots come once, after the loop: M = diag(d)^-1/2 W.
__device__ __forceinline__ bool chol8_warp(double c0, double c1, double& m0, double& m1, const int lane) {
  const int g = lane >> 2, q = lane & 3;
  double w0 = g == 2 * q ? 1.0 : 0.0, w1 = g == 2 * q + 1 ? 1.0 : 0.0;
  double dg = 1.0;                                                 // pivot of this lane's row
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double cs = (j & 1) ? c1 : c0;
    const double d = shfl_d(cs, j * 4 + (j >> 1));                 // pivot C[j][j]
    const double cgj = shfl_d(cs, (lane & ~3) | (j >> 1));         // C[g][j]
    const int sr = j * 4 + q;
    const double cj0 = shfl_d(c0, sr), cj1 = shfl_d(c1, sr);       // C[j][2q], C[j][2q + 1] (= C[2q][j], ... by symmetry)
    const double wj0 = shfl_d(w0, sr), wj1 = shfl_d(w1, sr);       // row j of W, final since step j - 1
    ok = ok && (d > 0.0) && isfinite(d);
    const double t0 = cgj * cj0, t1 = cgj * cj1;
    const double inv = rcp_newton(d);
    c0 = fma(-t0, inv, c0);                                        // (rows / columns <= j are dead from here on)
    c1 = fma(-t1, inv, c1);
    const double f = g > j ? cgj * inv : 0.0;                      // rows <= j of W are final
    w0 = fma(-f, wj0, w0);
    w1 = fma(-f, wj1, w1);
    dg = g == j ? d : dg;
  }
  const double r = rsqrt(dg);
  m0 = w0 * r; m1 = w1 * r;
  return ok;
}

// Packed triangle L (rows 0..n, row n = rhs) in shared memory -> Cholesky factor in place: tiles below the diagonal tiles
// hold L, the 8 x 8 diagonal tiles hold the INVERSE of their Cholesky block, row n holds y = inv(L) rhs.  S: scratch of
// chol_reg_scratch(n) doubles.  All kSolveThreads threads; the caller synchronises before (L complete) and after.
//
// Ownership: warp i holds TILE ROW i (tiles (i, 0..i), slot = tile column — every slot index in the update loop is a
// compile-time constant, shared-memory offsets are immediates; the first version dealt tiles round-robin and spent 500
// instructions per warp and panel on decoding which slot is what: with four warps per scheduler that was the limit, not
// any latency).  Panel k:
//   warps i > k wait for M_k (named barrier 1), turn their tile (i, k) into L_ik = A_ik M_k' (two DMMA), store it as row
//   block i of the panel; then warp k + 1 — whose only remaining tile is its diagonal tile — updates that tile from its own
//   L_{k+1,k} (no barrier: nobody else's data), factors it and publishes M_{k+1}: it is the pivot chain, one panel ahead of
//   the others, and retires.  The other warps meet at named barrier 2 (the panel is complete) and update tiles
//   (i, k+1..i) with two DMMA each, A fragment = their own row block, B fragment = row block j of the panel.
// Barrier counts shrink with the rows still alive.  Critical path per panel: panel product of one tile, diagonal update,
// one 8 x 8 factorisation.
__device__ __forceinline__ void packed_cholesky_reg(double* __restrict__ L, double* __restrict__ S, const int n, int* s_fail) {
  constexpr int NS = kCholRegMaxT - 1;           // slots: tile columns 0..14 (strictly below the diagonal)
  const int tid = threadIdx.x, lane = tid & 31, i = tid >> 5, g = lane >> 2, q = lane & 3;
  const int np8 = (n + 7) & ~7, T = (np8 >> 3) + 1, psr = chol_reg_pitch(n);
  double* const Pb = S;                                     // [2][8][psr] panel, column-major
  double* const Mb = S + 16 * psr;                          // [2][8][kCholMPitch] inverse of the diagonal Cholesky tile
  double* const Nb = Mb + 2 * 8 * kCholMPitch;              // [16][8][kCholMPitch] per warp: its tile of the NEXT panel column
  const int rhs_row = n * (n + 1) / 2;
  const bool is_rhs = i == T - 1;
  double c0[NS], c1[NS];                                    // tiles (i, j), j < i
  double dd0 = 0.0, dd1 = 0.0;                              // tile (i, i)
  double* const mine = Nb + i * 8 * kCholMPitch;
  if (i < T) {
    auto at = [&](const int r, const int c) -> double {
      if (r == np8) return c < n ? L[rhs_row + c] : 0.0;
      if (r > np8) return 0.0;
      if (r >= n || c >= n) return r == c ? 1.0 : 0.0;
      const int hi = max(r, c), lo = min(r, c);
      return L[hi * (hi + 1) / 2 + lo];
    };
#pragma unroll
    for (int j = 0; j < NS; j++) {
      c0[j] = 0.0; c1[j] = 0.0;
      if (j < i) { c0[j] = at(8 * i + g, 8 * j + 2 * q); c1[j] = at(8 * i + g, 8 * j + 2 * q + 1); }
    }
    if (!is_rhs) { dd0 = at(8 * i + g, 8 * i + 2 * q); dd1 = at(8 * i + g, 8 * i + 2 * q + 1); }
    if (i > 0) { mine[g * kCholMPitch + 2 * q] = c0[0]; mine[g * kCholMPitch + 2 * q + 1] = c1[0]; }
  }
  __syncthreads();
  if (i >= T) return;
#ifdef CMOS_CHOL_TIMING
  long long pk[6] = {0, 0, 0, 0, 0, 0}, tp;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(tp) :: "memory");
#define CPK(x) { long long tn; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tn) :: "memory"); pk[x] += tn - tp; tp = tn; }
#else
#define CPK(x)
#endif
  // factor diagonal tile k (this warp's), publish M_k and the packed copy; signals barrier 1 (T - k warps take part)
  auto factor_publish = [&](const int k) {
    double m0, m1;
    const bool ok = chol8_warp(dd0, dd1, m0, m1, lane);
    double* const M = Mb + (k & 1) * 8 * kCholMPitch;
    M[g * kCholMPitch + 2 * q] = m0; M[g * kCholMPitch + 2 * q + 1] = m1;
    const int R = 8 * k + g, C = 8 * k + 2 * q;
    if (R < n) {
      if (C <= R) L[R * (R + 1) / 2 + C] = m0;
      if (C + 1 <= R) L[R * (R + 1) / 2 + C + 1] = m1;
    }
    if (!ok && lane == 0) *s_fail = 1;
    __threadfence_block();
    __syncwarp();
    asm volatile("bar.arrive 1, %0;" ::"r"(32 * (T - k)) : "memory");
  };
  if (i == 0) {
    factor_publish(0);
  } else {
    const int R = 8 * i + g;                                        // this lane's row of the padded matrix
    const int pr = R == np8 ? n : (R < n ? R : -1);                 // ... and of the packed triangle
    double* const Lrow = L + (pr >= 0 ? pr * (pr + 1) / 2 : 0);
    const int kend = is_rhs ? T - 1 : i;                            // panels 0..kend-1 reach this row
    for (int k = 0; k < kend; k++) {
      double* const P = Pb + (k & 1) * 8 * psr;
      const double* const M = Mb + (k & 1) * 8 * kCholMPitch;
      // the tile of this panel column was left in `mine` by the previous update, in A-fragment order
      const double a0 = mine[g * kCholMPitch + q], a1 = mine[g * kCholMPitch + 4 + q];
      asm volatile("bar.sync 1, %0;" ::"r"(32 * (T - k)) : "memory");           // M_k is there
      CPK(0)
      // ---- panel product: L_ik = A_ik M_k'
      const double b0 = M[g * kCholMPitch + q], b1 = M[g * kCholMPitch + 4 + q];
      double x0 = 0.0, x1 = 0.0;
      dmma_884(x0, x1, a0, b0);
      dmma_884(x0, x1, a1, b1);
      P[(2 * q) * psr + R] = x0; P[(2 * q + 1) * psr + R] = x1;
      __syncwarp();
      const double* const pa = P + q * psr + g;                       // fragment (row block r): pa[8 r], pa[8 r + 4 psr]
      const double la0 = pa[8 * i], la1 = pa[8 * i + 4 * psr];        // own row block, as A (and B) fragment
      if (k == i - 1 && !is_rhs) {
        // ---- the pivot chain: my diagonal tile is the next panel's
        __threadfence_block();
        asm volatile("bar.arrive 2, %0;" ::"r"(32 * (T - 1 - k)) : "memory");
        dmma_884(dd0, dd1, -la0, la0);
        dmma_884(dd0, dd1, -la1, la1);
        CPK(2)
        factor_publish(i);
        CPK(3)
      }
      if (pr >= 0) {
        const int C = 8 * k + 2 * q;
        if (C < n) Lrow[C] = x0;
        if (C + 1 < n) Lrow[C + 1] = x1;
      }
      CPK(1)
      if (k == i - 1 && !is_rhs) break;
      asm volatile("bar.sync 2, %0;" ::"r"(32 * (T - 1 - k)) : "memory");       // the whole panel is in P
      CPK(4)
      // ---- trailing update in registers: C_ij -= L_ik L_jk', j = k+1..i-1, and the diagonal tile; four tiles per batch:
      // their eight operand loads first, then eight DMMA back to back
      const double na0 = -la0, na1 = -la1;
#pragma unroll
      for (int jb = 1; jb < NS; jb += 4) {
        if (jb + 3 <= k || jb >= i) continue;                          // no tile of this batch is live (warp-uniform)
        double f0[4], f1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int j = jb + u < NS ? jb + u : NS - 1; f0[u] = pa[8 * j]; f1[u] = pa[8 * j + 4 * psr]; }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int j = jb + u; if (j < NS && j > k && j < i) dmma_884(c0[j], c1[j], na0, f0[u]); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = jb + u;
          if (j < NS && j > k && j < i) {
            dmma_884(c0[j], c1[j], na1, f1[u]);
            if (j == k + 1) { mine[g * kCholMPitch + 2 * q] = c0[j]; mine[g * kCholMPitch + 2 * q + 1] = c1[j]; }
          }
        }
      }
      if (!is_rhs) { dmma_884(dd0, dd1, na0, la0); dmma_884(dd0, dd1, na1, la1); }
      __syncwarp();
      CPK(5)
    }
  }
#ifdef CMOS_CHOL_TIMING
  if (lane == 0 && (i == 1 || i == 7 || i == 14 || i == 15) && blockIdx.x == 0)
    printf("packed_cholesky_reg n %d T %d row %d: wait M %lld | panel product %lld | own diagonal %lld | factor+publish %lld | wait panel %lld | update %lld\n",
           n, T, i, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5]);
#endif
#undef CPK
}

// After packed_cholesky_reg: the 24 x 24 diagonal blocks [[A,0,0],[B,C,0],[D,E,F]] become their inverses,
// M10 = -M11 (B M00), M21 = -M22 (E M11), M20 = -M22 (D M00 + E M10); the 8 x 8 diagonal tiles are already inverted.  The
// last block may be 6, 12 or 18 rows.  Tb: scratch of 192 ceil(n / 24) doubles.  All kSolveThreads threads.
__device__ __forceinline__ void invert_diag24_r8(double* __restrict__ L, double* __restrict__ Tb, const int n) {
  const int tid = threadIdx.x, nb = (n + 23) / 24;
  auto el = [&](const int r, const int c) -> double& { return L[r * (r + 1) / 2 + c]; };
  // phase 1: T1 = L10 M00, T2 = L21 M11
  for (int e = tid; e < nb * 128; e += kSolveThreads) {
    const int p = e >> 7, w = (e >> 6) & 1, r = (e >> 3) & 7, j = e & 7;
    const int co = 24 * p + 8 * w, ro = co + 8 + r;
    if (ro >= n) continue;
    double v = 0.0;
    for (int k = j; k < 8; k++) v += el(ro, co + k) * el(co + k, co + j);
    Tb[p * 192 + w * 64 + r * 8 + j] = v;
  }
  __syncthreads();
  // phase 2: M10 = -M11 T1 (over L10)
  for (int e = tid; e < nb * 64; e += kSolveThreads) {
    const int p = e >> 6, r = (e >> 3) & 7, j = e & 7;
    const int a = 24 * p, ro = a + 8 + r;
    if (ro >= n) continue;
    double v = 0.0;
    for (int k = 0; k <= r; k++) v += el(ro, a + 8 + k) * Tb[p * 192 + k * 8 + j];
    el(ro, a + j) = -v;
  }
  __syncthreads();
  // phase 3: T3 = L20 M00 + L21 M10
  for (int e = tid; e < nb * 64; e += kSolveThreads) {
    const int p = e >> 6, r = (e >> 3) & 7, j = e & 7;
    const int a = 24 * p, ro = a + 16 + r;
    if (ro >= n) continue;
    double v = 0.0;
    for (int k = j; k < 8; k++) v += el(ro, a + k) * el(a + k, a + j);
    for (int k = 0; k < 8; k++) v += el(ro, a + 8 + k) * el(a + 8 + k, a + j);
    Tb[p * 192 + 128 + r * 8 + j] = v;
  }
  __syncthreads();
  // phase 4: M21 = -M22 T2 (over L21), M20 = -M22 T3 (over L20)
  for (int e = tid; e < nb * 128; e += kSolveThreads) {
    const int p = e >> 7, w = (e >> 6) & 1, r = (e >> 3) & 7, j = e & 7;
    const int a = 24 * p, ro = a + 16 + r;
    if (ro >= n) continue;
    const double* Tq = Tb + p * 192 + (w ? 64 : 128);
    double v = 0.0;
    for (int k = 0; k <= r; k++) v += el(ro, a + 16 + k) * Tq[k * 8 + j];
    el(ro, a + (w ? 8 : 0) + j) = -v;
  }
  __syncthreads();
}

}  // namespace cmos
