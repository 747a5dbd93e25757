// Parallel Cholesky of the block-banded reduced camera system (GlobalBundleAdjustemnt, CeresOptimizer.cc:178-187 — the
// 6K x 6K system Ceres hands to CHOLMOD): nested dissection by BLOCK CYCLIC REDUCTION instead of one CTA walking down
// 1000 dependent block columns.
//
// The Kv keyframe blocks are grouped into N nodes of Wb >= W consecutive blocks (W = block half-bandwidth), so S is block
// tridiagonal in nodes of n = 6 Wb unknowns.  Level l = 1, 2, ... eliminates the nodes i = 2^(l-1) (2t + 1) (1-based) in
// parallel; the neighbours of i at that level are l = i - 2^(l-1) and r = i + 2^(l-1).  Eliminating i:
//     D_i = L L',  y_i = L^-1 b_i,  V_l = L^-1 S_il,  V_r = L^-1 S_ir                       (k_cr_factor, k_cr_spike)
//     S_ll -= V_l' V_l,  S_rr -= V_r' V_r,  S_lr = -V_l' V_r,  b_l -= V_l' y_i,  b_r -= V_r' y_i   (k_cr_schur)
// and, from the last level back to the first,  x_i = L^-T (y_i - V_l x_l - V_r x_r)            (k_cr_back).
// L^-1 is never formed: V = L^-1 S and x = L^-T z are blocked substitutions against the packed factor (a first version
// inverted L explicitly by pairwise merging; that cost 88 k cycles per node, more than the factorisation itself).
// This is a Cholesky factorisation in nested-dissection order: the pivot chain is log2(N) dense n x n factorisations
// (6 x 120 pivots at configs[4]) instead of 6 Kv = 6000.  Every accumulation has one writer and a fixed order, so repeated
// solves are bit-identical.  The dense products run on the fp64 tensor cores (mma.sync m8n8k4, SASS DMMA).
//
// Node storage (in the allocation of the dense fallback S): arrays of N x n x n doubles
//   D0    node diagonal block from Sblk (lower triangle used)
//   AccL  sum of V_r' V_r of the nodes eliminated on i's LEFT  (i was their right neighbour)
//   AccR  sum of V_l' V_l of the nodes eliminated on i's RIGHT (i was their left neighbour)
//   Ep    S_{i,i-1} from Sblk (rows of i, columns of i-1)
//   Lp    the packed factor of D_i: rows 0..n-1 at r (r + 1) / 2; blocks below the 24 x 24 diagonal blocks = L, the
//         24 x 24 diagonal blocks = the INVERSE of their Cholesky block (substitutions are tile products, not divisions)
//   Vl, Vr, Clr = V_l' V_r (rows of l, columns of r)
// and vectors of N x n: bL, bR (like AccL / AccR), y, x.
#pragma once

namespace cmos {

struct CrArgs {
  int n, Wb, N, W, levels;
  const int* band_blk;      // [Kv][W+1]: id of block (b - off, b), or -1
  double* base;
  // BORDERED systems (band + arrow: the pose graph of OptimizeEssentialGraph, whose loop edges tie a few keyframes to
  // far-away ones — those keyframes are ordered last and form the border; posegraph.cuh).  nbp = 0: no border.
  int nbp;                  // border unknowns, padded to a multiple of 24
  double* bbase;            // border storage (crb_* below)
  const double* rhs_nodes;  // [N][n] right-hand side by node (null: BaDev::rhs in 6-wide camera blocks)
  double* x_nodes;          // [N][n] solution by node (null: BaDev::yc)
};
enum { CR_D0 = 0, CR_ACCL, CR_ACCR, CR_EP, CR_LP, CR_VL, CR_VR, CR_CLR, CR_NARR };
enum { CR_BL = 0, CR_BR, CR_Y, CR_X, CR_NVEC };

__host__ __device__ inline size_t cr_doubles(int N, int n) { return (size_t)N * n * ((size_t)CR_NARR * n + CR_NVEC); }
__device__ __forceinline__ double* cr_arr(const CrArgs& a, int k, int node) {      // node is 1-based
  return a.base + ((size_t)k * a.N + (node - 1)) * a.n * a.n;
}
__device__ __forceinline__ double* cr_vec(const CrArgs& a, int k, int node) {
  return a.base + (size_t)CR_NARR * a.N * a.n * a.n + ((size_t)k * a.N + (node - 1)) * a.n;
}
__device__ __forceinline__ int cr_node_at(int level, int t) { return (1 << (level - 1)) * (2 * t + 1); }

// Node i, eliminated at `level`, has received contributions from its left side at every level below (the node i - 2^(k-1)
// always exists) and from its right side iff node i + 1 exists; the first contribution of either side comes at level 1.
__device__ __forceinline__ bool cr_has_acc_left(int level) { return level >= 2; }
__device__ __forceinline__ bool cr_has_acc_right(const CrArgs& a, int node, int level) { return level >= 2 && node < a.N; }

// Border: the system is [[B, F], [F', C]] with B block tridiagonal in nodes.  Eliminating node i also produces
//   V_b = L^-1 F_i;  F_l -= V_l' V_b,  F_r -= V_r' V_b  (accumulated like AccR / AccL: one writer per level),
//   C -= V_b' V_b,  g_C -= V_b' y_i  (one partial per node, summed in node order by k_crb_solve),
// after the last level C x_C = g_C is one small dense solve, and the back substitution subtracts V_b x_C.
//   F, FaccL, FaccR, Vb: N x n x nbp;  Cpart: N x nbp x nbp;  gpart: N x nbp;  C0: nbp x nbp;  gB, xB: nbp.
enum { CRB_F = 0, CRB_FACCL, CRB_FACCR, CRB_VB, CRB_NARR };
__host__ __device__ inline size_t crb_doubles(int N, int n, int nbp) {
  return (size_t)CRB_NARR * N * n * nbp + (size_t)N * nbp * nbp + (size_t)N * nbp + (size_t)nbp * nbp + 2 * (size_t)nbp;
}
__device__ __forceinline__ double* crb_arr(const CrArgs& a, int k, int node) {
  return a.bbase + ((size_t)k * a.N + (node - 1)) * a.n * a.nbp;
}
__device__ __forceinline__ double* crb_cpart(const CrArgs& a, int node) {
  return a.bbase + (size_t)CRB_NARR * a.N * a.n * a.nbp + (size_t)(node - 1) * a.nbp * a.nbp;
}
__device__ __forceinline__ double* crb_gpart(const CrArgs& a, int node) {
  return a.bbase + (size_t)CRB_NARR * a.N * a.n * a.nbp + (size_t)a.N * a.nbp * a.nbp + (size_t)(node - 1) * a.nbp;
}
__device__ __forceinline__ double* crb_c0(const CrArgs& a) {
  return a.bbase + (size_t)CRB_NARR * a.N * a.n * a.nbp + (size_t)a.N * a.nbp * a.nbp + (size_t)a.N * a.nbp;
}
__device__ __forceinline__ double* crb_gb(const CrArgs& a) { return crb_c0(a) + (size_t)a.nbp * a.nbp; }
__device__ __forceinline__ double* crb_xb(const CrArgs& a) { return crb_gb(a) + a.nbp; }

constexpr int kCrMaxN = 144;            // 6 * kBandMaxW
inline size_t cr_factor_smem(int n) {
  return ((size_t)(n + 1) * (n + 2) / 2 + chol_scratch_doubles(n)) * sizeof(double);
}

// ---------------------------------------------------------------------------------------------------------------------
// Assembly: kCrAsmSplit CTAs per node, each zero-fills a quarter of the node's block rows in D0 / Ep and scatters the Sblk
// blocks that land there (one CTA per node spent 53 us writing 230 KB of zeros with 256 threads: latency of one SM).
constexpr int kCrAsmSplit = 4;          // Wb is a multiple of 4
__global__ void __launch_bounds__(256) k_cr_assemble(BaDev d, CrArgs a) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int node = blockIdx.x + 1, part = blockIdx.y, tid = threadIdx.x, n = a.n;
  double* D0 = cr_arr(a, CR_D0, node);
  double* Ep = cr_arr(a, CR_EP, node);
  const int bl0 = part * (a.Wb / kCrAsmSplit), bl1 = bl0 + a.Wb / kCrAsmSplit;     // block rows of this CTA
  // (AccL / AccR / bL / bR need no clearing: their first contribution, at level 1, is a plain store — cr_has_acc)
  const double2 z2 = make_double2(0.0, 0.0);
  {
    double2* d2 = (double2*)(D0 + (size_t)6 * bl0 * n);
    double2* e2 = (double2*)(Ep + (size_t)6 * bl0 * n);
    for (int e = tid; e < 6 * (bl1 - bl0) * n / 2; e += 256) { d2[e] = z2; e2[e] = z2; }
  }
  __syncthreads();
  const int b0 = (node - 1) * a.Wb, NB = a.W + 1;
  for (int e = tid; e < (bl1 - bl0) * NB * 36; e += 256) {
    const int bl = bl0 + e / (NB * 36), rem = e % (NB * 36), off = rem / 36, rc = rem - 36 * off, r = rc / 6, c = rc - 6 * r;
    const int b = b0 + bl, aa = b - off;
    if (b >= d.Kv) {                                   // padding block: identity
      if (off == 0 && r == c) D0[(size_t)(6 * bl + r) * n + 6 * bl + r] = 1.0;
      continue;
    }
    if (aa < 0) continue;
    const int id = a.band_blk[(size_t)b * NB + off];
    if (id < 0) continue;
    const double v = d.Sblk[(size_t)id * 36 + rc];     // S[6 aa + r][6 b + c]
    if (aa >= b0) {
      if (aa == b && c < r) continue;                  // lower triangle: row 6b + c >= column 6a + r
      D0[(size_t)(6 * bl + c) * n + 6 * (aa - b0) + r] = v;
    } else {
      Ep[(size_t)(6 * bl + c) * n + 6 * (aa - (b0 - a.Wb)) + r] = v;   // S_{node, node-1}
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Factor one node per CTA: gather D and b, packed Cholesky in shared memory (the rhs rides as row n and leaves as y), the
// packed factor and y to HBM.
__global__ void __launch_bounds__(kSolveThreads) k_cr_factor(BaDev d, CrArgs a, int level) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;                // (k_point_prep raises solve_failed for a singular point block)
  const int node = cr_node_at(level, blockIdx.x), n = a.n, tid = threadIdx.x;
  double* L = smem_d;
  double* P = L + (size_t)(n + 1) * (n + 2) / 2;
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  const double* __restrict__ D0 = cr_arr(a, CR_D0, node);
  const double* __restrict__ AccL = cr_arr(a, CR_ACCL, node);
  const double* __restrict__ AccR = cr_arr(a, CR_ACCR, node);
  const bool hl = cr_has_acc_left(level), hr = cr_has_acc_right(a, node, level);
#ifdef CMOS_CR_TIMING
  long long tk[6]; tk[0] = clock64();
#endif
  // gather the lower triangle, flattened over its packed index so that every thread has 4 x 3 independent loads in flight
  {
    const int ne = n * (n + 1) / 2;
    for (int e0 = tid; e0 < ne; e0 += 4 * kSolveThreads) {
      double v[4]; int idx[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * kSolveThreads;
        idx[u] = -1; v[u] = 0.0;
        if (e < ne) {
          int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
          while ((r + 1) * (r + 2) / 2 <= e) r++;
          while (r * (r + 1) / 2 > e) r--;
          const int c = e - r * (r + 1) / 2;
          const size_t g = (size_t)r * n + c;
          idx[u] = e;
          const double al = hl ? AccL[g] : 0.0, ar = hr ? AccR[g] : 0.0;
          v[u] = (D0[g] - al) - ar;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) if (idx[u] >= 0) L[idx[u]] = v[u];
    }
  }
  {
    const int b0 = (node - 1) * a.Wb;
    const double* bL = cr_vec(a, CR_BL, node);
    const double* bR = cr_vec(a, CR_BR, node);
    for (int k = tid; k < n; k += kSolveThreads) {
      const int cam = b0 + k / 6;
      const double r0 = a.rhs_nodes ? a.rhs_nodes[(size_t)(node - 1) * n + k] : (cam < d.Kv ? d.rhs[6 * b0 + k] : 0.0);
      L[n * (n + 1) / 2 + k] = (r0 - (hl ? bL[k] : 0.0)) - (hr ? bR[k] : 0.0);
    }
  }
  __syncthreads();
#ifdef CMOS_CR_TIMING
  tk[1] = clock64();
#endif
  // (the 24 x 24 diagonal blocks leave inverted: the spike and back substitutions run in 24-row block steps of tensor-core
  // tile products)
  factor_and_invert24(L, P, n, &s_fail);
  __syncthreads();
#ifdef CMOS_CR_TIMING
  tk[2] = clock64();
#endif
  if (s_fail) {                                          // not positive definite: the LM step is invalid
    if (tid == 0) st.solve_failed = 1;
    return;
  }
  double2* Lp = (double2*)cr_arr(a, CR_LP, node);
  const int ne2 = (n * (n + 1) / 2 + 1) / 2;             // packed triangle, in double2 (n (n + 1) / 2 is even for n % 4 == 0)
  for (int e = tid; e < ne2; e += kSolveThreads) Lp[e] = ((const double2*)L)[e];
  double* y = cr_vec(a, CR_Y, node);
  for (int k = tid; k < n; k += kSolveThreads) y[k] = L[n * (n + 1) / 2 + k];
#ifdef CMOS_CR_TIMING
  tk[3] = clock64();
  if (tid == 0 && blockIdx.x == 0 && st.iteration == 1)
    printf("k_cr_factor level %d n %d: gather %lld potrf %lld store %lld cycles\n", level, n, tk[1] - tk[0], tk[2] - tk[1],
           tk[3] - tk[2]);
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// fp64 tensor-core tile: one warp accumulates a 24 x 24 output tile as 3 x 3 mma.m8n8k4 tiles.
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): A[row = lane / 4][col = lane % 4], B[row = lane % 4][col = lane / 4],
// C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
struct CrTile { double c[3][3][2]; };
__device__ __forceinline__ void cr_tile_zero(CrTile& t) {
#pragma unroll
  for (int mi = 0; mi < 3; mi++)
#pragma unroll
    for (int ni = 0; ni < 3; ni++) t.c[mi][ni][0] = t.c[mi][ni][1] = 0.0;
}
// t += A B over k in [k_begin, k_end) (multiples of 4): load_a(row in tile, k), load_b(k, column in tile)
template <typename FA, typename FB>
__device__ __forceinline__ void cr_tile_acc(CrTile& t, const int k_begin, const int k_end, FA load_a, FB load_b) {
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  for (int k0 = k_begin; k0 < k_end; k0 += 4) {
    double fa[3], fb[3];
#pragma unroll
    for (int mi = 0; mi < 3; mi++) fa[mi] = load_a(8 * mi + g, k0 + q);
#pragma unroll
    for (int ni = 0; ni < 3; ni++) fb[ni] = load_b(k0 + q, 8 * ni + g);
#pragma unroll
    for (int mi = 0; mi < 3; mi++)
#pragma unroll
      for (int ni = 0; ni < 3; ni++) dmma_884(t.c[mi][ni][0], t.c[mi][ni][1], fa[mi], fb[ni]);
  }
}
template <typename FA, typename FB>
__device__ __forceinline__ void cr_tile_mma(CrTile& t, const int k_begin, const int k_end, FA load_a, FB load_b) {
  cr_tile_zero(t);
  cr_tile_acc(t, k_begin, k_end, load_a, load_b);
}

// The node that connected i with its level-`level` neighbour on side `right` (0 = left), and how to read S_{i,nb} from it:
// level 1: the assembled S_{i,i-1} (left) or S_{i+1,i} transposed (right); above: -C_lr of the node eliminated between them.
struct CrCoupling { const double* src; bool trans; double sign; };
__device__ __forceinline__ CrCoupling cr_coupling(const CrArgs& a, int node, int level, int right) {
  CrCoupling c;
  if (level == 1) {
    c.src = cr_arr(a, CR_EP, right ? node + 1 : node); c.trans = right != 0; c.sign = 1.0;
  } else {
    const int h = 1 << (level - 2);
    c.src = cr_arr(a, CR_CLR, right ? node + h : node - h); c.trans = right == 0; c.sign = -1.0;
  }
  return c;
}

// V_side = L^-1 S_{i,side} by blocked forward substitution in 24-row steps on the fp64 tensor cores.  Grid (slab groups, side,
// node); the CTA holds the packed factor in shared memory, each of its warps one 24-column slab of S:
//   p = 0 .. n/24 - 1:   W = S_p - sum_{q < p} L_pq X_q,   X_p = inv(L_pp) W        (24 x 24 x 24 tile products)
constexpr int kCrGemmWarps = 4;
constexpr int kCrSlab = 24;
inline int cr_spike_warps(int n) {                         // slabs per CTA: a divisor of n / 24 whose staging fits 200 KB
  const int nt = n / kCrSlab;
  for (int w = nt; w >= 1; w--)
    if (nt % w == 0 && ((size_t)n * (n + 1) / 2 + (size_t)n * (kCrSlab * w + 4)) * sizeof(double) <= 200 * 1024) return w;
  return 1;
}
inline size_t cr_spike_smem(int n) {
  return ((size_t)n * (n + 1) / 2 + (size_t)n * (kCrSlab * cr_spike_warps(n) + 4)) * sizeof(double);
}
// side 2 (launched on its own, side0 = 2): V_b = L^-1 (F_i - FaccL - FaccR), nbp columns.
__global__ void __launch_bounds__(192) k_cr_spike(BaDev d, CrArgs a, int level, int side0) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int node = cr_node_at(level, blockIdx.z), side = side0 + blockIdx.y, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbr = side == 1 ? node + (1 << (level - 1)) : node - (1 << (level - 1));
  if (side < 2 && (nbr < 1 || nbr > a.N)) return;
  const int kw = blockDim.x >> 5, ld = kCrSlab * kw + 4;   // + 4: the four k rows of a B fragment fall into different banks
  double* L = smem_d;                                    // packed factor
  double* Xs = L + (size_t)n * (n + 1) / 2;              // [n][ld]: solved block rows, warp w owns columns 24 w .. 24 w + 23
  {
    const double2* __restrict__ Lp = (const double2*)cr_arr(a, CR_LP, node);
    const int ne2 = (n * (n + 1) / 2 + 1) / 2;
    for (int e = tid; e < ne2; e += blockDim.x) ((double2*)L)[e] = Lp[e];
  }
  __syncthreads();
  const bool border = side == 2;
  const CrCoupling cp = border ? CrCoupling{crb_arr(a, CRB_F, node), false, 1.0} : cr_coupling(a, node, level, side);
  const double* __restrict__ src = cp.src;
  const int j0 = kCrSlab * (blockIdx.x * kw + warp), wc = kCrSlab * warp;
  const int g = lane >> 2, q = lane & 3;
  const int ldv = border ? a.nbp : n;                    // columns of the coupling block and of V
  if (j0 >= ldv) return;                                 // (only __syncwarp below)
  double* V = border ? crb_arr(a, CRB_VB, node) : cr_arr(a, side ? CR_VR : CR_VL, node);
  const double* __restrict__ fl = border && cr_has_acc_left(level) ? crb_arr(a, CRB_FACCL, node) : nullptr;
  const double* __restrict__ fr = border && cr_has_acc_right(a, node, level) ? crb_arr(a, CRB_FACCR, node) : nullptr;
  for (int p = 0; p < n / kCrSlab; p++) {
    const int r0 = kCrSlab * p;
    CrTile t;
#pragma unroll
    for (int mi = 0; mi < 3; mi++)
#pragma unroll
      for (int ni = 0; ni < 3; ni++) {
        const int row = r0 + 8 * mi + g, col = j0 + 8 * ni + 2 * q;
        if (border) {
          const size_t o = (size_t)row * ldv + col;
          double2 v = __ldg((const double2*)(src + o));
          if (fl) { const double2 u = *(const double2*)(fl + o); v.x -= u.x; v.y -= u.y; }
          if (fr) { const double2 u = *(const double2*)(fr + o); v.x -= u.x; v.y -= u.y; }
          t.c[mi][ni][0] = v.x; t.c[mi][ni][1] = v.y;
        } else if (cp.trans) {
          t.c[mi][ni][0] = cp.sign * __ldg(src + (size_t)col * n + row);
          t.c[mi][ni][1] = cp.sign * __ldg(src + (size_t)(col + 1) * n + row);
        } else {
          const double2 v = __ldg((const double2*)(src + (size_t)row * n + col));
          t.c[mi][ni][0] = cp.sign * v.x; t.c[mi][ni][1] = cp.sign * v.y;
        }
      }
    cr_tile_acc(t, 0, r0,                                 // W -= L_p,0..p-1 X_0..p-1
                [&](int r, int k) { const int R = r0 + r; return -L[R * (R + 1) / 2 + k]; },
                [&](int k, int c) { return Xs[k * ld + wc + c]; });
#pragma unroll
    for (int mi = 0; mi < 3; mi++)
#pragma unroll
      for (int ni = 0; ni < 3; ni++)
        *(double2*)(Xs + (r0 + 8 * mi + g) * ld + wc + 8 * ni + 2 * q) = make_double2(t.c[mi][ni][0], t.c[mi][ni][1]);
    __syncwarp();
    CrTile x;
    cr_tile_mma(x, 0, kCrSlab,                            // X_p = inv(L_pp) W, the inverse is lower triangular
                [&](int r, int k) { const int R = r0 + r; return k <= r ? L[R * (R + 1) / 2 + r0 + k] : 0.0; },
                [&](int k, int c) { return Xs[(r0 + k) * ld + wc + c]; });
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 3; mi++)
#pragma unroll
      for (int ni = 0; ni < 3; ni++) {
        const double2 v = make_double2(x.c[mi][ni][0], x.c[mi][ni][1]);
        *(double2*)(Xs + (r0 + 8 * mi + g) * ld + wc + 8 * ni + 2 * q) = v;
        *(double2*)(V + (size_t)(r0 + 8 * mi + g) * ldv + j0 + 8 * ni + 2 * q) = v;
      }
    __syncwarp();
  }
}

// Schur products of an eliminated node: grid (tile groups, product, node).
//   product 0: AccR[l] += V_l' V_l (lower tiles)   1: AccL[r] += V_r' V_r (lower tiles)   2: Clr[i] = V_l' V_r
//   product 3: bR[l] += V_l' y_i, bL[r] += V_r' y_i
__device__ __forceinline__ void cr_schur_body(const BaDev& d, const CrArgs& a, const int level, const int prod) {
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int node = cr_node_at(level, blockIdx.z), n = a.n, nt = n / 24, h = 1 << (level - 1);
  const int l = node - h, r = node + h;
  const bool has_l = l >= 1, has_r = r <= a.N;
  const double* __restrict__ Vl = cr_arr(a, CR_VL, node);
  const double* __restrict__ Vr = cr_arr(a, CR_VR, node);
  if (prod == 3) {
    if (blockIdx.x != 0) return;
    const double* __restrict__ y = cr_vec(a, CR_Y, node);
    for (int p = threadIdx.x; p < 2 * n; p += 32 * kCrGemmWarps) {
      const int sd = p >= n, c = p - sd * n;
      if (sd ? !has_r : !has_l) continue;
      const double* V = sd ? Vr : Vl;
      double v = 0.0;
      for (int k = 0; k < n; k++) v += V[(size_t)k * n + c] * y[k];
      double* dst = sd ? cr_vec(a, CR_BL, r) : cr_vec(a, CR_BR, l);
      dst[c] = level == 1 ? v : dst[c] + v;               // the first contribution of a side is a store (no clearing pass)
    }
    return;
  }
  if ((prod == 0 && !has_l) || (prod == 1 && !has_r) || (prod == 2 && !(has_l && has_r))) return;
  const int tile = blockIdx.x * kCrGemmWarps + (threadIdx.x >> 5);
  if (tile >= nt * nt) return;
  const int p0 = 24 * (tile / nt), q0 = 24 * (tile % nt);
  if (prod < 2 && q0 > p0) return;                       // symmetric products: lower tiles only
  const double* __restrict__ Va = prod == 1 ? Vr : Vl;
  const double* __restrict__ Vb = prod == 0 ? Vl : Vr;
  CrTile t;
  cr_tile_mma(t, 0, n,
              [&](int rr, int k) { return __ldg(Va + (size_t)k * n + p0 + rr); },
              [&](int k, int c) { return __ldg(Vb + (size_t)k * n + q0 + c); });
  double* dst = prod == 0 ? cr_arr(a, CR_ACCR, l) : prod == 1 ? cr_arr(a, CR_ACCL, r) : cr_arr(a, CR_CLR, node);
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 3; mi++)
#pragma unroll
    for (int ni = 0; ni < 3; ni++) {
      double2* o = (double2*)(dst + (size_t)(p0 + 8 * mi + g) * n + q0 + 8 * ni + 2 * q);
      if (prod == 2 || level == 1) *o = make_double2(t.c[mi][ni][0], t.c[mi][ni][1]);
      else { const double2 old = *o; *o = make_double2(old.x + t.c[mi][ni][0], old.y + t.c[mi][ni][1]); }
    }
}

// Border products of an eliminated node: grid (tile groups, product, node).
//   product 0: FaccR[l] += V_l' V_b   1: FaccL[r] += V_r' V_b   2: Cpart[i] = V_b' V_b (lower tiles)   3: gpart[i] = V_b' y_i
__device__ __forceinline__ void crb_schur_body(const BaDev& d, const CrArgs& a, const int level, const int prod) {
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int node = cr_node_at(level, blockIdx.z), n = a.n, nbp = a.nbp, h = 1 << (level - 1);
  const int l = node - h, r = node + h;
  const bool has_l = l >= 1, has_r = r <= a.N;
  const double* __restrict__ Vb = crb_arr(a, CRB_VB, node);
  if (prod == 3) {
    if (blockIdx.x != 0) return;
    const double* __restrict__ y = cr_vec(a, CR_Y, node);
    double* gp = crb_gpart(a, node);
    for (int c = threadIdx.x; c < nbp; c += 32 * kCrGemmWarps) {
      double v = 0.0;
      for (int k = 0; k < n; k++) v += Vb[(size_t)k * nbp + c] * y[k];
      gp[c] = v;
    }
    return;
  }
  if ((prod == 0 && !has_l) || (prod == 1 && !has_r)) return;
  const int ntr = (prod == 2 ? nbp : n) / 24, ntc = nbp / 24;
  const int tile = blockIdx.x * kCrGemmWarps + (threadIdx.x >> 5);
  if (tile >= ntr * ntc) return;
  const int p0 = 24 * (tile / ntc), q0 = 24 * (tile % ntc);
  if (prod == 2 && q0 > p0) return;
  const double* __restrict__ Va = prod == 0 ? cr_arr(a, CR_VL, node) : prod == 1 ? cr_arr(a, CR_VR, node) : Vb;
  const int lda = prod == 2 ? nbp : n;
  CrTile t;
  cr_tile_mma(t, 0, n,
              [&](int rr, int k) { return __ldg(Va + (size_t)k * lda + p0 + rr); },
              [&](int k, int c) { return __ldg(Vb + (size_t)k * nbp + q0 + c); });
  double* dst = prod == 0 ? crb_arr(a, CRB_FACCR, l) : prod == 1 ? crb_arr(a, CRB_FACCL, r) : crb_cpart(a, node);
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 3; mi++)
#pragma unroll
    for (int ni = 0; ni < 3; ni++) {
      double2* o = (double2*)(dst + (size_t)(p0 + 8 * mi + g) * nbp + q0 + 8 * ni + 2 * q);
      if (prod == 2 || level == 1) *o = make_double2(t.c[mi][ni][0], t.c[mi][ni][1]);
      else { const double2 old = *o; *o = make_double2(old.x + t.c[mi][ni][0], old.y + t.c[mi][ni][1]); }
    }
}

__global__ void __launch_bounds__(32 * kCrGemmWarps) k_cr_schur(BaDev d, CrArgs a, int level) { pdl_begin(); cr_schur_body(d, a, level, blockIdx.y); }
__global__ void __launch_bounds__(32 * kCrGemmWarps) k_crb_schur(BaDev d, CrArgs a, int level) { pdl_begin(); crb_schur_body(d, a, level, blockIdx.y); }
// both in one launch (bordered systems, every level but the last): grid (tile groups, 8 products, node)
__global__ void __launch_bounds__(32 * kCrGemmWarps) k_cr_schur_all(BaDev d, CrArgs a, int level) { pdl_begin();
  if (blockIdx.y < 4) cr_schur_body(d, a, level, blockIdx.y);
  else crb_schur_body(d, a, level, blockIdx.y - 4);
}

// The border system after every node is eliminated: (C0 - sum_i Cpart[i]) x_C = gB - sum_i gpart[i], sums in node order;
// one CTA, the same small dense solver as LocalBundleAdjustment's reduced camera system.
inline size_t crb_solve_smem(int nbp) { return ((size_t)(nbp + 1) * (nbp + 2) / 2 + chol_scratch_doubles(nbp)) * sizeof(double); }
__global__ void __launch_bounds__(kSolveThreads) k_crb_solve(BaDev d, CrArgs a) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int nb = a.nbp, tid = threadIdx.x;
  double* L = smem_d;
  double* P = L + (size_t)(nb + 1) * (nb + 2) / 2;
  __shared__ int s_fail;
  __shared__ double s_x24[24];
  if (tid == 0) s_fail = 0;
  const double* __restrict__ C0 = crb_c0(a);
  const double* __restrict__ gB = crb_gb(a);
  const double* __restrict__ Cp = crb_cpart(a, 1);
  const double* __restrict__ gp = crb_gpart(a, 1);
  const int ne = nb * (nb + 1) / 2;
  for (int e = tid; e < ne + nb; e += kSolveThreads) {
    double v, s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;       // four interleaved partial sums, fixed order
    size_t off, stride;
    if (e < ne) {
      int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while ((r + 1) * (r + 2) / 2 <= e) r++;
      while (r * (r + 1) / 2 > e) r--;
      const int c = e - r * (r + 1) / 2;
      off = (size_t)r * nb + c; stride = (size_t)nb * nb;
      v = C0[off];
    } else {
      off = e - ne; stride = nb;
      v = gB[off];
    }
    const double* __restrict__ src = e < ne ? Cp : gp;
    int i = 0;
    for (; i + 4 <= a.N; i += 4) {
      s0 += src[off + (size_t)i * stride]; s1 += src[off + (size_t)(i + 1) * stride];
      s2 += src[off + (size_t)(i + 2) * stride]; s3 += src[off + (size_t)(i + 3) * stride];
    }
    for (; i < a.N; i++) s0 += src[off + (size_t)i * stride];
    L[e] = v - ((s0 + s1) + (s2 + s3));
  }
  __syncthreads();
  factor_and_solve(L, P, nb, &s_fail, s_x24);
  __syncthreads();
  if (s_fail) { if (tid == 0) st.solve_failed = 1; return; }
  const double* y = L + ne;
  double* xb = crb_xb(a);
  for (int k = tid; k < nb; k += kSolveThreads) xb[k] = y[k];
}

// Back substitution of one level, one CTA per node: z = y_i - V_l x_l - V_r x_r (rows are independent: a warp takes four
// rows per step so that 16-40 loads per lane are in flight), then L' x = z by blocked substitution against the packed
// factor in shared memory, from the last 24-row block up (right-looking: x_p = inv(L_pp)' z_p, then z_q -= L_pq' x_p, q < p).
constexpr int kCrBackWarps = 16;
inline size_t cr_back_smem(int n) { return ((size_t)n * (n + 1) / 2 + 2) * sizeof(double); }
__global__ void __launch_bounds__(32 * kCrBackWarps) k_cr_back(BaDev d, CrArgs a, int level) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double s_z[kCrMaxN], s_xl[kCrMaxN], s_xr[kCrMaxN], s_x[kCrMaxN];
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int node = cr_node_at(level, blockIdx.x), n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = 1 << (level - 1), l = node - h, r = node + h;
  const bool has_l = l >= 1, has_r = r <= a.N;
  double* L = smem_d;
  {
    const double2* __restrict__ Lp = (const double2*)cr_arr(a, CR_LP, node);
    const int ne2 = (n * (n + 1) / 2 + 1) / 2;
    for (int e = tid; e < ne2; e += 32 * kCrBackWarps) ((double2*)L)[e] = Lp[e];
  }
  for (int k = tid; k < n; k += 32 * kCrBackWarps) {
    s_xl[k] = has_l ? cr_vec(a, CR_X, l)[k] : 0.0;
    s_xr[k] = has_r ? cr_vec(a, CR_X, r)[k] : 0.0;
  }
  __syncthreads();
  const double* __restrict__ Vl = cr_arr(a, CR_VL, node);
  const double* __restrict__ Vr = cr_arr(a, CR_VR, node);
  const double* __restrict__ y = cr_vec(a, CR_Y, node);
  const double* __restrict__ Vb = a.nbp ? crb_arr(a, CRB_VB, node) : nullptr;
  const double* __restrict__ xb = a.nbp ? crb_xb(a) : nullptr;
  constexpr int kCols = (kCrMaxN + 31) / 32;               // column groups per lane
  for (int k0 = 4 * warp; k0 < n; k0 += 4 * kCrBackWarps) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < kCols; u++) {
      const int j = lane + 32 * u;
      if (j < n) {
        const double xl = s_xl[j], xr = s_xr[j];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = min(k0 + q, n - 1);
          const double vl = has_l ? Vl[(size_t)k * n + j] : 0.0;
          const double vr = has_r ? Vr[(size_t)k * n + j] : 0.0;
          acc[q] += vl * xl + vr * xr;
        }
      }
    }
    if (Vb) {                                                // bordered system: ... - V_b x_C
#pragma unroll
      for (int u = 0; u < kCols; u++) {
        const int j = lane + 32 * u;
        if (j < a.nbp) {
          const double xj = xb[j];
#pragma unroll
          for (int q = 0; q < 4; q++) acc[q] += Vb[(size_t)min(k0 + q, n - 1) * a.nbp + j] * xj;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    if (lane < 4 && k0 + lane < n) {
      const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      s_z[k0 + lane] = y[k0 + lane] - v;
    }
  }
  __syncthreads();
  back_substitute24(L, s_z, s_x, n);       // s_z becomes x
  double* x = cr_vec(a, CR_X, node);
  const int b0 = (node - 1) * a.Wb;
  for (int j = tid; j < n; j += 32 * kCrBackWarps) {
    const double v = s_z[j];
    x[j] = v;
    if (a.x_nodes) a.x_nodes[(size_t)(node - 1) * n + j] = v;
    else if (b0 + j / 6 < d.Kv) d.yc[6 * b0 + j] = v;
  }
}

}  // namespace cmos
