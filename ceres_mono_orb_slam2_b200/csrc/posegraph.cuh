// CeresOptimizer::OptimizeEssentialGraph on the device (src/CeresOptimizer.cc:736-957; EssentialGraphErrorTerm
// include/CeresOptimizer.h:270-328; Sim3Parameterization src/CeresOptimizer.cc:24-47).  Included by ba.cu after the blocked
// Cholesky kernels, which this solve reuses on its own dense system through a BaDev view (S, rhs, yc, nc, st).
//
// Unknowns: the Sim3 logs [upsilon, omega, sigma] of every keyframe but the loop keyframe (7 per keyframe).  One residual
// block per essential-graph edge: r = log(Sji * exp(x_i) * exp(x_j)^-1), Jacobian wrt x_i = (I + ad(r)/2 + ad(r)^2/12) *
// Adj(exp(x_j)), wrt x_j its negative — so per edge ONE 7x7 product A = J'J and one 7-vector v = J'r feed both diagonal
// blocks (+A), the off-diagonal block (-A) and the two gradients (+v, -v).
//
// Kernels (E edges, Kv variable keyframes, n = 7 Kv):
//   k_eg_logs / k_eg_meas    x0 = log(Scw); per edge the measurement Sji from the initial values (:797-809, :822-895)
//   k_eg_linearize           thread per edge: r, J, A, v, cost term
//   k_eg_assemble            warp per variable keyframe: fixed-order sums over its incident edges -> diagonal block, gradient,
//                            || x - Plus(x, -g) ||_inf and |x|^2 terms
//   k_eg_post_lin            one CTA: fixed-order reductions, Jacobi scaling at iteration 0, LM bookkeeping
//   k_eg_build               warp per variable keyframe: its block row of the scaled, damped normal equations into the dense
//                            lower triangle (only this warp writes the row, duplicates of a keyframe pair accumulate in order)
//   solve                    nested dissection of the band + border structure (band_cr.cuh: keyframes with long-range edges form
//                            the border, the rest is block tridiagonal in nodes of Wb keyframes), or, when the graph has no
//                            such structure, blocked Cholesky k_potrf_diag / k_trsm_panel / k_syrk_tile inside the row envelope
//   k_eg_step                thread per keyframe: delta = -y * scale, candidate = Plus(x, delta)
//   k_eg_eval                thread per edge: model cost change -(J d)'(r + J d / 2) and the candidate's cost term
//   k_eg_decide              one CTA: reductions, accept / reject, radius, termination tests
//   k_eg_finish / k_eg_points  Tiw = [R | t / s] and X' = exp(x_r)^-1 * (exp(x0_r) * X) (:903-956)
// No floating-point atomics; every sum has a fixed order.
#pragma once

namespace cmos {

struct EgDev {
  int n_kf, Kv, E, n;
  double* x[2];                       // [n_kf][7]  current / candidate (st.cur)
  double* x0;                         // [n_kf][7]  Scw_original_datas
  const double *Scw, *Snc;            // [n_kf][13] scale, R (row-major), t
  const uint8_t* kf_flags;            // bit 0 constant, bit 1 has a non-corrected Sim3
  const int *var, *var_kf;            // [n_kf] keyframe -> variable index or -1; [Kv]
  const int *edge_j, *edge_i;         // [E]
  const uint8_t* edge_kind;           // [E]
  Sim3D* meas;                        // [E]
  double *r, *J, *A, *v;              // [E][7], [E][49], [E][49], [E][7]
  const int *inc_start, *inc_edge, *inc_other, *inc_sign;   // CSR of the incident edges of every variable keyframe
  double *Hd, *g, *scale, *delta;     // [Kv][49], [n], [n], [n]
  double *S, *rhs, *yc;               // [n][n] lower, [n], [n]
  double *p_cost, *p_mcc, *p_cand;    // [E]
  double *p_gmax, *p_xn2, *p_sn2;     // [Kv]
  LmState* st;
  double* trace;
  // Nested-dissection solve of the normal equations (band_cr.cuh with a border): bcr != 0 -> k_eg_build writes the node
  // blocks instead of the dense lower triangle.  pos[a] >= 0: position of variable keyframe a among the interior
  // (banded) keyframes; pos[a] < 0: border keyframe -1 - pos[a] (a keyframe with an edge to a far-away one).
  int bcr, n_interior, n_border;
  const int* pos;
  CrArgs ca;
};

// Where entry (scalar r of variable keyframe a, scalar c of variable keyframe b <= a) of the lower triangle lives.
__device__ __forceinline__ double* eg_entry(const EgDev& d, int a, int r, int b, int c) {
  if (!d.bcr) return d.S + (size_t)(7 * a + r) * d.n + 7 * b + c;
  const CrArgs& ca = d.ca;
  const int pa = d.pos[a], pb = d.pos[b];
  if (pa >= 0 && pb >= 0) {                                  // interior x interior: same node, or node and its left neighbour
    const int na = pa / ca.Wb, nb = pb / ca.Wb;
    const int la = 7 * (pa - na * ca.Wb) + r, lb = 7 * (pb - nb * ca.Wb) + c;
    return cr_arr(ca, na == nb ? CR_D0 : CR_EP, na + 1) + (size_t)la * ca.n + lb;
  }
  if (pa < 0 && pb < 0) return crb_c0(ca) + (size_t)(7 * (-1 - pa) + r) * ca.nbp + 7 * (-1 - pb) + c;
  if (pa < 0) {                                              // border row, interior column: F_node(b)[b's scalar][a's scalar]
    const int nb = pb / ca.Wb;
    return crb_arr(ca, CRB_F, nb + 1) + (size_t)(7 * (pb - nb * ca.Wb) + c) * ca.nbp + 7 * (-1 - pa) + r;
  }
  const int na = pa / ca.Wb;                                 // interior row, border column (a border keyframe with a smaller index)
  return crb_arr(ca, CRB_F, na + 1) + (size_t)(7 * (pa - na * ca.Wb) + r) * ca.nbp + 7 * (-1 - pb) + c;
}
__device__ __forceinline__ double* eg_rhs_entry(const EgDev& d, int a, int r) {
  if (!d.bcr) return d.rhs + 7 * a + r;
  const int pa = d.pos[a];
  if (pa < 0) return crb_gb(d.ca) + 7 * (-1 - pa) + r;
  const int na = pa / d.ca.Wb;
  return const_cast<double*>(d.ca.rhs_nodes) + (size_t)na * d.ca.n + 7 * (pa - na * d.ca.Wb) + r;
}

// Zero the node blocks the assembly accumulates into and put ones on the diagonal of the padding rows (a node holds Wb
// keyframes = 7 Wb unknowns, padded to a multiple of 24; the border likewise).  One CTA per node + one for the border.
__global__ void __launch_bounds__(256) k_eg_cr_clear(EgDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done) return;
  const CrArgs& ca = d.ca;
  const int tid = threadIdx.x, n = ca.n, nbp = ca.nbp;
  if ((int)blockIdx.x < ca.N) {
    const int node = blockIdx.x + 1;
    double* D0 = cr_arr(ca, CR_D0, node);
    double* Ep = cr_arr(ca, CR_EP, node);
    for (int e = tid; e < n * n; e += 256) { D0[e] = 0.0; Ep[e] = 0.0; }
    if (nbp) { double* F = crb_arr(ca, CRB_F, node); for (int e = tid; e < n * nbp; e += 256) F[e] = 0.0; }
    double* rn = const_cast<double*>(ca.rhs_nodes) + (size_t)(node - 1) * n;
    for (int e = tid; e < n; e += 256) rn[e] = 0.0;
    __syncthreads();
    const int real = 7 * max(0, min(ca.Wb, d.n_interior - (node - 1) * ca.Wb));
    for (int e = real + tid; e < n; e += 256) D0[(size_t)e * n + e] = 1.0;
  } else if (nbp) {
    double* C0 = crb_c0(ca);
    double* gB = crb_gb(ca);
    for (int e = tid; e < nbp * nbp; e += 256) C0[e] = 0.0;
    for (int e = tid; e < nbp; e += 256) gB[e] = 0.0;
    __syncthreads();
    for (int e = 7 * d.n_border + tid; e < nbp; e += 256) C0[(size_t)e * nbp + e] = 1.0;
  }
}

// solution by node / border -> yc in variable-keyframe order
__global__ void __launch_bounds__(256) k_eg_cr_scatter(EgDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= d.n) return;
  const int a = q / 7, r = q - 7 * a, pa = d.pos[a];
  double v;
  if (pa < 0) v = crb_xb(d.ca)[7 * (-1 - pa) + r];
  else { const int na = pa / d.ca.Wb; v = d.ca.x_nodes[(size_t)na * d.ca.n + 7 * (pa - na * d.ca.Wb) + r]; }
  d.yc[q] = v;
}

__device__ __forceinline__ void sim3_from_srt(const double* v13, Sim3D& S) {
  S.s = v13[0];
#pragma unroll
  for (int i = 0; i < 9; i++) S.R[i] = v13[1 + i];
#pragma unroll
  for (int i = 0; i < 3; i++) S.t[i] = v13[10 + i];
}

__global__ void __launch_bounds__(128) k_eg_logs(EgDev d, int max_iterations) { pdl_begin();
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k == 0) lm_init(*d.st, max_iterations, 0);
  if (k >= d.n_kf) return;
  Sim3D S;
  sim3_from_srt(d.Scw + 13 * (size_t)k, S);
  double v[7];
  sim3_log(S, v);
  for (int a = 0; a < 7; a++) { d.x0[7 * (size_t)k + a] = v[a]; d.x[0][7 * (size_t)k + a] = v[a]; d.x[1][7 * (size_t)k + a] = v[a]; }
}

__global__ void __launch_bounds__(128) k_eg_meas(EgDev d) { pdl_begin();
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= d.E) return;
  const int j = d.edge_j[e], i = d.edge_i[e];
  Sim3D Sjw, Siw, Swi;
  if (d.edge_kind[e] != 0 && (d.kf_flags[i] & 2)) sim3_from_srt(d.Snc + 13 * (size_t)i, Siw); else sim3_exp(d.x0 + 7 * (size_t)i, Siw);
  if (d.edge_kind[e] != 0 && (d.kf_flags[j] & 2)) sim3_from_srt(d.Snc + 13 * (size_t)j, Sjw); else sim3_exp(d.x0 + 7 * (size_t)j, Sjw);
  sim3_inverse(Siw, Swi);
  sim3_mul(Sjw, Swi, d.meas[e]);
}

// r = log(Sji * exp(xi) * exp(xj)^-1); Sj out for the adjoint
__device__ inline void eg_residual(const Sim3D& M, const double* xj, const double* xi, double* r, Sim3D* Sj_out) {
  Sim3D Si, Sj, Sji, T, Er;
  sim3_exp(xi, Si);
  sim3_exp(xj, Sj);
  sim3_mul(M, Si, T);
  sim3_inverse(Sj, Sji);
  sim3_mul(T, Sji, Er);
  sim3_log(Er, r);
  if (Sj_out) *Sj_out = Sj;
}

__global__ void __launch_bounds__(64) k_eg_linearize(EgDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  const int e = blockIdx.x * 64 + threadIdx.x;
  if (e >= d.E) return;
  const double* x = d.x[st.cur];
  double r[7];
  Sim3D Sj;
  eg_residual(d.meas[e], x + 7 * (size_t)d.edge_j[e], x + 7 * (size_t)d.edge_i[e], r, &Sj);
  // ad(r) = [[hat(w) + sigma I, hat(u), -u], [0, hat(w), 0], [0, 0, 0]]   (CeresOptimizer.h:292-296)
  double A[49];
  for (int i = 0; i < 49; i++) A[i] = 0.0;
  double Ow[9], Ou[9];
  hat3(r + 3, Ow); hat3(r, Ou);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[7 * i + j] = Ow[3 * i + j] + (i == j ? r[6] : 0.0);
      A[7 * i + 3 + j] = Ou[3 * i + j];
      A[7 * (3 + i) + 3 + j] = Ow[3 * i + j];
    }
  for (int i = 0; i < 3; i++) A[7 * i + 6] = -r[i];
  double* Jr = d.A + 49 * (size_t)e;          // scratch: the series, overwritten by J'J below
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double a2 = 0.0;
      for (int k = 0; k < 7; k++) a2 += A[7 * i + k] * A[7 * k + j];
      Jr[7 * i + j] = (i == j ? 1.0 : 0.0) + 0.5 * A[7 * i + j] + 1.0 / 12. * a2;
    }
  // Adj(Sj) = [[s R, hat(t) R, -t], [0, R, 0], [0, 0, 1]]
  double TR[9], Ot[9];
  hat3(Sj.t, Ot); mul33(Ot, Sj.R, TR);
  for (int i = 0; i < 49; i++) A[i] = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[7 * i + j] = Sj.s * Sj.R[3 * i + j]; A[7 * i + 3 + j] = TR[3 * i + j]; A[7 * (3 + i) + 3 + j] = Sj.R[3 * i + j]; }
  for (int i = 0; i < 3; i++) A[7 * i + 6] = -Sj.t[i];
  A[48] = 1.0;
  double* J = d.J + 49 * (size_t)e;
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double a = 0.0;
      for (int k = 0; k < 7; k++) a += Jr[7 * i + k] * A[7 * k + j];
      J[7 * i + j] = a;
    }
  double cost = 0.0;
  for (int a = 0; a < 7; a++) { d.r[7 * (size_t)e + a] = r[a]; cost += r[a] * r[a]; }
  d.p_cost[e] = 0.5 * cost;
  double* AtA = d.A + 49 * (size_t)e;
  for (int a = 0; a < 7; a++) {
    double gsum = 0.0;
    for (int k = 0; k < 7; k++) gsum += J[7 * k + a] * r[k];
    d.v[7 * (size_t)e + a] = gsum;
    for (int b = 0; b < 7; b++) {
      double h = 0.0;
      for (int k = 0; k < 7; k++) h += J[7 * k + a] * J[7 * k + b];
      AtA[7 * a + b] = h;
    }
  }
}

__global__ void __launch_bounds__(128) k_eg_assemble(EgDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  const int a = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (a >= d.Kv) return;
  double h0 = 0.0, h1 = 0.0, gl = 0.0;
  for (int p = d.inc_start[a]; p < d.inc_start[a + 1]; p++) {
    const int e = d.inc_edge[p];
    const double* A = d.A + 49 * (size_t)e;
    h0 += A[lane];
    if (lane + 32 < 49) h1 += A[lane + 32];
    if (lane < 7) gl += (double)d.inc_sign[p] * d.v[7 * (size_t)e + lane];
  }
  d.Hd[49 * (size_t)a + lane] = h0;
  if (lane + 32 < 49) d.Hd[49 * (size_t)a + lane + 32] = h1;
  if (lane < 7) d.g[7 * (size_t)a + lane] = gl;
  __syncwarp();
  if (lane == 0) {
    const double* x = d.x[st.cur] + 7 * (size_t)d.var_kf[a];
    double ng[7], xp[7], m = 0.0, xn = 0.0;
    for (int k = 0; k < 7; k++) ng[k] = -d.g[7 * (size_t)a + k];
    sim3_plus(x, ng, xp);
    for (int k = 0; k < 7; k++) { m = fmax(m, fabs(x[k] - xp[k])); xn += x[k] * x[k]; }
    d.p_gmax[a] = m; d.p_xn2[a] = xn;
  }
}

__global__ void __launch_bounds__(256) k_eg_post_lin(EgDev d) { pdl_begin();
  __shared__ double scratch[33];
  LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  const int tid = threadIdx.x;
  double c = 0.0, m = 0.0, xn = 0.0;
  for (int e = tid; e < d.E; e += 256) c += d.p_cost[e];
  for (int a = tid; a < d.Kv; a += 256) { m = fmax(m, d.p_gmax[a]); xn += d.p_xn2[a]; }
  c = block_sum(c, scratch);
  m = block_max(m, scratch);
  xn = block_sum(xn, scratch);
  if (st.first)
    for (int q = tid; q < d.n; q += 256) d.scale[q] = 1.0 / (1.0 + sqrt(d.Hd[49 * (size_t)(q / 7) + 8 * (q % 7)]));
  __syncthreads();
  if (tid == 0) lm_after_linearize(st, c, m, sqrt(xn), d.trace);
}

__global__ void __launch_bounds__(128) k_eg_build(EgDev d) { pdl_begin();
  LmState& st = *d.st;
  if (st.done) return;
  const int a = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) st.solve_failed = 0;
  if (a >= d.Kv) return;
  const double* sa = d.scale + 7 * (size_t)a;
  const double* Hd = d.Hd + 49 * (size_t)a;
  for (int q = lane; q < 49; q += 32) {
    const int r = q / 7, c = q - 7 * r;
    if (c > r) continue;
    double v = sa[r] * Hd[q] * sa[c];
    if (r == c) v += fmin(fmax(sa[r] * sa[r] * Hd[q], kMinLmDiag), kMaxLmDiag) / st.radius;
    *eg_entry(d, a, r, a, c) = v;
  }
  if (lane < 7) *eg_rhs_entry(d, a, lane) = sa[lane] * d.g[7 * (size_t)a + lane];
  for (int p = d.inc_start[a]; p < d.inc_start[a + 1]; p++) {
    const int b = d.inc_other[p];
    if (b < 0 || b >= a) continue;
    const double* A = d.A + 49 * (size_t)d.inc_edge[p];
    const double* sb = d.scale + 7 * (size_t)b;
    for (int q = lane; q < 49; q += 32) {
      const int r = q / 7, c = q - 7 * r;
      *eg_entry(d, a, r, b, c) -= sa[r] * A[q] * sb[c];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) k_eg_step(EgDev d) { pdl_begin();
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= d.n_kf) return;
  const int a = d.var[k];
  if (a < 0) return;                               // constant keyframes are identical in both buffers
  const double* x = d.x[st.cur] + 7 * (size_t)k;
  double* xc = d.x[st.cur ^ 1] + 7 * (size_t)k;
  double delta[7], cand[7];
  bool fin = true;
  for (int q = 0; q < 7; q++) {
    delta[q] = -d.yc[7 * a + q] * d.scale[7 * a + q];
    fin = fin && isfinite(delta[q]);
    d.delta[7 * a + q] = delta[q];
  }
  if (!fin) { st.solve_failed = 1; return; }
  sim3_plus(x, delta, cand);
  double sn2 = 0.0;
  for (int q = 0; q < 7; q++) { xc[q] = cand[q]; sn2 += (x[q] - cand[q]) * (x[q] - cand[q]); }
  d.p_sn2[a] = sn2;
}

__global__ void __launch_bounds__(64) k_eg_eval(EgDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int e = blockIdx.x * 64 + threadIdx.x;
  if (e >= d.E) return;
  const int kj = d.edge_j[e], ki = d.edge_i[e];
  const int vi = d.var[ki], vj = d.var[kj];
  double dd[7];
  for (int a = 0; a < 7; a++) dd[a] = (vi >= 0 ? d.delta[7 * vi + a] : 0.0) - (vj >= 0 ? d.delta[7 * vj + a] : 0.0);
  const double* J = d.J + 49 * (size_t)e;
  double mcc = 0.0;
  for (int row = 0; row < 7; row++) {
    double mm = 0.0;
    for (int a = 0; a < 7; a++) mm += J[7 * row + a] * dd[a];
    mcc -= mm * (d.r[7 * (size_t)e + row] + mm / 2.0);
  }
  d.p_mcc[e] = mcc;
  const double* xc = d.x[st.cur ^ 1];
  double r[7], c = 0.0;
  eg_residual(d.meas[e], xc + 7 * (size_t)kj, xc + 7 * (size_t)ki, r, nullptr);
  for (int a = 0; a < 7; a++) c += r[a] * r[a];
  d.p_cand[e] = 0.5 * c;
}

__global__ void __launch_bounds__(256) k_eg_decide(EgDev d, int* done_host) { pdl_begin();
  __shared__ double scratch[33];
  LmState& st = *d.st;
  if (st.done) { if (threadIdx.x == 0) *done_host = 1; return; }
  const int tid = threadIdx.x;
  const bool ok = !st.solve_failed;
  double mcc = 0.0, cc = 0.0, sn = 0.0;
  if (ok) {
    for (int e = tid; e < d.E; e += 256) { mcc += d.p_mcc[e]; cc += d.p_cand[e]; }
    for (int a = tid; a < d.Kv; a += 256) sn += d.p_sn2[a];
  }
  mcc = block_sum(mcc, scratch);
  cc = block_sum(cc, scratch);
  sn = block_sum(sn, scratch);
  if (tid == 0) {
    lm_decide(st, ok, mcc, cc, sqrt(sn), d.trace);
    *done_host = st.done;
  }
}

// rejected step: the candidate buffer must hold x again for the constant-keyframe invariant — not needed: only variable
// keyframes are rewritten by k_eg_step, and every variable keyframe is rewritten before the candidate is read.

__global__ void k_eg_summary(EgDev d, cmos_ba_summary* out) { pdl_begin();
  const LmState& st = *d.st;
  cmos_ba_summary s;
  s.iterations = st.iteration; s.successful_steps = st.successful; s.termination = st.termination;
  s.jacobian_evaluations = st.jac_evals; s.initial_cost = st.initial_cost; s.final_cost = st.x_cost;
  *out = s;
}

// lie_out [n_kf][7], Tiw_out [n_kf][16], Swc [n_kf] (corrected_Swcs)
__global__ void __launch_bounds__(128) k_eg_finish(EgDev d, double* lie_out, double* Tiw_out, Sim3D* Swc) { pdl_begin();
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= d.n_kf) return;
  const double* x = d.x[d.st->cur] + 7 * (size_t)k;
  Sim3D S;
  sim3_exp(x, S);
  sim3_inverse(S, Swc[k]);
  for (int a = 0; a < 7; a++) lie_out[7 * (size_t)k + a] = x[a];
  double* T = Tiw_out + 16 * (size_t)k;
  const double is = 1. / S.s;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = S.R[3 * i + j];
    T[4 * i + 3] = is * S.t[i];
  }
  T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
}

__global__ void __launch_bounds__(256) k_eg_points(EgDev d, int n_points, const double* __restrict__ Xw, const int* __restrict__ ref_kf,
                                                   const Sim3D* __restrict__ Swc, double* __restrict__ Xo) { pdl_begin();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n_points) return;
  const int rk = ref_kf[p];
  Sim3D Srw;
  sim3_exp(d.x0 + 7 * (size_t)rk, Srw);
  double a[3], c[3], b[3];
  mul3v(Srw.R, Xw + 3 * (size_t)p, a);
  for (int i = 0; i < 3; i++) c[i] = Srw.s * a[i] + Srw.t[i];
  const Sim3D& W = Swc[rk];
  mul3v(W.R, c, b);
  for (int i = 0; i < 3; i++) Xo[3 * (size_t)p + i] = W.s * b[i] + W.t[i];
}

}  // namespace cmos
