// Bundle-adjustment engine for sm_100a behind CeresOptimizer::{PoseOptimization, LocalBundleAdjustment,
// BundleAdjustment / GlobalBundleAdjustemnt, OptimizeSim3, OptimizeEssentialGraph} (reference
// src/CeresOptimizer.cc:59-225,275-342,344-599,601-957).
//
// What the reference does per ceres::Solve — autodiff residual blocks, Huber corrector, quaternion manifold,
// SPARSE_NORMAL_CHOLESKY inside a Levenberg-Marquardt trust region — is restated here as a fixed sequence of
// fp64 kernels whose control flow lives on the device (LmState): the host enqueues max_iterations rounds and
// never synchronises inside a solve.
//
//   k_linearize      thread per map point: residual + analytic 2x(6+3) Jacobians of its observations, loss
//                    weights, the 3x3 point block H_pp, g_p (no atomics: observations are stored point-major)
//   k_cam_blocks     CTA per variable keyframe: 6x6 block H_cc, g_c over the keyframe's observation list
//   k_point_prep     (H_pp + D_p^2)^-1 and (H_pp + D_p^2)^-1 g_p per point
//   k_schur          warp per non-zero 6x6 block (a,b) of the reduced camera system: S_ab = [a==b](H_cc + D_c^2)
//                    - sum over co-observing points of Jc_a' (Jp_a Hpp^-1 Jp_b') Jc_b; no atomics, fixed order
//   k_solve_small    Cholesky of the 6Kv x 6Kv reduced system in the shared memory of one CTA (Kv <= 38): look-ahead
//                    factorisation of the next 6x6 diagonal block, stored as its inverse; candidate keyframe poses
//   k_solve_band     larger systems of a windowed co-visibility graph: persistent CTA over the block band
//   blocked path     anything else (loop-closure edges; also OptimizeEssentialGraph, posegraph.cuh): right-looking
//                    factorisation in HBM inside the row envelope (k_potrf_diag / k_trsm_panel / k_syrk_tile), the whole
//                    back substitution in one launch (k_backsolve_all)
//   k_backsub        thread per point: back-substitution, candidate point, candidate cost of its observations
//   post_lin_body /  gradient / cost reductions and the LM state machine (step quality, radius schedule, termination
//   decide_body      tests): on one GPU they run in the LAST CTA of k_cam_blocks / k_backsub (threadfence + ticket),
//                    sharded solves run them as k_post_lin / k_decide around the NCCL all-reduces
// PoseOptimization (6 unknowns, constant points) and OptimizeSim3 (7 unknowns) are one persistent CTA per problem that
// runs the whole solve.
//
// Everything is deterministic (no floating-point atomics), so 1-GPU and N-GPU runs and repeated runs agree.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ba_device.cuh"
#include "cmos_common.h"

namespace cmos {

#define SYM6(a, b) ((a) * 6 - (a) * ((a) - 1) / 2 + ((b) - (a)))   // a <= b, packed upper triangle of a 6x6

constexpr int kPoseThreads = 256;
constexpr int kLinThreads = 128;
constexpr int kPointLanes = 4;       // lanes per map point in k_linearize / k_backsub (a power of two <= 32)
#ifndef CMOS_CAM_THREADS
#define CMOS_CAM_THREADS 256   // A/B on B200: LocalBA 1.496 (128) -> 1.477 ms (256), 1.497 (512); GlobalBA unchanged
#endif
constexpr int kCamThreads = CMOS_CAM_THREADS;
#ifndef CMOS_SOLVE_THREADS
#define CMOS_SOLVE_THREADS 512
#endif
// Trailing-update variants measured on B200 at configs[3] (tools/ab_solve.py): row-per-warp 1.768 ms per LocalBA solve,
// 4x4 register tiles 1.796 ms, tiles that skip empty column groups 1.95 ms; 1024 threads row-per-warp 1.799 ms.
#ifndef CMOS_SOLVE_TILE
#define CMOS_SOLVE_TILE 0
#endif
#ifndef CMOS_SOLVE_NU
#define CMOS_SOLVE_NU 0
#endif
constexpr int kSolveThreads = CMOS_SOLVE_THREADS;
constexpr int kSmallMaxN = 228;      // packed lower triangle + rhs row + panel buffer of a 228-column system = 223.5 KB
constexpr int kNB = 64;              // panel width of the blocked factorisation
constexpr int kTraceCols = 8;
constexpr size_t kPanelSmem = 2 * kNB * (kNB + 1) * sizeof(double);   // two 64 x 65 fp64 tiles

// =================================================================================================
// PoseOptimization: one CTA per frame, whole LM solve in one launch
// =================================================================================================
struct PoseArgs {
  double* pose7;              // [B][7] in/out
  const int* n_corr;          // [B]
  const double* xw;           // [B][stride][3]
  const float* uv;            // [B][stride][2]
  const float* inv_sigma2;    // [B][stride]
  int stride;
  double fx, fy, cx, cy;
  int max_iterations;
  uint8_t* is_outlier;        // [B][stride]
  int* n_inliers;             // [B]
  cmos_ba_summary* summaries; // [B] or null
  double* trace;              // [B][trace_rows][8] or null
  int trace_rows;
};

__device__ bool chol6_solve(double A[6][6], double* b) {
  for (int j = 0; j < 6; j++) {
    double d = A[j][j];
    for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  for (int i = 0; i < 6; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A[i][k] * b[k];
    b[i] = s / A[i][i];
  }
  for (int i = 5; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < 6; k++) s -= A[k][i] * b[k];
    b[i] = s / A[i][i];
  }
  return true;
}

__global__ void __launch_bounds__(kPoseThreads) k_pose_opt(PoseArgs a) { pdl_begin();
  __shared__ double s_pose[2][7];
  __shared__ double s_red[kPoseThreads / 32][28];
  __shared__ double s_acc[28];        // 21 H (packed upper) | 6 g | cost
  __shared__ double s_scale[6];
  __shared__ double s_scratch[33];
  __shared__ LmState st;
  __shared__ double s_mcc, s_step_norm;
  __shared__ int s_ok;

  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(a.n_corr[f], a.stride);
  const double* X = a.xw + (size_t)f * a.stride * 3;
  const float* uv = a.uv + (size_t)f * a.stride * 2;
  const float* ws = a.inv_sigma2 + (size_t)f * a.stride;
  double* trace = a.trace ? a.trace + (size_t)f * a.trace_rows * kTraceCols : nullptr;
  if (n < 3) {   // CeresOptimizer.cc:330: fewer than 3 correspondences -> return 0, pose untouched
    if (tid == 0) {
      a.n_inliers[f] = 0;
      if (a.summaries) { cmos_ba_summary z; memset(&z, 0, sizeof(z)); a.summaries[f] = z; }
    }
    return;
  }
  if (tid < 7) s_pose[0][tid] = a.pose7[(size_t)f * 7 + tid];
  if (tid == 0) lm_init(st, a.max_iterations, 0);
  __syncthreads();

  for (;;) {
    if (st.need_lin) {
      // ---- EvaluateGradientAndJacobian at x ----
      const double* pose = s_pose[st.cur];
      double acc[28];
#pragma unroll
      for (int k = 0; k < 28; k++) acc[k] = 0.0;
      for (int i = tid; i < n; i += kPoseThreads) {
        const Proj P = project_obs(pose, X + 3 * i, a.fx, a.fy, a.cx, a.cy, (double)uv[2 * i], (double)uv[2 * i + 1], (double)ws[i]);
        double Jc[12], w;
        jac_cam(P, a.fx, a.fy, (double)ws[i], Jc);
        acc[27] += loss_eval(P.r0 * P.r0 + P.r1 * P.r1, 1, &w);
#pragma unroll
        for (int r = 0; r < 6; r++) {
          const double j0 = w * Jc[r], j1 = w * Jc[6 + r];
#pragma unroll
          for (int c = r; c < 6; c++) acc[SYM6(r, c)] += j0 * Jc[c] + j1 * Jc[6 + c];
          acc[21 + r] += j0 * P.r0 + j1 * P.r1;
        }
      }
#pragma unroll
      for (int k = 0; k < 28; k++) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_red[warp][k] = v;
      }
      __syncthreads();
      if (tid < 28) {
        double t = 0.0;
        for (int w2 = 0; w2 < kPoseThreads / 32; w2++) t += s_red[w2][tid];
        s_acc[tid] = t;
      }
      __syncthreads();
      if (tid == 0) {
        if (st.first)
          for (int k = 0; k < 6; k++) s_scale[k] = 1.0 / (1.0 + sqrt(s_acc[SYM6(k, k)]));
        double xn = 0.0;
        for (int k = 0; k < 7; k++) xn += pose[k] * pose[k];
        lm_after_linearize(st, s_acc[27], pose_gradient_max(pose, s_acc + 21), sqrt(xn), trace);
      }
      __syncthreads();
    }
    if (st.done) break;
    // ---- LevenbergMarquardtStrategy::ComputeStep on the Jacobi-scaled system (thread 0: 6x6) ----
    if (tid == 0) {
      const double* pose = s_pose[st.cur];
      double* cand = s_pose[st.cur ^ 1];
      double A[6][6], b[6], D2[6], gs[6];
      for (int r = 0; r < 6; r++) {
        D2[r] = fmin(fmax(s_scale[r] * s_scale[r] * s_acc[SYM6(r, r)], kMinLmDiag), kMaxLmDiag) / st.radius;
        gs[r] = s_scale[r] * s_acc[21 + r];
        b[r] = gs[r];
        for (int c = 0; c < 6; c++) {
          const int lo = r < c ? r : c, hi = r < c ? c : r;
          A[r][c] = s_scale[r] * s_scale[c] * s_acc[SYM6(lo, hi)] + (r == c ? D2[r] : 0.0);
        }
      }
      bool ok = chol6_solve(A, b);
      double mcc = 0.0, delta[6];
      for (int r = 0; r < 6; r++) {
        const double step = -b[r];
        if (!isfinite(step)) ok = false;
        mcc += step * (D2[r] * step - gs[r]);
        delta[r] = step * s_scale[r];
      }
      mcc *= 0.5;
      double sn = 0.0;
      if (ok) {
        for (int k = 0; k < 3; k++) cand[k] = pose[k] + delta[k];
        quat_plus(pose + 3, delta + 3, cand + 3);
        for (int k = 0; k < 7; k++) sn += (pose[k] - cand[k]) * (pose[k] - cand[k]);
      }
      s_ok = ok; s_mcc = mcc; s_step_norm = sqrt(sn);
    }
    __syncthreads();
    double cand_cost = 0.0;
    if (s_ok && s_mcc > 0.0) {
      const double* cand = s_pose[st.cur ^ 1];
      double c = 0.0, w;
      for (int i = tid; i < n; i += kPoseThreads) {
        const Proj P = project_obs(cand, X + 3 * i, a.fx, a.fy, a.cx, a.cy, (double)uv[2 * i], (double)uv[2 * i + 1], (double)ws[i]);
        c += loss_eval(P.r0 * P.r0 + P.r1 * P.r1, 1, &w);
      }
      cand_cost = block_sum(c, s_scratch);
    }
    if (tid == 0) lm_decide(st, s_ok != 0, s_mcc, cand_cost, s_step_norm, trace);
    __syncthreads();
    if (st.done) break;
  }
  // ---- CheckOutliers (:243-269) on the optimised pose, then q.normalized() (:335) ----
  const double* pose = s_pose[st.cur];
  int bad = 0;
  for (int i = tid; i < n; i += kPoseThreads) {
    const Proj P = project_obs(pose, X + 3 * i, a.fx, a.fy, a.cx, a.cy, (double)uv[2 * i], (double)uv[2 * i + 1], 1.0);
    const double chi2 = (P.r0 * P.r0 + P.r1 * P.r1) * (double)ws[i];
    const bool out = chi2 > kChi2;
    a.is_outlier[(size_t)f * a.stride + i] = out;
    bad += out;
  }
  const double nbad = block_sum((double)bad, s_scratch);
  if (tid == 0) {
    a.n_inliers[f] = n - (int)nbad;
    const double nq = sqrt(pose[3] * pose[3] + pose[4] * pose[4] + pose[5] * pose[5] + pose[6] * pose[6]);
    for (int k = 0; k < 3; k++) a.pose7[(size_t)f * 7 + k] = pose[k];
    for (int k = 3; k < 7; k++) a.pose7[(size_t)f * 7 + k] = pose[k] / nq;
    if (a.summaries) {
      cmos_ba_summary s;
      s.iterations = st.iteration; s.successful_steps = st.successful; s.termination = st.termination;
      s.jacobian_evaluations = st.jac_evals; s.initial_cost = st.initial_cost; s.final_cost = st.x_cost;
      a.summaries[f] = s;
    }
  }
}

// =================================================================================================
// OptimizeSim3 (CeresOptimizer.cc:601-735): one CTA runs the whole LM solve over the 7-vector sim12
// =================================================================================================
struct Sim3Args {
  int n;                       // correspondences (2 residual blocks each)
  double s12, R12[9], t12[3];  // initial S12
  double K1[4], K2[4];
  const float* obs1; const float* inv_sigma1; const double* P3D2c;
  const float* obs2; const float* inv_sigma2; const double* P3D1c;
  double huber_a;
  int max_iterations;
  uint8_t* is_bad;             // [n]
  double* out;                 // [24]: lie(7), s, R(9), t(3), return value, iterations, successful, termination
  cmos_ba_summary* summary;
  double* trace;
};

template <int N>
__device__ bool chol_solve_n(double (*A)[N], double* b) {
  for (int j = 0; j < N; j++) {
    double d = A[j][j];
    for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < N; i++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  for (int i = 0; i < N; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A[i][k] * b[k];
    b[i] = s / A[i][i];
  }
  for (int i = N - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < N; k++) s -= A[k][i] * b[k];
    b[i] = s / A[i][i];
  }
  return true;
}

#define SYM7(r, c) ((r) * 7 - (r) * ((r) - 1) / 2 + ((c) - (r)))   // packed upper triangle of a 7x7, r <= c

__device__ __forceinline__ double huber_eval(double s, double a, double* w) {
  const double b = a * a;
  if (s > b) { const double r = sqrt(s); *w = fmax(2.2250738585072014e-308, a / r); return 0.5 * (2.0 * a * r - b); }
  *w = 1.0;
  return 0.5 * s;
}

__global__ void __launch_bounds__(kPoseThreads) k_sim3_opt(Sim3Args a) { pdl_begin();
  __shared__ double s_x[2][7];
  __shared__ Sim3D s_S, s_Si;            // exp(x) and its inverse for the point being evaluated
  __shared__ double s_red[kPoseThreads / 32][36];
  __shared__ double s_acc[36];           // 28 H (packed upper) | 7 g | cost
  __shared__ double s_scale[7];
  __shared__ double s_scratch[33];
  __shared__ LmState st;
  __shared__ double s_mcc, s_step_norm;
  __shared__ int s_ok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = a.n;
  if (tid == 0) {
    Sim3D S0; S0.s = a.s12;
    for (int i = 0; i < 9; i++) S0.R[i] = a.R12[i];
    for (int i = 0; i < 3; i++) S0.t[i] = a.t12[i];
    sim3_log(S0, s_x[0]);
    lm_init(st, a.max_iterations, 0);
  }
  __syncthreads();
  auto set_point = [&](const double* x) {          // thread 0
    sim3_exp(x, s_S);
    sim3_inverse(s_S, s_Si);
  };
  auto block_eval = [&](int b, double* r, double* J) {
    const int i = b >> 1;
    if ((b & 1) == 0)
      sim3_error_term(s_S, a.P3D2c + 3 * i, a.K1[0], a.K1[1], a.K1[2], a.K1[3], (double)a.obs1[2 * i], (double)a.obs1[2 * i + 1],
                      (double)a.inv_sigma1[i], r, J);
    else
      sim3_error_term(s_Si, a.P3D1c + 3 * i, a.K2[0], a.K2[1], a.K2[2], a.K2[3], (double)a.obs2[2 * i], (double)a.obs2[2 * i + 1],
                      (double)a.inv_sigma2[i], r, J);
  };
  for (;;) {
    if (st.need_lin) {
      if (tid == 0) set_point(s_x[st.cur]);
      __syncthreads();
      double acc[36];
#pragma unroll
      for (int k = 0; k < 36; k++) acc[k] = 0.0;
      for (int b = tid; b < 2 * n; b += kPoseThreads) {
        double r[2], J[14], w;
        block_eval(b, r, J);
        acc[35] += huber_eval(r[0] * r[0] + r[1] * r[1], a.huber_a, &w);
#pragma unroll
        for (int p = 0; p < 7; p++) {
          const double j0 = w * J[p], j1 = w * J[7 + p];
#pragma unroll
          for (int c = p; c < 7; c++) acc[SYM7(p, c)] += j0 * J[c] + j1 * J[7 + c];
          acc[28 + p] += j0 * r[0] + j1 * r[1];
        }
      }
#pragma unroll
      for (int k = 0; k < 36; k++) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_red[warp][k] = v;
      }
      __syncthreads();
      if (tid < 36) {
        double t = 0.0;
        for (int w2 = 0; w2 < kPoseThreads / 32; w2++) t += s_red[w2][tid];
        s_acc[tid] = t;
      }
      __syncthreads();
      if (tid == 0) {
        const double* x = s_x[st.cur];
        if (st.first)
          for (int k = 0; k < 7; k++) s_scale[k] = 1.0 / (1.0 + sqrt(s_acc[SYM7(k, k)]));
        double ng[7], xp[7], xn = 0.0, gm = 0.0;
        for (int k = 0; k < 7; k++) { ng[k] = -s_acc[28 + k]; xn += x[k] * x[k]; }
        sim3_plus(x, ng, xp);
        for (int k = 0; k < 7; k++) gm = fmax(gm, fabs(x[k] - xp[k]));
        lm_after_linearize(st, s_acc[35], gm, sqrt(xn), a.trace);
      }
      __syncthreads();
    }
    if (st.done) break;
    if (tid == 0) {
      const double* x = s_x[st.cur];
      double* cand = s_x[st.cur ^ 1];
      double A[7][7], b[7], D2[7], gs[7];
      for (int r = 0; r < 7; r++) {
        D2[r] = fmin(fmax(s_scale[r] * s_scale[r] * s_acc[SYM7(r, r)], kMinLmDiag), kMaxLmDiag) / st.radius;
        gs[r] = s_scale[r] * s_acc[28 + r];
        b[r] = gs[r];
        for (int c = 0; c < 7; c++) {
          const int lo = r < c ? r : c, hi = r < c ? c : r;
          A[r][c] = s_scale[r] * s_scale[c] * s_acc[SYM7(lo, hi)] + (r == c ? D2[r] : 0.0);
        }
      }
      bool ok = chol_solve_n<7>(A, b);
      double mcc = 0.0, delta[7];
      for (int r = 0; r < 7; r++) {
        const double step = -b[r];
        if (!isfinite(step)) ok = false;
        mcc += step * (D2[r] * step - gs[r]);
        delta[r] = step * s_scale[r];
      }
      mcc *= 0.5;
      double sn = 0.0;
      if (ok) {
        sim3_plus(x, delta, cand);
        for (int k = 0; k < 7; k++) sn += (x[k] - cand[k]) * (x[k] - cand[k]);
        set_point(cand);
      }
      s_ok = ok; s_mcc = mcc; s_step_norm = sqrt(sn);
    }
    __syncthreads();
    double cand_cost = 0.0;
    if (s_ok && s_mcc > 0.0) {
      double c = 0.0, w;
      for (int b = tid; b < 2 * n; b += kPoseThreads) {
        double r[2];
        block_eval(b, r, nullptr);
        c += huber_eval(r[0] * r[0] + r[1] * r[1], a.huber_a, &w);
      }
      cand_cost = block_sum(c, s_scratch);
    }
    if (tid == 0) lm_decide(st, s_ok != 0, s_mcc, cand_cost, s_step_norm, a.trace);
    __syncthreads();
    if (st.done) break;
  }
  // S12 = exp(sim12); CheckOutlier with Eigen::Quaterniond(s R) — Eigen's matrix -> quaternion conversion applied to the
  // SCALED rotation, then its unit-quaternion rotation polynomial (CeresOptimizer.cc:702-726, 227-241)
  if (tid == 0) set_point(s_x[st.cur]);
  __syncthreads();
  auto check = [&](const Sim3D& T, const double* K4, double u, double v, double inv_sigma, const double* P) {
    double M[9], q[4];
    for (int i = 0; i < 9; i++) M[i] = T.s * T.R[i];
    const double tr = M[0] + M[4] + M[8];
    if (tr > 0) {
      double t = sqrt(tr + 1.0);
      q[3] = 0.5 * t; t = 0.5 / t;
      q[0] = (M[7] - M[5]) * t; q[1] = (M[2] - M[6]) * t; q[2] = (M[3] - M[1]) * t;
    } else {
      int i = 0;
      if (M[4] > M[0]) i = 1;
      if (M[8] > M[4 * i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      double t = sqrt(M[4 * i] - M[4 * j] - M[4 * k] + 1.0);
      q[i] = 0.5 * t; t = 0.5 / t;
      q[3] = (M[3 * k + j] - M[3 * j + k]) * t; q[j] = (M[3 * j + i] + M[3 * i + j]) * t; q[k] = (M[3 * k + i] + M[3 * i + k]) * t;
    }
    const double uv0 = 2 * (q[1] * P[2] - q[2] * P[1]), uv1 = 2 * (q[2] * P[0] - q[0] * P[2]), uv2 = 2 * (q[0] * P[1] - q[1] * P[0]);
    const double p0 = P[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1) + T.t[0];
    const double p1 = P[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2) + T.t[1];
    const double p2 = P[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0) + T.t[2];
    const double px = K4[0] * p0 + K4[2] * p2, py = K4[1] * p1 + K4[3] * p2;
    const double eu = u - px / p2, ev = v - py / p2;
    return (eu * eu + ev * ev) * inv_sigma > a.huber_a * a.huber_a;
  };
  int bad = 0;
  for (int i = tid; i < n; i += kPoseThreads) {
    const bool b12 = check(s_S, a.K1, (double)a.obs1[2 * i], (double)a.obs1[2 * i + 1], (double)a.inv_sigma1[i], a.P3D2c + 3 * i);
    const bool b21 = check(s_Si, a.K2, (double)a.obs2[2 * i], (double)a.obs2[2 * i + 1], (double)a.inv_sigma2[i], a.P3D1c + 3 * i);
    a.is_bad[i] = b12 || b21;
    bad += b12 || b21;
  }
  const double nbad = block_sum((double)bad, s_scratch);
  if (tid == 0) {
    const double* x = s_x[st.cur];
    for (int k = 0; k < 7; k++) a.out[k] = x[k];
    a.out[7] = s_S.s;
    for (int k = 0; k < 9; k++) a.out[8 + k] = s_S.R[k];
    for (int k = 0; k < 3; k++) a.out[17 + k] = s_S.t[k];
    const int good = n - (int)nbad;
    a.out[20] = good < 10 ? 0.0 : (double)good;
    if (a.summary) {
      cmos_ba_summary s;
      s.iterations = st.iteration; s.successful_steps = st.successful; s.termination = st.termination;
      s.jacobian_evaluations = st.jac_evals; s.initial_cost = st.initial_cost; s.final_cost = st.x_cost;
      *a.summary = s;
    }
  }
}

// =================================================================================================
// General engine (LocalBundleAdjustment / BundleAdjustment)
// =================================================================================================
struct BaDev {
  int K, Kv, M, N, n_blocks, nc;            // nc = 6 * Kv
  double fx, fy, cx, cy;
  double* cams[2];                          // [K][7]
  double* pts[2];                           // [M][3]
  const int* cam_var;                       // [K]
  const int *o_cam, *o_cv, *o_pt;           // [N] point-major
  const float2* o_uv;
  const float* o_w;
  uint8_t* o_mode;
  const int* pt_start;                      // [M+1]
  const int *cam_start, *cam_obs;           // [Kv+1], [#obs of variable keyframes]
  const int *blk_a, *blk_b, *blk_start;     // [n_blocks], [n_blocks], [n_blocks+1]
  const int *chunk_start, *chunk_blk, *blk_chunk0;   // [n_chunks+1] pair ranges, [n_chunks] block of a chunk, [n_blocks+1]
  double* schur_part;                       // [n_chunks][42] per-chunk partial sums of k_schur_chunks
  int n_chunks;
  const int *pair_a, *pair_b;               // observation pairs of every block
  double *Jc, *Jp, *res;                    // [N][12], [N][6], [N][2]  (rows already weighted by sqrt(rho'))
  double *Hpp, *gp, *Hinv, *tp, *scale_p;   // [M][6], [M][3], [M][6], [M][3], [M][3]
  double *Hcc, *gc, *scale_c;               // [Kv][21], [Kv][6], [Kv][6]
  double *S, *rhs, *yc;                     // [nc][nc] lower, [nc], [nc]
  double *Sblk;                             // [n_blocks][36] reduced-system blocks (a,b) as assembled by k_schur
  const int* var_cam;                       // [Kv] variable index -> keyframe
  double* red;                              // [8] locally reduced scalars (all-reduced across ranks when multi)
  unsigned int* ticket;                     // [2] arrival counters of the fused reduce-and-decide tails (single GPU)
  int multi, is_root;                       // multi: points are sharded over ranks; is_root: adds the keyframe-only terms
  int rank, n_ranks;
  double* lin_tail;                         // multi: [cost, |x|^2, gmax slot per rank] behind H_cc / g_c in the same message
  double *part;                             // partial sums, see offsets
  int n_lin_blocks;
  // offsets into part
  int o_lin_cost, o_lin_gmax, o_lin_xn2, o_cam_gmax, o_cam_xn2, o_bs_cost, o_bs_mcc, o_bs_sn2, o_cam_mcc, o_cam_sn2;
  LmState* st;
  double* trace;                            // [rows][8] of the running pass, or null
  const volatile uint8_t* stop_flag;        // mapped host byte or null
  int pass, dbg_stop_pass, dbg_stop_iter;   // test hook: the device raises the stop flag itself at (pass, iteration)
};

// `pad` of LmState doubles as the "aborted" flag: LocalBundleAdjustment returns without writing anything back
// when the stop flag is up at the start of a pass (CeresOptimizer.cc:509-512).
__global__ void k_lm_init(BaDev d, int max_iterations) { pdl_begin();
  const int cur = d.st->cur, aborted = d.st->pad;
  lm_init(*d.st, max_iterations, cur & 1);
  d.st->pad = aborted;
  if (aborted || (d.stop_flag && *d.stop_flag)) { d.st->pad = 1; d.st->done = 1; d.st->termination = TERM_USER; }
}

__device__ bool last_cta_arrives(unsigned int* ticket, unsigned int n_ctas);
// cmos_ba_debug_stop_at: emulates another thread raising the stop flag at a reproducible point of the solve
__device__ __forceinline__ void dbg_raise_stop(const BaDev& d, const LmState& st) {
  if (d.dbg_stop_iter >= 0 && d.stop_flag && d.pass == d.dbg_stop_pass && st.iteration >= d.dbg_stop_iter)
    *const_cast<volatile uint8_t*>(d.stop_flag) = 1;
}
__device__ void post_lin_body(const BaDev& d, LmState& st, int phase, double* scratch);
__device__ void decide_body(const BaDev& d, LmState& st, int phase, double* scratch);

__global__ void __launch_bounds__(kLinThreads) k_linearize(BaDev d) { pdl_begin();
  __shared__ double scratch[33];
  const LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  // kPointLanes lanes share a map point: each takes every kPointLanes-th observation (the dependent loads and the fp64
  // chain of a point's 4-5 observations run side by side instead of one after the other), two shuffles add the parts
  const int gt = blockIdx.x * kLinThreads + threadIdx.x;
  const int j = gt / kPointLanes, q = gt % kPointLanes;
  const double* cams = d.cams[st.cur];
  const double* pts = d.pts[st.cur];
  double cost = 0.0, gmax = 0.0, xn2 = 0.0;
  double H[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  double X[3] = {0, 0, 0};
  if (j < d.M) {
    X[0] = pts[3 * j]; X[1] = pts[3 * j + 1]; X[2] = pts[3 * j + 2];
    for (int p = d.pt_start[j] + q; p < d.pt_start[j + 1]; p += kPointLanes) {
      const int cam = d.o_cam[p];
      const float2 uv = d.o_uv[p];
      const double w = (double)d.o_w[p];
      const double* c = cams + 7 * (size_t)cam;
      const Proj P = project_obs(c, X, d.fx, d.fy, d.cx, d.cy, (double)uv.x, (double)uv.y, w);
      double wsum;
      cost += loss_eval(P.r0 * P.r0 + P.r1 * P.r1, d.o_mode[p], &wsum);
      const double sw = sqrt(wsum);
      double Jp[6];
      jac_point(P, c, d.fx, d.fy, w, Jp);
#pragma unroll
      for (int k = 0; k < 6; k++) Jp[k] *= sw;
      const double r0 = sw * P.r0, r1 = sw * P.r1;
      H[0] += Jp[0] * Jp[0] + Jp[3] * Jp[3]; H[1] += Jp[0] * Jp[1] + Jp[3] * Jp[4]; H[2] += Jp[0] * Jp[2] + Jp[3] * Jp[5];
      H[3] += Jp[1] * Jp[1] + Jp[4] * Jp[4]; H[4] += Jp[1] * Jp[2] + Jp[4] * Jp[5]; H[5] += Jp[2] * Jp[2] + Jp[5] * Jp[5];
      g[0] += Jp[0] * r0 + Jp[3] * r1; g[1] += Jp[1] * r0 + Jp[4] * r1; g[2] += Jp[2] * r0 + Jp[5] * r1;
      double2* jp = (double2*)(d.Jp + 6 * (size_t)p);
      jp[0] = make_double2(Jp[0], Jp[1]); jp[1] = make_double2(Jp[2], Jp[3]); jp[2] = make_double2(Jp[4], Jp[5]);
      *(double2*)(d.res + 2 * (size_t)p) = make_double2(r0, r1);
      if (d.o_cv[p] >= 0) {
        double Jc[12];
        jac_cam(P, d.fx, d.fy, w, Jc);
        double2* jc = (double2*)(d.Jc + 12 * (size_t)p);
#pragma unroll
        for (int k = 0; k < 6; k++) jc[k] = make_double2(sw * Jc[2 * k], sw * Jc[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < kPointLanes; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 6; k++) H[k] += __shfl_xor_sync(0xffffffffu, H[k], o);
#pragma unroll
    for (int k = 0; k < 3; k++) g[k] += __shfl_xor_sync(0xffffffffu, g[k], o);
  }
  if (j < d.M && q == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) d.Hpp[6 * (size_t)j + k] = H[k];
#pragma unroll
    for (int k = 0; k < 3; k++) d.gp[3 * (size_t)j + k] = g[k];
    if (st.first) {
      d.scale_p[3 * (size_t)j] = 1.0 / (1.0 + sqrt(H[0]));
      d.scale_p[3 * (size_t)j + 1] = 1.0 / (1.0 + sqrt(H[3]));
      d.scale_p[3 * (size_t)j + 2] = 1.0 / (1.0 + sqrt(H[5]));
    }
    gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
    xn2 = X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
  }
  cost = block_sum(cost, scratch);
  gmax = block_max(gmax, scratch);
  xn2 = block_sum(xn2, scratch);
  if (threadIdx.x == 0) {
    d.part[d.o_lin_cost + blockIdx.x] = cost;
    d.part[d.o_lin_gmax + blockIdx.x] = gmax;
    d.part[d.o_lin_xn2 + blockIdx.x] = xn2;
  }
}

__device__ void cam_finish(const BaDev& d, const LmState& st, int a) {
  const double* Hc = d.Hcc + 21 * (size_t)a;
  const double* g = d.gc + 6 * (size_t)a;
  if (st.first)
    for (int k = 0; k < 6; k++) d.scale_c[6 * (size_t)a + k] = 1.0 / (1.0 + sqrt(Hc[SYM6(k, k)]));
  const double* c = d.cams[st.cur] + 7 * (size_t)d.var_cam[a];
  double xn2 = 0.0;
  for (int k = 0; k < 7; k++) xn2 += c[k] * c[k];
  d.part[d.o_cam_gmax + a] = pose_gradient_max(c, g);
  d.part[d.o_cam_xn2 + a] = xn2;
}

// n_cam_ctas: CTAs of the launch that run this body (the ticket of the fused tail counts them)
__device__ __forceinline__ void cam_blocks_body(const BaDev& d, int n_cam_ctas) {
  __shared__ double s_red[kCamThreads / 32][27];     // 108 doubles; reused as the 33-double scratch of the fused tail
  const LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0.0;
  for (int e = d.cam_start[a] + tid; e < d.cam_start[a + 1]; e += kCamThreads) {
    const int p = d.cam_obs[e];
    double Jc[12];
    const double2* jc = (const double2*)(d.Jc + 12 * (size_t)p);
#pragma unroll
    for (int k = 0; k < 6; k++) { const double2 v = jc[k]; Jc[2 * k] = v.x; Jc[2 * k + 1] = v.y; }
    const double2 r = *(const double2*)(d.res + 2 * (size_t)p);
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
#pragma unroll
      for (int c = rr; c < 6; c++) acc[SYM6(rr, c)] += Jc[rr] * Jc[c] + Jc[6 + rr] * Jc[6 + c];
      acc[21 + rr] += Jc[rr] * r.x + Jc[6 + rr] * r.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (tid < 27) {
    double t = 0.0;
    for (int w = 0; w < kCamThreads / 32; w++) t += s_red[w][tid];
    if (tid < 21) d.Hcc[21 * (size_t)a + tid] = t; else d.gc[6 * (size_t)a + tid - 21] = t;
  }
  __syncthreads();
  if (d.multi) return;
  if (tid == 0) cam_finish(d, st, a);
  if (last_cta_arrives(d.ticket, n_cam_ctas)) post_lin_body(d, *d.st, 0, &s_red[0][0]);
}
__global__ void __launch_bounds__(kCamThreads) k_cam_blocks(BaDev d) { pdl_begin();
  cam_blocks_body(d, gridDim.x);
}

// Jacobi scale, gradient norm and parameter norm of one variable keyframe from its (global) H_cc, g_c
__global__ void __launch_bounds__(128) k_cam_finish(BaDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done || !st.need_lin) return;
  const int a = blockIdx.x * 128 + threadIdx.x;
  if (a < d.Kv) cam_finish(d, st, a);
}

// strided, fixed-order sum / max of a partial array by one CTA
// (ld.cg: the partials may have been written by other CTAs of the same launch — see last_cta_arrives)
__device__ double part_sum(const double* p, int n, double* scratch) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += __ldcg(p + i);
  return block_sum(v, scratch);
}
__device__ double part_max(const double* p, int n, double* scratch) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v = fmax(v, __ldcg(p + i));
  return block_max(v, scratch);
}
// Reduce-and-decide tails run inside the producing kernel on a single GPU: every CTA publishes its partials, fences and
// takes a ticket; the CTA that draws the last one sees all of them (threadfence reduction) and runs the tail.  One launch
// less per stage, and the order of every sum is unchanged.
__device__ bool last_cta_arrives(unsigned int* ticket, unsigned int n_ctas) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == n_ctas - 1);
    if (s_last) *ticket = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// phase 1: reduce this rank's per-CTA partials into red[0..2] = cost, |x|^2 of the points, gradient max;
// phase 2: add the keyframes' share and run the state machine; phase 0: both (single GPU)
__device__ void post_lin_body(const BaDev& d, LmState& st, int phase, double* scratch) {
  if (phase != 2) {
    const double cost = part_sum(d.part + d.o_lin_cost, d.n_lin_blocks, scratch);
    const double x1 = part_sum(d.part + d.o_lin_xn2, d.n_lin_blocks, scratch);
    const double g1 = part_max(d.part + d.o_lin_gmax, d.n_lin_blocks, scratch);
    if (threadIdx.x == 0) {
      if (d.multi) {   // rides behind H_cc / g_c in ONE sum all-reduce: the max becomes a slot per rank
        d.lin_tail[0] = cost; d.lin_tail[1] = x1;
        for (int r = 0; r < d.n_ranks; r++) d.lin_tail[2 + r] = r == d.rank ? g1 : 0.0;
      } else {
        d.red[0] = cost; d.red[1] = x1; d.red[2] = g1;
      }
    }
    __syncthreads();
  }
  if (phase != 1) {
    const double g2 = part_max(d.part + d.o_cam_gmax, d.Kv, scratch);
    const double x2 = part_sum(d.part + d.o_cam_xn2, d.Kv, scratch);
    if (threadIdx.x == 0) {
      if (d.multi) {
        double g = 0.0;
        for (int r = 0; r < d.n_ranks; r++) g = fmax(g, d.lin_tail[2 + r]);
        d.red[0] = d.lin_tail[0]; d.red[1] = d.lin_tail[1]; d.red[2] = g;
      }
      lm_after_linearize(st, d.red[0], fmax(d.red[2], g2), sqrt(d.red[1] + x2), d.trace);
      dbg_raise_stop(d, st);
      if (!st.done && d.stop_flag && *d.stop_flag) { st.done = 1; st.termination = TERM_USER; }   // StopFlagCallback
    }
  }
}
__global__ void __launch_bounds__(256) k_post_lin(BaDev d, int phase) { pdl_begin();
  __shared__ double scratch[33];
  LmState& st = *d.st;
  if (st.done) return;
  if (!st.need_lin) {
    // Rejected step: nothing was re-linearised, but the collective still runs.  The root's message buffer holds the
    // totals of the last linearisation; every other rank contributes zeros so that the sum reproduces them (summing the
    // already reduced buffers would multiply H_cc / g_c by the number of ranks).
    if (phase == 1 && d.multi && !d.is_root)
      for (int i = threadIdx.x; i < 27 * d.Kv + 2 + d.n_ranks; i += blockDim.x) d.Hcc[i] = 0.0;
    return;
  }
  post_lin_body(d, st, phase, scratch);
}

__device__ __forceinline__ void point_prep_body(const BaDev& d, int j) {
  LmState& st = *d.st;
  if (st.done) return;
  if (j >= d.M) return;
  const double* H = d.Hpp + 6 * (size_t)j;
  const double s0 = d.scale_p[3 * (size_t)j], s1 = d.scale_p[3 * (size_t)j + 1], s2 = d.scale_p[3 * (size_t)j + 2];
  const double ir = 1.0 / st.radius;
  (void)ir;
  double A[6];
  A[0] = s0 * s0 * H[0]; A[1] = s0 * s1 * H[1]; A[2] = s0 * s2 * H[2];
  A[3] = s1 * s1 * H[3]; A[4] = s1 * s2 * H[4]; A[5] = s2 * s2 * H[5];
  A[0] += fmin(fmax(A[0], kMinLmDiag), kMaxLmDiag) / st.radius;
  A[3] += fmin(fmax(A[3], kMinLmDiag), kMaxLmDiag) / st.radius;
  A[5] += fmin(fmax(A[5], kMinLmDiag), kMaxLmDiag) / st.radius;
  double inv[6];
  if (!invert3_sym(A, inv)) {
    st.solve_failed = 1;   // benign race: every writer stores 1
#pragma unroll
    for (int k = 0; k < 6; k++) inv[k] = 0.0;
  }
  const double g0 = s0 * d.gp[3 * (size_t)j], g1 = s1 * d.gp[3 * (size_t)j + 1], g2 = s2 * d.gp[3 * (size_t)j + 2];
#pragma unroll
  for (int k = 0; k < 6; k++) d.Hinv[6 * (size_t)j + k] = inv[k];
  d.tp[3 * (size_t)j] = inv[0] * g0 + inv[1] * g1 + inv[2] * g2;
  d.tp[3 * (size_t)j + 1] = inv[1] * g0 + inv[3] * g1 + inv[4] * g2;
  d.tp[3 * (size_t)j + 2] = inv[2] * g0 + inv[4] * g1 + inv[5] * g2;
}
__global__ void __launch_bounds__(kLinThreads) k_point_prep(BaDev d) { pdl_begin();
  point_prep_body(d, blockIdx.x * kLinThreads + threadIdx.x);
}
// Both consumers of a linearisation in ONE launch (single-GPU solves): CTAs [0, Kv) sum the camera blocks (the last of them runs
// the reduce-and-decide tail), the CTAs behind them invert the damped point blocks.  The two parts read what k_linearize wrote
// and the trust-region radius, which the tail does not touch; a point CTA that still sees done == 0 while the tail raises it
// computes blocks nobody reads.  One launch less per LM iteration, and the 5 us of point work run beside the 13 us of camera work.
__global__ void __launch_bounds__(kCamThreads) k_cam_blocks_prep(BaDev d) { pdl_begin();
  if ((int)blockIdx.x < d.Kv) cam_blocks_body(d, d.Kv);
  else point_prep_body(d, ((int)blockIdx.x - d.Kv) * kCamThreads + (int)threadIdx.x);
}

__device__ __forceinline__ void load_jc_scaled(const BaDev& d, int p, const double* sc, double* Jc) {
  const double2* jc = (const double2*)(d.Jc + 12 * (size_t)p);
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 v = jc[k]; Jc[2 * k] = v.x; Jc[2 * k + 1] = v.y; }
#pragma unroll
  for (int k = 0; k < 6; k++) { Jc[k] *= sc[k]; Jc[6 + k] *= sc[k]; }
}
__device__ __forceinline__ void load_jp_scaled(const BaDev& d, int p, int j, double* Jp) {
  const double2* jp = (const double2*)(d.Jp + 6 * (size_t)p);
  const double2 v0 = jp[0], v1 = jp[1], v2 = jp[2];
  const double s0 = d.scale_p[3 * (size_t)j], s1 = d.scale_p[3 * (size_t)j + 1], s2 = d.scale_p[3 * (size_t)j + 2];
  Jp[0] = v0.x * s0; Jp[1] = v0.y * s1; Jp[2] = v1.x * s2; Jp[3] = v1.y * s0; Jp[4] = v2.x * s1; Jp[5] = v2.y * s2;
}

// One co-observation pair (pa, pb) of block (a, b): acc += Jc_a' (Jp_a Hpp^-1 Jp_b') Jc_b; the diagonal pair also feeds the rhs.
__device__ __forceinline__ void schur_pair(const BaDev& d, const int e, const double* sca, const double* scb, double* acc,
                                           double* racc) {
    const int pa = d.pair_a[e], pb = d.pair_b[e];
    const int j = d.o_pt[pa];
    double Jca[12], Jpa[6], Jpb[6], U[12];
    load_jc_scaled(d, pa, sca, Jca);
    load_jp_scaled(d, pa, j, Jpa);
    const double* Hi = d.Hinv + 6 * (size_t)j;
    const double h0 = Hi[0], h1 = Hi[1], h2 = Hi[2], h3 = Hi[3], h4 = Hi[4], h5 = Hi[5];
    // V = Jp_a * Hinv (2x3)
    const double V00 = Jpa[0] * h0 + Jpa[1] * h1 + Jpa[2] * h2, V01 = Jpa[0] * h1 + Jpa[1] * h3 + Jpa[2] * h4,
                 V02 = Jpa[0] * h2 + Jpa[1] * h4 + Jpa[2] * h5;
    const double V10 = Jpa[3] * h0 + Jpa[4] * h1 + Jpa[5] * h2, V11 = Jpa[3] * h1 + Jpa[4] * h3 + Jpa[5] * h4,
                 V12 = Jpa[3] * h2 + Jpa[4] * h4 + Jpa[5] * h5;
    if (pa == pb) {
#pragma unroll
      for (int k = 0; k < 6; k++) Jpb[k] = Jpa[k];
#pragma unroll
      for (int k = 0; k < 12; k++) U[k] = Jca[k];
      // rhs: Jc_a' (Jp_a (Hpp^-1 g_p))
      const double* t = d.tp + 3 * (size_t)j;
      const double q0 = Jpa[0] * t[0] + Jpa[1] * t[1] + Jpa[2] * t[2], q1 = Jpa[3] * t[0] + Jpa[4] * t[1] + Jpa[5] * t[2];
#pragma unroll
      for (int k = 0; k < 6; k++) racc[k] += Jca[k] * q0 + Jca[6 + k] * q1;
    } else {
      load_jp_scaled(d, pb, j, Jpb);
      load_jc_scaled(d, pb, scb, U);
    }
    // Q = V * Jp_b' (2x2)
    const double Q00 = V00 * Jpb[0] + V01 * Jpb[1] + V02 * Jpb[2], Q01 = V00 * Jpb[3] + V01 * Jpb[4] + V02 * Jpb[5];
    const double Q10 = V10 * Jpb[0] + V11 * Jpb[1] + V12 * Jpb[2], Q11 = V10 * Jpb[3] + V11 * Jpb[4] + V12 * Jpb[5];
    // U = Q * Jc_b (2x6), acc += Jc_a' U
    double U0[6], U1[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { U0[k] = Q00 * U[k] + Q01 * U[6 + k]; U1[k] = Q10 * U[k] + Q11 * U[6 + k]; }
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) acc[r * 6 + c] += Jca[r] * U0[c] + Jca[6 + r] * U1[c];
}

// CTA (4 warps) per non-zero block (a,b), a <= b (variable-keyframe indices): threads stride over the block's
// observation pairs, warp shuffles + a fixed-order cross-warp sum reduce the 6x6 (+ rhs); writes the
// lower-triangle copy S[b][a].
constexpr int kSchurThreads = 128;
__global__ void __launch_bounds__(kSchurThreads) k_schur(BaDev d) { pdl_begin();
  __shared__ double s_part[kSchurThreads / 32][42];
  const LmState& st = *d.st;
  if (st.done) return;
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = d.blk_a[blk], b = d.blk_b[blk];
  double sca[6], scb[6];
#pragma unroll
  for (int k = 0; k < 6; k++) { sca[k] = d.scale_c[6 * (size_t)a + k]; scb[k] = d.scale_c[6 * (size_t)b + k]; }
  double acc[36], racc[6];
#pragma unroll
  for (int k = 0; k < 36; k++) acc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) racc[k] = 0.0;
  for (int e = d.blk_start[blk] + tid; e < d.blk_start[blk + 1]; e += kSchurThreads) schur_pair(d, e, sca, scb, acc, racc);
#pragma unroll
  for (int k = 0; k < 36; k++) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_part[warp][k] = v;
  }
  if (a == b) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double v = warp_sum(racc[k]);
      if (lane == 0) s_part[warp][36 + k] = v;
    }
  }
  __syncthreads();
  if (tid < 36) {
    const int r = tid / 6, c = tid - 6 * r;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kSchurThreads / 32; w++) v += s_part[w][tid];
    v = -v;
    if (a == b && d.is_root) {   // keyframe-only terms are added once (by the root rank when points are sharded)
      const double* Hc = d.Hcc + 21 * (size_t)a;
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      const double h = sca[r] * sca[c] * Hc[SYM6(lo, hi)];
      v += h;
      if (r == c) v += fmin(fmax(h, kMinLmDiag), kMaxLmDiag) / st.radius;
    }
    d.Sblk[(size_t)blk * 36 + tid] = v;   // entry (r, c) of block (a, b)
  } else if (tid < 42 && a == b) {
    const int k = tid - 36;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kSchurThreads / 32; w++) v += s_part[w][36 + k];
    d.rhs[6 * a + k] = (d.is_root ? sca[k] * d.gc[6 * (size_t)a + k] : 0.0) - v;
  }
}

// Chunked Schur complement.  The pair lists of the blocks are very uneven (configs[4]: 500 pairs on a diagonal block, a
// handful on the band's edge; with a warp or a CTA per block the diagonal blocks were a 16-round dependent-load chain that
// set the kernel's duration while most warps had nothing to do).  A WARP takes one CHUNK of <= kSchurChunk pairs of one
// block — one pair per lane per round — and the 36 + 6 partial sums are combined by a REDUCE-SCATTER over the lanes (each
// step a lane keeps one half of its values and receives the partner's partials of that half: 16 + 8 + 4 + 2 + 1 shuffles
// for 32 values instead of 32 x 5), the remaining 4 + 6 values by butterflies.  k_schur_combine then sums the chunks of a
// block in chunk order and adds the keyframe-only terms.  Fixed order everywhere: repeated runs are bit-identical.
constexpr int kSchurWarps = 4;
constexpr int kSchurPart = 42;             // 36 block entries + 6 rhs entries per chunk
#ifndef CMOS_SCHUR_MINB
#define CMOS_SCHUR_MINB 1   // resident CTAs per SM the register allocation must allow (A/B: tools/build_variant.sh)
#endif
__global__ void __launch_bounds__(32 * kSchurWarps, CMOS_SCHUR_MINB) k_schur_chunks(BaDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done) return;
  const int lane = threadIdx.x & 31, chunk = blockIdx.x * kSchurWarps + (threadIdx.x >> 5);
  if (chunk >= d.n_chunks) return;
  const int blk = d.chunk_blk[chunk];
  const int a = d.blk_a[blk], b = d.blk_b[blk];
  double sca[6], scb[6];
#pragma unroll
  for (int k = 0; k < 6; k++) { sca[k] = d.scale_c[6 * (size_t)a + k]; scb[k] = d.scale_c[6 * (size_t)b + k]; }
  double acc[36], racc[6];
#pragma unroll
  for (int k = 0; k < 36; k++) acc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) racc[k] = 0.0;
  for (int e = d.chunk_start[chunk] + lane; e < d.chunk_start[chunk + 1]; e += 32) schur_pair(d, e, sca, scb, acc, racc);
  // reduce-scatter of acc[0..31]: after the step with offset o a lane holds o values; lane L ends with the total of value
  // index L (bit 4 of the lane selects +16, bit 3 +8, ...)
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const double keep = up ? acc[o + i] : acc[i];
      const double send = up ? acc[i] : acc[o + i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  double tail = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double v = warp_sum(acc[32 + k]);
    if (lane == k) tail = v;
  }
  if (a == b) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double v = warp_sum(racc[k]);
      if (lane == 4 + k) tail = v;
    }
  }
  double* out = d.schur_part + (size_t)chunk * kSchurPart;
  out[lane] = acc[0];
  if (lane < 10) out[32 + lane] = tail;
}

// thread per (block, entry): sums the block's chunk partials in chunk order, adds H_cc + D_c^2 / g_c on the diagonal
// entry idx of reduced-system block blk = (a, b): idx < 36 the block's entry (r, c) = (idx / 6, idx % 6), 36..41 the rhs of
// keyframe a (diagonal blocks only): the block's chunks summed in chunk order, plus the keyframe-only terms
__device__ __forceinline__ double schur_combined(const BaDev& d, const LmState& st, const int blk, const int idx, const int a,
                                                 const int b) {
  double sum = 0.0;
  for (int c = d.blk_chunk0[blk]; c < d.blk_chunk0[blk + 1]; c++) sum += d.schur_part[(size_t)c * kSchurPart + idx];
  if (idx < 36) {
    const int r = idx / 6, c = idx - 6 * r;
    double v = -sum;
    if (a == b && d.is_root) {   // keyframe-only terms are added once (by the root rank when points are sharded)
      const double* Hc = d.Hcc + 21 * (size_t)a;
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      const double h = d.scale_c[6 * (size_t)a + r] * d.scale_c[6 * (size_t)a + c] * Hc[SYM6(lo, hi)];
      v += h;
      if (r == c) v += fmin(fmax(h, kMinLmDiag), kMaxLmDiag) / st.radius;
    }
    return v;
  }
  const int k = idx - 36;
  return (d.is_root ? d.scale_c[6 * (size_t)a + k] * d.gc[6 * (size_t)a + k] : 0.0) - sum;
}

__global__ void __launch_bounds__(256) k_schur_combine(BaDev d) { pdl_begin();
  const LmState& st = *d.st;
  if (st.done) return;
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int blk = t / 48, idx = t - 48 * blk;
  if (blk >= d.n_blocks || idx >= kSchurPart) return;
  const int a = d.blk_a[blk], b = d.blk_b[blk];
  if (idx >= 36 && a != b) return;
  const double v = schur_combined(d, st, blk, idx, a, b);
  if (idx < 36) d.Sblk[(size_t)blk * 36 + idx] = v;
  else d.rhs[6 * a + idx - 36] = v;
}

// candidate keyframe poses from the solved reduced system + the keyframes' share of the step statistics
__device__ void cam_candidate(const BaDev& d, const LmState& st, int cam) {
  const int a = d.cam_var[cam];
  const double* c = d.cams[st.cur] + 7 * (size_t)cam;
  double* cc = d.cams[st.cur ^ 1] + 7 * (size_t)cam;
  if (a < 0) {   // constant keyframes are identical in both buffers (set at upload)
    return;
  }
  const double* Hc = d.Hcc + 21 * (size_t)a;
  double delta[6], mcc = 0.0;
  for (int k = 0; k < 6; k++) {
    const double s = d.scale_c[6 * (size_t)a + k];
    const double step = -d.yc[6 * a + k];
    const double D2 = fmin(fmax(s * s * Hc[SYM6(k, k)], kMinLmDiag), kMaxLmDiag) / st.radius;
    mcc += step * (D2 * step - s * d.gc[6 * (size_t)a + k]);
    delta[k] = step * s;
  }
  for (int k = 0; k < 3; k++) cc[k] = c[k] + delta[k];
  quat_plus(c + 3, delta + 3, cc + 3);
  double sn2 = 0.0;
  for (int k = 0; k < 7; k++) sn2 += (c[k] - cc[k]) * (c[k] - cc[k]);
  d.part[d.o_cam_mcc + a] = 0.5 * mcc;
  d.part[d.o_cam_sn2 + a] = sn2;
}

// Cholesky of the reduced camera system in shared memory: packed lower triangle, the rhs carried as row n so
// that the forward substitution rides on the factorisation.  Right-looking in panels of 6 columns (one keyframe
// block): the 6x6 diagonal block is factored by one warp with shuffles, every row below solves its 6 entries
// against it independently, the trailing update is a rank-6 update with one warp per row.  Back substitution by
// one warp, then the candidate keyframe poses.
constexpr int kPB = 6;
// generated straight-line code (tools: see DESIGN.md §4): arrays indexed in nested unrolled loops were kept in local memory by nvcc
__device__ __forceinline__ void factor_diag6(double* __restrict__ L, int k0, int* s_fail) {
  double* const row0 = L + (k0 + 0) * (k0 + 0 + 1) / 2 + k0;
  double* const row1 = L + (k0 + 1) * (k0 + 1 + 1) / 2 + k0;
  double* const row2 = L + (k0 + 2) * (k0 + 2 + 1) / 2 + k0;
  double* const row3 = L + (k0 + 3) * (k0 + 3 + 1) / 2 + k0;
  double* const row4 = L + (k0 + 4) * (k0 + 4 + 1) / 2 + k0;
  double* const row5 = L + (k0 + 5) * (k0 + 5 + 1) / 2 + k0;
  double a00 = row0[0];
  double a10 = row1[0], a11 = row1[1];
  double a20 = row2[0], a21 = row2[1], a22 = row2[2];
  double a30 = row3[0], a31 = row3[1], a32 = row3[2], a33 = row3[3];
  double a40 = row4[0], a41 = row4[1], a42 = row4[2], a43 = row4[3], a44 = row4[4];
  double a50 = row5[0], a51 = row5[1], a52 = row5[2], a53 = row5[3], a54 = row5[4], a55 = row5[5];
  bool ok = true;
  ok = ok && (a00 > 0.0) && isfinite(a00);
  { const double inv = __drcp_rn(a00);
    { const double t = a10 * inv; a11 -= a10 * t; a21 -= a20 * t; a31 -= a30 * t; a41 -= a40 * t; a51 -= a50 * t; }
    { const double t = a20 * inv; a22 -= a20 * t; a32 -= a30 * t; a42 -= a40 * t; a52 -= a50 * t; }
    { const double t = a30 * inv; a33 -= a30 * t; a43 -= a40 * t; a53 -= a50 * t; }
    { const double t = a40 * inv; a44 -= a40 * t; a54 -= a50 * t; }
    { const double t = a50 * inv; a55 -= a50 * t; }
  }
  ok = ok && (a11 > 0.0) && isfinite(a11);
  { const double inv = __drcp_rn(a11);
    { const double t = a21 * inv; a22 -= a21 * t; a32 -= a31 * t; a42 -= a41 * t; a52 -= a51 * t; }
    { const double t = a31 * inv; a33 -= a31 * t; a43 -= a41 * t; a53 -= a51 * t; }
    { const double t = a41 * inv; a44 -= a41 * t; a54 -= a51 * t; }
    { const double t = a51 * inv; a55 -= a51 * t; }
  }
  ok = ok && (a22 > 0.0) && isfinite(a22);
  { const double inv = __drcp_rn(a22);
    { const double t = a32 * inv; a33 -= a32 * t; a43 -= a42 * t; a53 -= a52 * t; }
    { const double t = a42 * inv; a44 -= a42 * t; a54 -= a52 * t; }
    { const double t = a52 * inv; a55 -= a52 * t; }
  }
  ok = ok && (a33 > 0.0) && isfinite(a33);
  { const double inv = __drcp_rn(a33);
    { const double t = a43 * inv; a44 -= a43 * t; a54 -= a53 * t; }
    { const double t = a53 * inv; a55 -= a53 * t; }
  }
  ok = ok && (a44 > 0.0) && isfinite(a44);
  { const double inv = __drcp_rn(a44);
    { const double t = a54 * inv; a55 -= a54 * t; }
  }
  ok = ok && (a55 > 0.0) && isfinite(a55);
  const double rs0 = rsqrt(a00), rs1 = rsqrt(a11), rs2 = rsqrt(a22), rs3 = rsqrt(a33), rs4 = rsqrt(a44), rs5 = rsqrt(a55);
  a10 *= rs0;
  a20 *= rs0; a21 *= rs1;
  a30 *= rs0; a31 *= rs1; a32 *= rs2;
  a40 *= rs0; a41 *= rs1; a42 *= rs2; a43 *= rs3;
  a50 *= rs0; a51 *= rs1; a52 *= rs2; a53 *= rs3; a54 *= rs4;
  { const double d0 = rs0; row0[0] = d0;
    const double d1 = -(a10 * d0) * rs1; row1[0] = d1;
    const double d2 = -(a20 * d0 + a21 * d1) * rs2; row2[0] = d2;
    const double d3 = -(a30 * d0 + a31 * d1 + a32 * d2) * rs3; row3[0] = d3;
    const double d4 = -(a40 * d0 + a41 * d1 + a42 * d2 + a43 * d3) * rs4; row4[0] = d4;
    const double d5 = -(a50 * d0 + a51 * d1 + a52 * d2 + a53 * d3 + a54 * d4) * rs5; row5[0] = d5;
  }
  { const double d1 = rs1; row1[1] = d1;
    const double d2 = -(a21 * d1) * rs2; row2[1] = d2;
    const double d3 = -(a31 * d1 + a32 * d2) * rs3; row3[1] = d3;
    const double d4 = -(a41 * d1 + a42 * d2 + a43 * d3) * rs4; row4[1] = d4;
    const double d5 = -(a51 * d1 + a52 * d2 + a53 * d3 + a54 * d4) * rs5; row5[1] = d5;
  }
  { const double d2 = rs2; row2[2] = d2;
    const double d3 = -(a32 * d2) * rs3; row3[2] = d3;
    const double d4 = -(a42 * d2 + a43 * d3) * rs4; row4[2] = d4;
    const double d5 = -(a52 * d2 + a53 * d3 + a54 * d4) * rs5; row5[2] = d5;
  }
  { const double d3 = rs3; row3[3] = d3;
    const double d4 = -(a43 * d3) * rs4; row4[3] = d4;
    const double d5 = -(a53 * d3 + a54 * d4) * rs5; row5[3] = d5;
  }
  { const double d4 = rs4; row4[4] = d4;
    const double d5 = -(a54 * d4) * rs5; row5[4] = d5;
  }
  { const double d5 = rs5; row5[5] = d5;
  }
  if (!ok) *s_fail = 1;
}

// rank-6 update of kGroups x 32 columns of one row of the packed triangle (lane = column within a group)
template <int kGroups>
__device__ __forceinline__ void solve_row_update(const double* li, const double* __restrict__ P, int ps, double* __restrict__ rowp,
                                                 int cb, int cmax) {
  double v[kGroups];
#pragma unroll
  for (int u = 0; u < kGroups; u++) v[u] = 0.0;
#pragma unroll
  for (int q = 0; q < kPB; q++) {
    const double* pq = P + q * ps;
#pragma unroll
    for (int u = 0; u < kGroups; u++) { const int c = cb + 32 * u; if (c <= cmax) v[u] += li[q] * pq[c]; }
  }
#pragma unroll
  for (int u = 0; u < kGroups; u++) { const int c = cb + 32 * u; if (c <= cmax) rowp[c] -= v[u]; }
}

constexpr int kPBuf = 8;      // rows of the panel buffer: 6 panel columns + 2 rows of zeros (the k = 8 of two m8n8k4 steps)

// fp64 tensor-core tile product: D (8 x 8) += A (8 x 4) B (4 x 8).  Fragment layout (PTX ISA, mma.m8n8k4 .f64):
// A[row = lane / 4][col = lane % 4], B[row = lane % 4][col = lane / 4], C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
__device__ __forceinline__ void dmma_884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#ifndef CMOS_CHOL_V1
// Right-looking Cholesky of a packed lower triangle in shared memory, panels of 6 columns, rows 0..n (row n = the rhs,
// which leaves as the forward-substituted y).  On return every 6x6 diagonal block holds the INVERSE of its Cholesky
// block (factor_diag6).  All kSolveThreads threads of the CTA call it; P is the [kPBuf][ps] panel buffer.
//
// The pivot chain — factor the 6 x 6 diagonal block, solve the six rows of the next block against it, update the next
// diagonal block, factor again — is the critical path of every BA solve (LocalBA: 15 factorisations of 120 unknowns;
// GlobalBA: 6 per LM iteration).  Warp 0 does nothing else: it runs one block ahead of the other 15 warps, which solve the
// remaining rows and apply the rank-6 trailing update on the fp64 tensor cores (8 x 8 tiles, the panel's six columns padded
// to k = 8 with zero rows of P).  One named barrier hands the solved panel from all warps to the updaters without stopping
// warp 0; one __syncthreads per panel.  (Round 1's version: every warp took part in every phase, two __syncthreads per
// panel, scalar trailing update bound by shared-memory wavefronts: 3760 cycles per panel at n = 114, now measured in
// profiles/.)
__device__ __forceinline__ void packed_cholesky(double* __restrict__ L, double* __restrict__ P, const int n, const int ps,
                                                int* s_fail) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kUpd = kSolveThreads - 32, kUpdWarps = kUpd / 32;
  for (int i = tid; i < 2 * ps; i += kSolveThreads) P[6 * ps + i] = 0.0;       // k = 6, 7 of the padded panel
  // (r, c) of the next diagonal block's element a lane of warp 0 updates
  int dr = 0, dc = lane;
  while (dc > dr) { dc -= dr + 1; dr++; }
  if (tid == 0 && n > 0) factor_diag6(L, 0, s_fail);
  __syncthreads();
  const int g = lane >> 2, q = lane & 3;
#ifdef CMOS_CR_TIMING
  long long pk[6] = {0, 0, 0, 0, 0, 0}, tp = clock64();
#define PK(i) { const long long tn = clock64(); pk[i] += tn - tp; tp = tn; }
#else
#define PK(i)
#endif
  for (int k0 = 0; k0 < n; k0 += kPB) {
    if (*s_fail) break;
    const int t0 = k0 + kPB;
    const int nc = t0 < n ? kPB : 0;                       // rows of the next diagonal block (warp 0's)
    const int R0 = t0 + nc;                                // first row of the other warps
    // ---- panel solve: row i times inv(L11)'
    auto solve_row = [&](const int i) {
      double* rowp = L + i * (i + 1) / 2 + k0;
      double a[kPB], x[kPB];
#pragma unroll
      for (int c = 0; c < kPB; c++) a[c] = rowp[c];
#pragma unroll
      for (int c = 0; c < kPB; c++) {
        const double* di = L + (k0 + c) * (k0 + c + 1) / 2 + k0;
        double v = 0.0;
#pragma unroll
        for (int qq = 0; qq <= c; qq++) v += a[qq] * di[qq];
        x[c] = v;
      }
#pragma unroll
      for (int c = 0; c < kPB; c++) { rowp[c] = x[c]; P[c * ps + i] = x[c]; }
    };
    if (warp == 0) {
      if (lane < nc) solve_row(t0 + lane);
      __syncwarp();
      PK(0)
      asm volatile("bar.arrive 1, %0;" ::"n"(kSolveThreads) : "memory");   // hands the chain rows' X to the updaters
      if (nc) {
        if (lane < kPB * (kPB + 1) / 2) {
          const int i = t0 + dr, cc = t0 + dc;
          double v = 0.0;
#pragma unroll
          for (int qq = 0; qq < kPB; qq++) v += P[qq * ps + i] * P[qq * ps + cc];
          L[i * (i + 1) / 2 + cc] -= v;
        }
        __syncwarp();
        PK(1)
        if (lane == 0) factor_diag6(L, t0, s_fail);
        PK(2)
      }
    } else {
      for (int i = R0 + tid - 32; i <= n; i += kUpd) solve_row(i);
      PK(0)
      asm volatile("bar.sync 1, %0;" ::"n"(kSolveThreads) : "memory");
      PK(1)     // every X of this panel is in P
      // ---- trailing update of rows R0..n, columns t0..min(row, n - 1): 8 x 8 tiles, row tile ti has min(ti + 2, max_ct)
      // column tiles (the triangle), flattened over the 15 updater warps
      const int n_rt = (n - R0 + 8) >> 3;                  // ceil((n - R0 + 1) / 8)
      const int max_ct = ((n - 1 - t0) >> 3) + 1;
      const int lin = max(0, min(n_rt, max_ct - 1));       // row tiles that still have ti + 2 column tiles
      const int total = lin * (lin + 3) / 2 + (n_rt - lin) * max_ct;
      for (int e = warp - 1; e < total; e += kUpdWarps) {
        int ti, tj;
        if (e < lin * (lin + 3) / 2) {
          ti = (int)((sqrtf(8.0f * (float)e + 9.0f) - 3.0f) * 0.5f);
          while ((ti + 1) * (ti + 4) / 2 <= e) ti++;
          while (ti * (ti + 3) / 2 > e) ti--;
          tj = e - ti * (ti + 3) / 2;
        } else {
          const int r = e - lin * (lin + 3) / 2;
          ti = lin + r / max_ct; tj = r - (r / max_ct) * max_ct;
        }
        const int row = R0 + 8 * ti + g, col = t0 + 8 * tj + 2 * q;          // C fragment: (row, col), (row, col + 1)
        const int arow = min(row, n), bcol = min(t0 + 8 * tj + g, n);         // clamped operand rows (masked at the store)
        const double a0 = -P[q * ps + arow], a1 = -P[(4 + q) * ps + arow];
        const double b0 = P[q * ps + bcol], b1 = P[(4 + q) * ps + bcol];
        const int cmax = min(row, n - 1);
        const bool v0 = row <= n && col <= cmax, v1 = row <= n && col + 1 <= cmax;
        double* cp = L + row * (row + 1) / 2 + col;
        double c0 = v0 ? cp[0] : 0.0, c1 = v1 ? cp[1] : 0.0;
        dmma_884(c0, c1, a0, b0);
        dmma_884(c0, c1, a1, b1);
        if (v0) cp[0] = c0;
        if (v1) cp[1] = c1;
      }
      PK(2)
    }
    __syncthreads();
    PK(3)
  }
#ifdef CMOS_CR_TIMING
  if ((tid == 0 || tid == 32 || tid == 480) && blockIdx.x == 0)
    printf("packed_cholesky v2 tid %d n %d: [w0: solve6 / diag-update / factor_diag6 | updaters: solve / barrier / dmma update] %lld %lld %lld | end sync %lld\n", tid, n, pk[0], pk[1], pk[2], pk[3]);
#endif
#undef PK
}
#else
// Right-looking Cholesky of a packed lower triangle in shared memory, panels of 6 columns, rows 0..n (row n = the rhs,
// which leaves as the forward-substituted y).  On return every 6x6 diagonal block holds the INVERSE of its Cholesky
// block (factor_diag6).  All kSolveThreads threads of the CTA call it; P is the [kPBuf][ps] panel buffer.
__device__ __forceinline__ void packed_cholesky_v1(double* __restrict__ L, double* __restrict__ P, const int n, const int ps,
                                                int* s_fail) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kSolveThreads / 32;
#ifdef CMOS_CR_TIMING
  long long pk[6] = {0, 0, 0, 0, 0, 0}, tp = clock64();
#define PK(i) { const long long tn = clock64(); pk[i] += tn - tp; tp = tn; }
#else
#define PK(i)
#endif
  if (tid == 0 && n > 0) factor_diag6(L, 0, s_fail);
  __syncthreads();
  for (int k0 = 0; k0 < n; k0 += kPB) {
    if (*s_fail) break;
    const int t0 = k0 + kPB;
    // panel solve: row i of the panel times inv(L11)'
    for (int i = t0 + tid; i <= n; i += kSolveThreads) {
      double* rowp = L + i * (i + 1) / 2 + k0;
      double a[kPB], x[kPB];
#pragma unroll
      for (int c = 0; c < kPB; c++) a[c] = rowp[c];
#pragma unroll
      for (int c = 0; c < kPB; c++) {
        const double* di = L + (k0 + c) * (k0 + c + 1) / 2 + k0;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q <= c; q++) v += a[q] * di[q];
        x[c] = v;
      }
#pragma unroll
      for (int c = 0; c < kPB; c++) { rowp[c] = x[c]; P[c * ps + i] = x[c]; }
    }
    PK(0)
    __syncthreads();
    PK(1)
    // trailing update.  Look-ahead: warp 0 updates the NEXT diagonal block (rows t0..t0+5 lie entirely inside it) and its
    // lane 0 factors it at once, while the other warps update the rows below — the next round starts with its panel solve.
    if (warp == 0) {
      if (t0 < n) {
        if (lane < kPB * (kPB + 1) / 2) {
          int r = 0, c = lane;
          while (c > r) { c -= r + 1; r++; }
          const int i = t0 + r, cc = t0 + c;
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < kPB; q++) v += P[q * ps + i] * P[q * ps + cc];
          L[i * (i + 1) / 2 + cc] -= v;
        }
        __syncwarp();
        PK(2)
        if (lane == 0) factor_diag6(L, t0, s_fail);
        PK(3)
      }
    } else {
#if CMOS_SOLVE_TILE
      // Register-tiled rank-6 update: a warp takes 4 consecutive rows, its lanes 4 column groups (c, c+32, c+64, c+96) —
      // 16 accumulators per thread.  The update is bound by shared-memory wavefronts, not by fp64 (one SM, 64 fp64 lanes);
      // the tile reads each panel value once per 4 rows / 4 columns instead of once per FMA.
      const int first = t0 + kPB;
      for (int i0 = first + 4 * (warp - 1); i0 <= n; i0 += 4 * (nw - 1)) {
        const int top = min(i0 + 3, n);                       // last row of the tile
        const int cend = top < n ? top : n - 1;               // widest row's last column
        for (int cb = t0 + lane; cb <= cend; cb += 128) {
          const int nu = min(4, (cend - (cb - lane)) / 32 + 1);
          double v[4][4];
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int u = 0; u < 4; u++) v[r][u] = 0.0;
#pragma unroll
          for (int q = 0; q < kPB; q++) {
            const double* pq = P + q * ps;
            double li[4], pc[4];
#pragma unroll
            for (int r = 0; r < 4; r++) li[r] = pq[min(i0 + r, n)];
#pragma unroll
            for (int u = 0; u < 4; u++)
              if (!CMOS_SOLVE_NU || u < nu) {                                     // warp-uniform: column groups past the tile's widest row
                pc[u] = pq[min(cb + 32 * u, n)];
#pragma unroll
                for (int r = 0; r < 4; r++) v[r][u] += li[r] * pc[u];
              }
          }
#pragma unroll
          for (int r = 0; r < 4; r++) {
            const int i = i0 + r;
            if (i > n) break;
            const int cmax = i < n ? i : n - 1;
            double* rowp = L + i * (i + 1) / 2;
#pragma unroll
            for (int u = 0; u < 4; u++) { const int c = cb + 32 * u; if (c <= cmax) rowp[c] -= v[r][u]; }
          }
        }
      }
    }
#else
      for (int i = t0 + kPB + warp - 1; i <= n; i += nw - 1) {
        double li[kPB];
#pragma unroll
        for (int q = 0; q < kPB; q++) li[q] = P[q * ps + i];
        const int cmax = i < n ? i : n - 1;
        double* rowp = L + i * (i + 1) / 2;
        // column groups of 32 the row really has (warp-uniform): short rows do not issue the other groups' instructions
        for (int c0 = t0; c0 <= cmax; c0 += 128) {
          const int ng = min(4, (cmax - c0) / 32 + 1);
          if (ng == 4) solve_row_update<4>(li, P, ps, rowp, c0 + lane, cmax);
          else if (ng == 3) solve_row_update<3>(li, P, ps, rowp, c0 + lane, cmax);
          else if (ng == 2) solve_row_update<2>(li, P, ps, rowp, c0 + lane, cmax);
          else solve_row_update<1>(li, P, ps, rowp, c0 + lane, cmax);
        }
      }
    }
#endif
    PK(4)
    __syncthreads();
    PK(5)
  }
#ifdef CMOS_CR_TIMING
  if ((tid == 0 || tid == 32) && blockIdx.x == 0)
    printf("packed_cholesky tid %d n %d: solve %lld sync %lld | w0 diag-update %lld factor_diag6 %lld | trailing %lld sync %lld\n", tid, n, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5]);
#endif
#undef PK
}

#define packed_cholesky packed_cholesky_v1
#endif

// After packed_cholesky: invert the 24 x 24 diagonal Cholesky blocks in place ([[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]],
// 6 -> 12 -> 24; the 6 x 6 diagonal blocks are already stored inverted; the last block may be 6, 12 or 18 wide).
// Substitutions then run in 24-row steps of independent dot products instead of 6-row steps.  T: scratch of 6 n + 36 doubles
// (the panel buffer).  All kSolveThreads threads.
__device__ __forceinline__ void invert_diag24(double* __restrict__ L, double* __restrict__ T, const int n) {
  const int tid = threadIdx.x, nb = n / 6;
  for (int sb = 1; sb <= 2; sb *= 2) {
    const int np = (nb + sb - 1) / (2 * sb), pe = 36 * sb * sb, sa = 6 * sb;
    for (int e = tid; e < np * pe; e += kSolveThreads) {         // T = L_CA * M_AA
      const int pr = e / pe, rem = e - pr * pe, r = rem / sa, j = rem - r * sa;
      const int a0 = 6 * (2 * pr * sb), c0 = a0 + sa;
      if (c0 + r >= min(c0 + sa, n)) continue;
      const double* Lrow = L + (c0 + r) * (c0 + r + 1) / 2 + a0;
      double v = 0.0;
      for (int k = j; k < sa; k++) v += Lrow[k] * L[(a0 + k) * (a0 + k + 1) / 2 + a0 + j];
      T[e] = v;
    }
    __syncthreads();
    for (int e = tid; e < np * pe; e += kSolveThreads) {         // M_CA = -M_CC * T, over L_CA
      const int pr = e / pe, rem = e - pr * pe, r = rem / sa, j = rem - r * sa;
      const int a0 = 6 * (2 * pr * sb), c0 = a0 + sa;
      if (c0 + r >= min(c0 + sa, n)) continue;
      const double* Mrow = L + (c0 + r) * (c0 + r + 1) / 2 + c0;
      const double* Tc = T + (size_t)pr * pe + j;
      double v = 0.0;
      for (int k = 0; k <= r; k++) v += Mrow[k] * Tc[k * sa];
      L[(c0 + r) * (c0 + r + 1) / 2 + a0 + j] = -v;
    }
    __syncthreads();
  }
}

// L' x = y in 24-row block steps against a factor whose 24 x 24 diagonal blocks are inverted (invert_diag24), from the last
// block up: x_p = inv(L_pp)' y_p, then y_q -= L_pq' x_p for the rows above.  y (n entries, shared memory) becomes x.
// s_x: 24 doubles of shared scratch.  All threads of the CTA (blockDim.x >= n, a multiple of 32).  Every dot product is
// split over 4 (or 2) adjacent lanes when the CTA has the threads for it — the 24 dependent load + FMA pairs of one thread
// were the whole cost of a step (8.4 k cycles at n = 120) — and added by shuffles in a fixed order.
// (kBlk = 8: the same against the 8 x 8 inverted diagonal tiles packed_cholesky_reg leaves — no 24-block inversion needed.)
template <int kBlk>
__device__ __forceinline__ void back_substitute_blocks(const double* __restrict__ L, double* __restrict__ y, double* __restrict__ s_x,
                                                       const int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lpu = 4 * n <= nt ? 4 : (2 * n <= nt ? 2 : 1);      // lanes per unknown
  const int u = tid / lpu, sub = tid - u * lpu;
  for (int r0 = ((n - 1) / kBlk) * kBlk; r0 >= 0; r0 -= kBlk) {
    const int bs = min(kBlk, n - r0);
    {
      double v = 0.0;
      if (u < bs)
        for (int qq = u + sub; qq < bs; qq += lpu) v += L[(r0 + qq) * (r0 + qq + 1) / 2 + r0 + u] * y[r0 + qq];
      if (lpu == 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (lpu >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (u < bs && sub == 0) s_x[u] = v;
    }
    __syncthreads();
    {
      double v = 0.0;
      if (u < r0)
        for (int qq = sub; qq < bs; qq += lpu) v += L[(r0 + qq) * (r0 + qq + 1) / 2 + u] * s_x[qq];
      if (lpu == 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (lpu >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (sub == 0) {
        if (u < r0) y[u] -= v;
        else if (u < r0 + bs) y[u] = s_x[u - r0];
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void back_substitute24(const double* __restrict__ L, double* __restrict__ y, double* __restrict__ s_x,
                                                  const int n) {
  back_substitute_blocks<24>(L, y, s_x, n);
}

}  // namespace cmos
#include "chol_reg.cuh"
namespace cmos {

// Doubles of scratch behind the packed triangle that factor_and_invert24 needs for an n-column system.
__host__ __device__ inline size_t chol_scratch_doubles(int n) {
  const size_t v2 = (size_t)kPBuf * (n + 2) + 40;
#ifndef CMOS_CHOL_SMEM
  if (chol_reg_supported(n)) { const size_t r = chol_reg_scratch(n); return r > v2 ? r : v2; }
#endif
  return v2;
}

// Cholesky of the packed triangle L (rows 0..n, row n = rhs -> y) followed by the inversion of the 24 x 24 diagonal blocks
// (the format invert_diag24 documents).  Systems of up to 120 unknowns take the register-resident tile factorisation
// (chol_reg.cuh), larger ones the shared-memory panels of packed_cholesky; CMOS_CHOL_SMEM at build time forces the latter
// (A/B).  P: chol_scratch_doubles(n) doubles.  All kSolveThreads threads; L complete (synchronised) on entry; on return
// *s_fail tells whether a pivot failed (then the diagonal blocks are not inverted).
__device__ __forceinline__ void factor_and_invert24(double* __restrict__ L, double* __restrict__ P, const int n, int* s_fail) {
#ifndef CMOS_CHOL_SMEM
  if (chol_reg_supported(n)) {
    packed_cholesky_reg(L, P, n, s_fail);
    __syncthreads();
#ifdef CMOS_CHOL_TIMING
    const long long ti0 = clock64();
#endif
    if (!*s_fail) invert_diag24_r8(L, P, n);
#ifdef CMOS_CHOL_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("invert_diag24_r8 n %d: %lld cycles\n", n, clock64() - ti0);
#endif
    return;
  }
#endif
  packed_cholesky(L, P, n, n + 2, s_fail);
  __syncthreads();
  if (!*s_fail) invert_diag24(L, P, n);
}

// The whole solve L L' x = rhs (k_solve_small, the border system of the bordered nested dissection).  y = row n of L becomes
// x.  s_x: 24 doubles of shared scratch.  Substituting against the 8 x 8 inverted tiles directly (back_substitute_blocks<8>,
// no 24-block inversion) was measured and is slower: 15 steps of two barrier-separated phases cost 16.7 k cycles at n = 120,
// the 24-block inversion + 5 steps 12 k.
__device__ __forceinline__ void factor_and_solve(double* __restrict__ L, double* __restrict__ P, const int n, int* s_fail,
                                                 double* __restrict__ s_x) {
  factor_and_invert24(L, P, n, s_fail);
  __syncthreads();
  if (!*s_fail) back_substitute24(L, L + n * (n + 1) / 2, s_x, n);
}

__global__ void __launch_bounds__(kSolveThreads) k_solve_small(BaDev d) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  LmState& st = *d.st;
  if (st.done) return;
  const int n = d.nc, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* L = smem_d;                                   // rows 0..n, row i at i*(i+1)/2
  double* P = L + (size_t)(n + 1) * (n + 2) / 2;        // scratch of the factorisation (chol_scratch_doubles)
  __shared__ int s_fail;
  if (tid == 0) s_fail = st.solve_failed;
  for (int i = tid; i < n * (n + 1) / 2; i += kSolveThreads) L[i] = 0.0;
  for (int k = tid; k < n; k += kSolveThreads) L[n * (n + 1) / 2 + k] = d.rhs[k];
  __syncthreads();
  for (int e = tid; e < d.n_blocks * 36; e += kSolveThreads) {
    const int blk = e / 36, rc = e - 36 * blk, r = rc / 6, c = rc - 6 * r;
    const int row = 6 * d.blk_b[blk] + c, col = 6 * d.blk_a[blk] + r;   // lower-triangle copy of block (a,b), a <= b
    if (col <= row) L[row * (row + 1) / 2 + col] = d.Sblk[e];
  }
  __syncthreads();
  // 6x6 diagonal block in registers of ONE thread (fully unrolled).  The pivot chain is the critical path of the whole
  // solve, so it carries one reciprocal per pivot and nothing else: the elimination runs on the unscaled columns
  // (A[r][c] -= A[r][j] A[c][j] / A[j][j]), the six rsqrt that turn them into Cholesky columns are independent and issued
  // together at the end, and what is stored in the block's place is the INVERSE of its Cholesky factor: the panel solve
  // and the back substitution then are short independent dot products instead of dependent substitution chains.
#ifdef CMOS_SOLVE_TIMING
  long long tk[8] = {0,0,0,0,0,0,0,0}; long long tprev = clock64(); const long long tstart = tprev;
#define TK(i) { const long long tn = clock64(); tk[i] += tn - tprev; tprev = tn; }
#else
#define TK(i)
#endif
  __shared__ double s_x24[24];
  factor_and_solve(L, P, n, &s_fail, s_x24);
  TK(4)
  __syncthreads();
  {
    double* y = L + n * (n + 1) / 2;      // forward-substituted rhs, now the solution
    if (warp == 0) {
      int bad = 0;
      for (int k = lane; k < n; k += 32) {
        const double v = s_fail ? 0.0 : y[k];
        d.yc[k] = v;
        bad |= !isfinite(v);
      }
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0 && bad) s_fail = 1;
    }
  }
  __syncthreads();
  TK(5)
  if (tid == 0) st.solve_failed = s_fail;
  if (!s_fail)
    for (int cam = tid; cam < d.K; cam += kSolveThreads) cam_candidate(d, st, cam);
  TK(6)
#ifdef CMOS_SOLVE_TIMING
  if ((tid == 0 || tid == 32 || tid == 1023) && st.iteration == 1)
    printf("solve_small tid %d n %d: setup+diag0 %lld | P2 %lld bar %lld P3 %lld bar %lld | backsub %lld cand %lld | total %lld\n", tid, n, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5], tk[6], clock64() - tstart);
#endif
}

// Test tap of the small dense solver (cmos_debug_solve_spd): A x = b through exactly the device code k_solve_small and
// k_cr_factor run — factor_and_invert24 + back_substitute24 — on a caller-supplied SPD matrix.
__global__ void __launch_bounds__(kSolveThreads) k_debug_solve_spd(const double* __restrict__ A, const double* __restrict__ b, int n,
                                                                   double* __restrict__ x, int* fail, long long* cycles, int whole) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  const int tid = threadIdx.x;
  double* L = smem_d;
  double* P = L + (size_t)(n + 1) * (n + 2) / 2;
  __shared__ int s_fail;
  __shared__ double s_x24[24];
  if (tid == 0) s_fail = 0;
  for (int e = tid; e < n * n; e += kSolveThreads) {
    const int r = e / n, c = e - r * n;
    if (c <= r) L[r * (r + 1) / 2 + c] = A[e];
  }
  for (int k = tid; k < n; k += kSolveThreads) L[n * (n + 1) / 2 + k] = b[k];
  __syncthreads();
  const long long t0 = clock64();
  double* y = L + n * (n + 1) / 2;
  long long t1;
  if (whole) {                           // the path of k_solve_small / k_crb_solve
    factor_and_solve(L, P, n, &s_fail, s_x24);
    t1 = clock64();
  } else {                               // the path of k_cr_factor + k_cr_back (24-block format)
    factor_and_invert24(L, P, n, &s_fail);
    __syncthreads();
    t1 = clock64();
    if (!s_fail) back_substitute24(L, y, s_x24, n);
  }
  __syncthreads();
  const long long t2 = clock64();
  for (int k = tid; k < n; k += kSolveThreads) x[k] = s_fail ? 0.0 : y[k];
  if (tid == 0) { *fail = s_fail; cycles[0] = t1 - t0; cycles[1] = t2 - t1; }
}

__global__ void __launch_bounds__(kLinThreads) k_backsub(BaDev d) { pdl_begin();
  __shared__ double scratch[33];
  const LmState& st = *d.st;
  if (st.done) return;
  const int gt = blockIdx.x * kLinThreads + threadIdx.x;
  const int j = gt / kPointLanes, q = gt % kPointLanes;
  double cost = 0.0, mcc = 0.0, sn2 = 0.0;
  const bool live = j < d.M && !st.solve_failed;       // solve_failed is uniform over the launch
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  if (live) {
    for (int p = d.pt_start[j] + q; p < d.pt_start[j + 1]; p += kPointLanes) {
      const int a = d.o_cv[p];
      if (a < 0) continue;
      double Jc[12], Jp[6];
      load_jc_scaled(d, p, d.scale_c + 6 * (size_t)a, Jc);
      load_jp_scaled(d, p, j, Jp);
      const double* y = d.yc + 6 * a;
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) { v0 += Jc[k] * y[k]; v1 += Jc[6 + k] * y[k]; }
      acc0 += Jp[0] * v0 + Jp[3] * v1; acc1 += Jp[1] * v0 + Jp[4] * v1; acc2 += Jp[2] * v0 + Jp[5] * v1;
    }
  }
#pragma unroll
  for (int o = 1; o < kPointLanes; o <<= 1) {
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
  }
  if (live) {
    const double* cams_c = d.cams[st.cur ^ 1];
    const double* X = d.pts[st.cur] + 3 * (size_t)j;
    const double* Hi = d.Hinv + 6 * (size_t)j;
    const double* t = d.tp + 3 * (size_t)j;
    // y_p = Hpp^-1 (g_p - W' y_c) ; step = -y_p
    const double step[3] = {-(t[0] - (Hi[0] * acc0 + Hi[1] * acc1 + Hi[2] * acc2)),
                            -(t[1] - (Hi[1] * acc0 + Hi[3] * acc1 + Hi[4] * acc2)),
                            -(t[2] - (Hi[2] * acc0 + Hi[4] * acc1 + Hi[5] * acc2))};
    const double* H = d.Hpp + 6 * (size_t)j;
    const double hd[3] = {H[0], H[3], H[5]};
    double Xc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double s = d.scale_p[3 * (size_t)j + k];
      const double D2 = fmin(fmax(s * s * hd[k], kMinLmDiag), kMaxLmDiag) / st.radius;
      mcc += step[k] * (D2 * step[k] - s * d.gp[3 * (size_t)j + k]);
      const double delta = step[k] * s;
      Xc[k] = X[k] + delta;
      sn2 += (X[k] - Xc[k]) * (X[k] - Xc[k]);
    }
    mcc *= 0.5;
    if (q == 0) {
      double* Xo = d.pts[st.cur ^ 1] + 3 * (size_t)j;
      Xo[0] = Xc[0]; Xo[1] = Xc[1]; Xo[2] = Xc[2];
    } else { mcc = 0.0; sn2 = 0.0; }               // the point's model change and step norm are counted once
    for (int p = d.pt_start[j] + q; p < d.pt_start[j + 1]; p += kPointLanes) {
      const float2 uv = d.o_uv[p];
      const Proj P = project_obs(cams_c + 7 * (size_t)d.o_cam[p], Xc, d.fx, d.fy, d.cx, d.cy, (double)uv.x, (double)uv.y,
                                 (double)d.o_w[p]);
      double w;
      cost += loss_eval(P.r0 * P.r0 + P.r1 * P.r1, d.o_mode[p], &w);
    }
  }
  cost = block_sum(cost, scratch);
  mcc = block_sum(mcc, scratch);
  sn2 = block_sum(sn2, scratch);
  if (threadIdx.x == 0) {
    d.part[d.o_bs_cost + blockIdx.x] = cost;
    d.part[d.o_bs_mcc + blockIdx.x] = mcc;
    d.part[d.o_bs_sn2 + blockIdx.x] = sn2;
  }
  if (d.multi) return;
  if (last_cta_arrives(d.ticket + 1, gridDim.x)) decide_body(d, *d.st, 0, scratch);
}

// phase 1: red[3..5] = candidate cost, model cost change, |step|^2 of this rank's points; phase 2: add the
// keyframes' share and decide; phase 0: both
__device__ void decide_body(const BaDev& d, LmState& st, int phase, double* scratch) {
  if (phase != 2) {
    const double cost = part_sum(d.part + d.o_bs_cost, d.n_lin_blocks, scratch);
    const double mcc = part_sum(d.part + d.o_bs_mcc, d.n_lin_blocks, scratch);
    const double sn2 = part_sum(d.part + d.o_bs_sn2, d.n_lin_blocks, scratch);
    if (threadIdx.x == 0) { d.red[3] = cost; d.red[4] = mcc; d.red[5] = sn2; d.red[6] = st.solve_failed ? 1.0 : 0.0; }
    __syncthreads();
  }
  if (phase != 1) {
    const bool ok = !(d.red[6] > 0.0);
    double mcc = 0, sn2 = 0;
    if (ok) {
      mcc = d.red[4] + part_sum(d.part + d.o_cam_mcc, d.Kv, scratch);
      sn2 = d.red[5] + part_sum(d.part + d.o_cam_sn2, d.Kv, scratch);
    }
    if (threadIdx.x == 0) {
      lm_decide(st, ok, mcc, d.red[3], sqrt(sn2), d.trace);
      st.solve_failed = 0;
      dbg_raise_stop(d, st);
      if (!st.done && d.stop_flag && *d.stop_flag) { st.done = 1; st.termination = TERM_USER; }
    }
  }
}
__global__ void __launch_bounds__(256) k_decide(BaDev d, int phase) { pdl_begin();
  __shared__ double scratch[33];
  LmState& st = *d.st;
  if (st.done) return;
  decide_body(d, st, phase, scratch);
}

// CheckOutlier (:227-241) + the z <= 0 test (:558-564) over the observations of local keyframes; also sets the
// per-observation loss mode of the next pass (quirk Q2) and copies the pass summary out.
__global__ void __launch_bounds__(256) k_outlier_scan(BaDev d, const uint8_t* __restrict__ cam_flags,
                                                      const int* __restrict__ perm, uint8_t* __restrict__ erase,
                                                      int set_mode) { pdl_begin();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= d.N) return;
  const int cur = d.st->cur;
  const int cam = d.o_cam[p];
  uint8_t e = 0;
  if (!(cam_flags[cam] & 2) && !d.st->pad) {
    const float2 uv = d.o_uv[p];
    const Proj P = project_obs(d.cams[cur] + 7 * (size_t)cam, d.pts[cur] + 3 * (size_t)d.o_pt[p], d.fx, d.fy, d.cx, d.cy,
                               (double)uv.x, (double)uv.y, 1.0);
    const double chi2 = (P.r0 * P.r0 + P.r1 * P.r1) * (double)d.o_w[p];
    e = (chi2 > kChi2) || (P.p2 <= 0.0);
  }
  erase[perm[p]] = e;
  if (set_mode) d.o_mode[p] = e ? 1 : 3;
}

__global__ void k_set_mode(BaDev d, int mode) { pdl_begin();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p < d.N) d.o_mode[p] = (uint8_t)mode;
}

// StopFlagCallback returns SOLVER_TERMINATE_SUCCESSFULLY (CeresOptimizer.h:340): ceres::Solve ends with USER_SUCCESS, a usable
// solution, so the parameter blocks keep the iterate reached when the flag was seen — nothing is rolled back.  What discards
// work is the reference's own `if (*stop_flag) return;` at the start of a LocalBundleAdjustment pass (`pad` above).
__global__ void k_summary(BaDev d, cmos_ba_summary* out) { pdl_begin();
  const LmState& st = *d.st;
  cmos_ba_summary s;
  s.iterations = st.iteration; s.successful_steps = st.successful; s.termination = st.termination;
  s.jacobian_evaluations = st.jac_evals; s.initial_cost = st.initial_cost; s.final_cost = st.x_cost;
  *out = s;
}

// final parameters always end up in buffer 0 of the handle's "result" view
__global__ void k_gather_result(BaDev d, const double* cams0, const double* pts0, double* cams_out, double* pts_out) { pdl_begin();
  const int cur = d.st->cur, aborted = d.st->pad;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < 7 * d.K) cams_out[i] = aborted ? cams0[i] : d.cams[cur][i];
  if (i < 3 * d.M) pts_out[i] = aborted ? pts0[i] : d.pts[cur][i];
}

__global__ void __launch_bounds__(256) k_scatter_S(BaDev d) { pdl_begin();
  if (d.st->done) return;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= d.n_blocks * 36) return;
  const int blk = e / 36, rc = e - 36 * blk, r = rc / 6, c = rc - 6 * r;
  const int row = 6 * d.blk_b[blk] + c, col = 6 * d.blk_a[blk] + r;
  if (col <= row) d.S[(size_t)row * d.nc + col] = d.Sblk[e];
}

// -------------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky in HBM for reduced systems that do not fit one CTA's shared memory
// (GlobalBundleAdjustemnt: 6000 x 6000 at 1000 keyframes).  Row-major lower triangle, panel width kNB.
// The rhs is carried as row n (so the forward substitution is part of the factorisation); the inverses of
// the diagonal panels are kept so that the panel solve and the back substitution are plain products.
// -------------------------------------------------------------------------------------------------
// factor the kb x kb diagonal block at (k0,k0) in place and store inv(L11) (lower) into Linv[panel].
// Right-looking on the UNSCALED columns — A[i][k] -= A[i][j] A[k][j] / A[j][j] needs one barrier per column and no
// square root on the critical path; column j is scaled by rsqrt(A[j][j]) once, at the end.
constexpr int kPotrfThreads = 1024;
__global__ void __launch_bounds__(kPotrfThreads) k_potrf_diag(BaDev d, int k0, int kb, double* Linv) { pdl_begin();
  extern __shared__ double dyn_smem[];
  double (*A)[kNB + 1] = (double (*)[kNB + 1])dyn_smem;
  double (*Li)[kNB + 1] = (double (*)[kNB + 1])(dyn_smem + kNB * (kNB + 1));
  __shared__ double s_inv[kNB];                 // 1 / pivot (unscaled columns), then 1 / L[i][i]
  __shared__ int s_bad;
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int n = d.nc, tid = threadIdx.x;
#ifdef CMOS_SOLVE_TIMING
  long long tq[6]; tq[0] = clock64();
#define TQ(i) tq[i] = clock64();
#else
#define TQ(i)
#endif
  for (int idx = tid; idx < kNB * kNB; idx += kPotrfThreads) {
    const int i = idx / kNB, k = idx - i * kNB;
    A[i][k] = (k <= i && i < kb) ? d.S[(size_t)(k0 + i) * n + k0 + k] : 0.0;
    Li[i][k] = 0.0;
  }
  if (tid == 0) s_bad = 0;
  __syncthreads();
  if (tid == 0) {
    const double d0 = A[0][0];
    if (!(d0 > 0.0) || !isfinite(d0)) s_bad = 1;
    s_inv[0] = __drcp_rn(d0);
  }
  __syncthreads();
  TQ(1)
  // Right-looking on the unscaled columns, 32 x 32 threads over the trailing block (rows by ty, columns by tx).  The thread
  // that owns the next pivot updates it first and takes its reciprocal at once, so the reciprocal is off the other
  // threads' path: one barrier per column.
  const int tx = tid & 31, ty = tid >> 5;
  for (int j = 0; j < kb; j++) {
    if (s_bad) break;                            // uniform: written before the last barrier
    const double inv = s_inv[j];
    for (int i = j + 1 + ty; i < kb; i += 32) {
      const double aij = A[i][j] * inv;
      for (int k = j + 1 + tx; k <= i; k += 32) {
        const double v = A[i][k] - aij * A[k][j];
        A[i][k] = v;
        if (i == j + 1 && k == j + 1) {          // the next pivot is final now
          if (!(v > 0.0) || !isfinite(v)) s_bad = 1;
          s_inv[j + 1] = __drcp_rn(v);
        }
      }
    }
    __syncthreads();
  }
  TQ(2)
  if (s_bad) { if (tid == 0) st.solve_failed = 1; return; }
  __syncthreads();
  if (tid < kb) s_inv[tid] = rsqrt(A[tid][tid]);        // 1 / L[j][j]
  __syncthreads();
  for (int idx = tid; idx < kb * kb; idx += kPotrfThreads) {     // L[i][j] = A[i][j] / sqrt(A[j][j])
    const int i = idx / kb, j = idx - i * kb;
    if (j < i) A[i][j] *= s_inv[j];
  }
  __syncthreads();
  if (tid < kb) A[tid][tid] = A[tid][tid] * s_inv[tid];
  __syncthreads();
  TQ(3)
  // inverse of the lower-triangular factor by forward substitution: 4 lanes per column split every row's dot product,
  // two shuffles add the parts in a fixed order; the 8 columns of a warp advance row by row together
  if (tid < 4 * kNB) {
    const int c = tid >> 2, part = tid & 3;
    for (int i = 0; i < kb; i++) {
      double s = 0.0;
      if (c < kb && i > c)
        for (int k = c + part; k < i; k += 4) s += A[i][k] * Li[k][c];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && c < kb && i >= c) Li[i][c] = ((i == c ? 1.0 : 0.0) - s) * s_inv[i];
      __syncwarp();
    }
  }
  __syncthreads();
  TQ(4)
  for (int idx = tid; idx < kb * kb; idx += kPotrfThreads) {
    const int i = idx / kb, k = idx - i * kb;
    if (k <= i) d.S[(size_t)(k0 + i) * n + k0 + k] = A[i][k];
    Linv[(size_t)(k0 / kNB) * kNB * kNB + i * kNB + k] = Li[i][k];
  }
  TQ(5)
#ifdef CMOS_SOLVE_TIMING
  if (tid == 0 && k0 == 640 && st.iteration == 1) printf("potrf: load %lld factor %lld scale %lld inverse %lld store %lld\n", tq[1]-tq[0], tq[2]-tq[1], tq[3]-tq[2], tq[4]-tq[3], tq[5]-tq[4]);
#endif
}

// panel solve: rows i > k0+kb (and the rhs row): L21[i][:] = A21[i][:] * inv(L11)'  -> L21[i][c] = sum_k A21[i][k] Linv[c][k]
// Only the ACTIVE 64-row tiles of the panel are touched: tiles (relative to the first row below the panel) whose
// rows reach into or before the panel's columns — the row envelope of the reduced system, inside which all fill
// of the factorisation stays.  For a windowed co-visibility graph that is a handful of tiles per panel.
__global__ void __launch_bounds__(256) k_trsm_panel(BaDev d, int k0, int kb, const double* __restrict__ Linv,
                                                    const int* __restrict__ tiles) { pdl_begin();
  extern __shared__ double dyn_smem[];
  double (*Li)[kNB + 1] = (double (*)[kNB + 1])dyn_smem;
  double (*Arow)[kNB + 1] = (double (*)[kNB + 1])(dyn_smem + kNB * (kNB + 1));
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int n = d.nc, tid = threadIdx.x;
  const double* Lp = Linv + (size_t)(k0 / kNB) * kNB * kNB;
  for (int idx = tid; idx < kNB * kNB; idx += 256) Li[idx / kNB][idx % kNB] = Lp[idx];
  const int row0 = k0 + kb + tiles[blockIdx.x >> 1] * 64 + (blockIdx.x & 1) * 32;   // rows row0..row0+31; logical row n = rhs
  for (int idx = tid; idx < 32 * kb; idx += 256) {
    const int r = idx / kb, k = idx - r * kb;
    const int i = row0 + r;
    double v = 0.0;
    if (i < n) v = d.S[(size_t)i * n + k0 + k];
    else if (i == n) v = d.rhs[k0 + k];
    Arow[r][k] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * kb; idx += 256) {
    const int r = idx / kb, c = idx - r * kb;
    const int i = row0 + r;
    if (i > n) continue;
    double s = 0.0;
    for (int k = 0; k <= c; k++) s += Arow[r][k] * Li[c][k];
    if (i < n) d.S[(size_t)i * n + k0 + c] = s; else d.rhs[k0 + c] = s;
  }
}

// trailing update: A22 -= L21 L21' on 64x64 tiles of the lower triangle (and the rhs row), 256 threads, 4x4 per thread
__global__ void __launch_bounds__(256) k_syrk_tile(BaDev d, int k0, int kb, const int* __restrict__ tiles) { pdl_begin();
  extern __shared__ double dyn_smem[];
  double (*As)[64 + 1] = (double (*)[64 + 1])dyn_smem;                      // [k][i]
  double (*Bs)[64 + 1] = (double (*)[64 + 1])(dyn_smem + kNB * (64 + 1));   // [k][j]
  const LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int n = d.nc, tid = threadIdx.x;
  const int t0 = k0 + kb;
  // tile coordinates in the trailing matrix from the linear block index (lower triangle incl. diagonal)
  int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) bi++;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) bi--;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  const int i0 = t0 + tiles[bi] * 64, j0 = t0 + tiles[bj] * 64;     // bi, bj index the panel's active-tile list
  for (int idx = tid; idx < 64 * kb; idx += 256) {
    const int r = idx / kb, k = idx - r * kb;
    const int i = i0 + r, j = j0 + r;
    double va = 0.0, vb = 0.0;
    if (i < n) va = d.S[(size_t)i * n + k0 + k]; else if (i == n) va = d.rhs[k0 + k];
    if (j < n) vb = d.S[(size_t)j * n + k0 + k];
    As[k][r] = va; Bs[k][r] = vb;
  }
  __syncthreads();
  const int ti = (tid >> 4) * 4, tj = (tid & 15) * 4;
  double acc[4][4] = {};
  for (int k = 0; k < kb; k++) {
    double av[4], bv[4];
#pragma unroll
    for (int x = 0; x < 4; x++) { av[x] = As[k][ti + x]; bv[x] = Bs[k][tj + x]; }
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
      for (int y = 0; y < 4; y++) acc[x][y] += av[x] * bv[y];
  }
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int i = i0 + ti + x;
    if (i > n) continue;
#pragma unroll
    for (int y = 0; y < 4; y++) {
      const int j = j0 + tj + y;
      if (j >= n || (i < n && j > i)) continue;
      if (i < n) d.S[(size_t)i * n + j] -= acc[x][y]; else d.rhs[j] -= acc[x][y];
    }
  }
}

// back substitution, panel by panel from the last: x_k = inv(L_kk)' y_k, then y_i -= L_ki' x_k for i < k0
__global__ void __launch_bounds__(256) k_backsolve_panel(BaDev d, int k0, int kb, const double* __restrict__ Linv, int c_begin) { pdl_begin();
  __shared__ double xk[kNB];
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int n = d.nc, tid = threadIdx.x;
  const double* Lp = Linv + (size_t)(k0 / kNB) * kNB * kNB;
  if (tid < kb) {
    double s = 0.0;
    for (int r = tid; r < kb; r++) s += Lp[r * kNB + tid] * d.rhs[k0 + r];   // (inv L)' y
    xk[tid] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < kb) d.yc[k0 + tid] = xk[tid];
  const int c = c_begin + blockIdx.x * 256 + tid;   // columns c_begin <= c < k0 (the panel rows are zero left of c_begin)
  if (c < k0) {
    double s = 0.0;
    for (int r = 0; r < kb; r++) s += d.S[(size_t)(k0 + r) * n + c] * xk[r];
    d.rhs[c] -= s;
  }
}

// The whole back substitution in ONE launch (one CTA, the forward-substituted rhs in shared memory): per panel, from the last,
// x_k = inv(L_kk)' y_k by 16 threads per unknown, then y_c -= L_kc' x_k for the columns c in the panel's row envelope by
// 16 row groups x 64 columns with a fixed-order reduction.  Replaces one k_backsolve_panel launch per panel.
constexpr int kBackAllMaxN = 12288;
inline size_t back_all_smem(int n) { return ((size_t)n + kNB + 16 * 64) * sizeof(double); }
inline bool use_back_all(int n) {      // CMOS_BA_PANEL_BACKSOLVE=1 forces the per-panel kernels (the path of systems beyond kBackAllMaxN)
  const char* e = std::getenv("CMOS_BA_PANEL_BACKSOLVE");
  return n <= kBackAllMaxN && !(e && e[0] == '1');
}
__global__ void __launch_bounds__(1024) k_backsolve_all(BaDev d, const double* __restrict__ Linv, const int* __restrict__ first_col) { pdl_begin();
  extern __shared__ double dyn_smem[];
  double* y = dyn_smem;                       // [n]
  double* xk = y + d.nc;                      // [kNB]
  double* red = xk + kNB;                     // [16][64]
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int n = d.nc, tid = threadIdx.x;
  for (int i = tid; i < n; i += 1024) y[i] = d.rhs[i];
  __syncthreads();
  for (int k0 = ((n - 1) / kNB) * kNB; k0 >= 0; k0 -= kNB) {
    const int kb = min(kNB, n - k0);
    const double* Lp = Linv + (size_t)(k0 / kNB) * kNB * kNB;
    {
      const int t = tid >> 4, part = tid & 15;          // unknown t of the panel, 16 partial sums over r
      double s = 0.0;
      if (t < kb)
        for (int r = t + part; r < kb; r += 16) s += Lp[r * kNB + t] * y[k0 + r];
#pragma unroll
      for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0 && t < kb) xk[t] = s;
    }
    __syncthreads();
    if (tid < kb) { y[k0 + tid] = xk[tid]; d.yc[k0 + tid] = xk[tid]; }
    const int c0 = min(first_col[k0 / kNB], k0);
    const int cl = tid & 63, rg = tid >> 6;             // column lane, row group (4 rows each)
    for (int cb = c0; cb < k0; cb += 64) {
      const int c = cb + cl;
      double s = 0.0;
      if (c < k0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int r = 4 * rg + q;
          if (r < kb) s += d.S[(size_t)(k0 + r) * n + c] * xk[r];
        }
      }
      red[rg * 64 + cl] = s;
      __syncthreads();
      if (tid < 64 && cb + tid < k0) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 16; g++) t += red[g * 64 + tid];
        y[cb + tid] -= t;
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// Banded reduced systems (GlobalBundleAdjustemnt over a windowed co-visibility graph: keyframes a, b share points only
// when |a - b| <= W): ONE persistent CTA walks down the block columns with the (W+1) x (W+1)-block active window of
// the trailing matrix in shared memory, addressed circularly (block b lives in slot b mod (W+1), so the block row
// that enters when column j retires takes column j's slots).  Per block column: 6x6 diagonal factor in the registers
// of one thread, the <= W block rows below solve against it, rank-6 update of the window, the rhs rides along
// (forward substitution), the L block column goes to HBM for the back substitution at the end.  No dense S, no
// memset, one launch per LM iteration instead of ~380.
struct BandArgs {
  int W;                    // block half-bandwidth
  const int* band_blk;      // [Kv][W+1]: id of block (b - off, b), or -1
  double* Lcol;             // [Kv][(W+1)*6][6] L block columns (row slot-major), reuses the dense-S allocation
};
constexpr int kBandThreads = 512;
constexpr int kBandMaxW = 24;
inline size_t band_smem_bytes(int W) {
  const size_t NW = 6 * (size_t)(W + 1);
  return (NW * (NW + 1) + NW * 6 + NW) * sizeof(double) + (size_t)(W * (W + 1) / 2 + 4) * sizeof(uint16_t);
}

// 6x6 Cholesky of the block at window slot sj (lower triangle in Wm) + forward substitution of its rhs block, in the
// registers of the calling thread.  Leaves L11 in s_D, 1/diag in s_rd, y in s_y; returns false if not positive definite.
__device__ __forceinline__ bool band_diag(const double* Wm, int ld, int sj, const double* r6, double (*s_D)[7], double* s_rd,
                                          double* s_y) {
  double A[6][6];
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) A[r][c] = Wm[(size_t)(sj + r) * ld + sj + c];
  bool ok = true;
#pragma unroll
  for (int jj = 0; jj < 6; jj++) {
    const double djj = A[jj][jj];
    if (!(djj > 0.0) || !isfinite(djj)) ok = false;
    const double inv = rsqrt(djj);
    A[jj][jj] = djj * inv;
    s_rd[jj] = inv;
#pragma unroll
    for (int r = jj + 1; r < 6; r++) A[r][jj] *= inv;
#pragma unroll
    for (int c = jj + 1; c < 6; c++)
#pragma unroll
      for (int r = c; r < 6; r++) A[r][c] -= A[r][jj] * A[c][jj];
  }
  double y[6];
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double v = r6[r];
#pragma unroll
    for (int q = 0; q < r; q++) v -= A[r][q] * y[q];
    y[r] = v * s_rd[r];
    s_y[r] = y[r];
  }
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = 0; c < 6; c++) s_D[r][c] = c <= r ? A[r][c] : 0.0;
  return ok;
}

__global__ void __launch_bounds__(kBandThreads) k_solve_band(BaDev d, BandArgs ba) { pdl_begin();
  extern __shared__ __align__(16) double smem_d[];
  LmState& st = *d.st;
  if (st.done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Kv = d.Kv, W = ba.W, NB = W + 1, NW = 6 * NB, ld = NW + 1;
  double* Wm = smem_d;                       // [NW][ld] window, entry (i, k) at [i % NW][k % NW], lower triangle
  double* X = Wm + (size_t)NW * ld;          // [NW][6] solved block column, by row slot
  double* s_r = X + (size_t)NW * 6;          // [NW] rhs of the window rows, by row slot
  uint16_t* s_pair = (uint16_t*)(s_r + NW);  // [W(W+1)/2] block pairs (bi << 8 | bk), bk <= bi, of the rank-6 update
  // double-buffered per-column results of the look-ahead diagonal factorisation
  __shared__ double s_D[2][6][7], s_rd[2][6], s_y[2][6];
  __shared__ int s_slot[2][kBandMaxW + 1];   // row slot of block j + 1 + b, per column parity (no runtime modulo inside)
  __shared__ int s_fail;
  const int n_pairs = W * (W + 1) / 2;
  for (int e = tid; e < n_pairs; e += kBandThreads) {
    int bi = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= e) bi++;
    while (bi * (bi + 1) / 2 > e) bi--;
    s_pair[e] = (uint16_t)((bi << 8) | (e - bi * (bi + 1) / 2));
  }
  if (tid == 0) s_fail = st.solve_failed;
  if (tid <= W) s_slot[0][tid] = ((1 + tid) % NB) * 6;
  // initial window: block rows 0..W.  Sblk entry (r, c) of block (a = b - off, b) -> window (row 6b + c, col 6a + r)
  for (int b = 0; b < min(NB, Kv); b++) {
    for (int e = tid; e < NB * 36; e += kBandThreads) {
      const int off = e / 36, rc = e - 36 * off, r = rc / 6, c = rc - 6 * r, a2 = b - off;
      if (a2 < 0 || (a2 == b && c < r)) continue;
      const int id = ba.band_blk[(size_t)b * NB + off];
      Wm[(size_t)((b % NB) * 6 + c) * ld + (a2 % NB) * 6 + r] = id >= 0 ? d.Sblk[(size_t)id * 36 + rc] : 0.0;
    }
    if (tid < 6) s_r[(b % NB) * 6 + tid] = d.rhs[6 * b + tid];
  }
  __syncthreads();
  if (tid == 0 && !s_fail && !band_diag(Wm, ld, 0, s_r, s_D[0], s_rd[0], s_y[0])) s_fail = 1;
  // per-thread constants of the entering block row (elements e = tid, tid + 512 of its (W+1) x 36 entries)
  int pf_off[2], pf_r[2], pf_c[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int e = tid + q * kBandThreads;
    pf_off[q] = e < NB * 36 ? e / 36 : -1;
    const int rc = e - 36 * (e / 36);
    pf_r[q] = rc / 6; pf_c[q] = rc - 6 * (rc / 6);
  }
  int row_slot = ((6 + tid) % NW);                            // slot of row 6 (j + 1) + tid for the row solves
  int sj = 0;
  __syncthreads();
  for (int j = 0; j < Kv && !s_fail; j++) {
    const int cur = j & 1;
    const int nbelow = min(W, Kv - 1 - j);
    const int* slot = s_slot[cur];
    // the block row that enters when column j retires: fetch early, the loads fly during the row solves
    const int bn = j + NB;
    double pre[2] = {0.0, 0.0}, pre_r = 0.0;
    if (bn < Kv) {
#pragma unroll
      for (int q = 0; q < 2; q++)
        if (pf_off[q] >= 0) {
          const int id = ba.band_blk[(size_t)bn * NB + pf_off[q]];
          if (id >= 0) pre[q] = d.Sblk[(size_t)id * 36 + pf_r[q] * 6 + pf_c[q]];
        }
      if (tid < 6) pre_r = d.rhs[6 * bn + tid];
    }
    // rows below: x = w * L11^-T (6 entries per row), one thread per row; the rhs follows
    double* Lc = ba.Lcol + (size_t)j * NW * 6;
    if (tid < 6 * nbelow) {
      const int si = row_slot;
      double x[6], dot = 0.0;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        double v = Wm[(size_t)si * ld + sj + c];
#pragma unroll
        for (int q = 0; q < c; q++) v -= x[q] * s_D[cur][c][q];
        x[c] = v * s_rd[cur][c];
        dot += x[c] * s_y[cur][c];
      }
      double2* xo = (double2*)(X + si * 6);
      xo[0] = make_double2(x[0], x[1]); xo[1] = make_double2(x[2], x[3]); xo[2] = make_double2(x[4], x[5]);
      double2* lo = (double2*)(Lc + (size_t)tid * 6);
      lo[0] = make_double2(x[0], x[1]); lo[1] = make_double2(x[2], x[3]); lo[2] = make_double2(x[4], x[5]);
      s_r[si] -= dot;
    }
    row_slot += 6; if (row_slot >= NW) row_slot -= NW;
    if (tid >= kBandThreads - 36) {              // the diagonal block (reciprocal diagonal) and y_j, for the back substitution
      const int q = tid - (kBandThreads - 36), r = q / 6, c = q - 6 * r;
      Lc[(size_t)(6 * W + r) * 6 + c] = r == c ? s_rd[cur][r] : s_D[cur][r][c];
      if (c == 0) d.rhs[6 * j + r] = s_y[cur][r];
    } else if (tid >= 256 && tid <= 256 + W) {   // slots of the next column's rows
      int v = slot[tid - 256] + 6; if (v >= NW) v -= NW;
      s_slot[cur ^ 1][tid - 256] = v;
    }
    __syncthreads();
    // rank-6 update of the window rows below.  Warp 0 takes the next diagonal block first and then factors it
    // (look-ahead) while the other warps update the rest.
    if (warp == 0) {
      if (nbelow > 0) {
        const int s1 = slot[0];
        for (int rc = lane; rc < 36; rc += 32) {
          const int r = rc / 6, c = rc - 6 * r;
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < 6; q++) v += X[(s1 + r) * 6 + q] * X[(s1 + c) * 6 + q];
          Wm[(size_t)(s1 + r) * ld + s1 + c] -= v;
        }
        __syncwarp();
        if (lane == 0 && !band_diag(Wm, ld, s1, s_r + s1, s_D[cur ^ 1], s_rd[cur ^ 1], s_y[cur ^ 1])) s_fail = 1;
      }
    } else {
      // work item = (block pair, 3x3 quadrant): 54 independent FMAs on 12 vector loads, enough ILP for the 15 warps
      const int np = nbelow * (nbelow + 1) / 2;
      for (int w = tid - 32; w < 4 * (np - 1); w += kBandThreads - 32) {     // pair 0 = (0,0) is warp 0's
        const int pr = s_pair[1 + (w >> 2)], qd = w & 3;
        const int si = slot[pr >> 8] + 3 * (qd >> 1), sk = slot[pr & 0xff] + 3 * (qd & 1);
        double xi[3][6];
#pragma unroll
        for (int a2 = 0; a2 < 3; a2++) {
          const double2* px = (const double2*)(X + (si + a2) * 6);
          const double2 t0 = px[0], t1 = px[1], t2 = px[2];
          xi[a2][0] = t0.x; xi[a2][1] = t0.y; xi[a2][2] = t1.x; xi[a2][3] = t1.y; xi[a2][4] = t2.x; xi[a2][5] = t2.y;
        }
#pragma unroll
        for (int b2 = 0; b2 < 3; b2++) {
          const double2* px = (const double2*)(X + (sk + b2) * 6);
          const double2 t0 = px[0], t1 = px[1], t2 = px[2];
#pragma unroll
          for (int a2 = 0; a2 < 3; a2++) {
            const double v = ((xi[a2][0] * t0.x + xi[a2][1] * t0.y) + (xi[a2][2] * t1.x + xi[a2][3] * t1.y)) +
                             (xi[a2][4] * t2.x + xi[a2][5] * t2.y);
            Wm[(size_t)(si + a2) * ld + sk + b2] -= v;
          }
        }
      }
    }
    // place the prefetched block row (its slots are those of the retired block j: nobody else touches them)
    if (bn < Kv) {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int off = pf_off[q];
        if (off < 0 || (off == 0 && pf_c[q] < pf_r[q])) continue;
        const int col = off == 0 ? sj : slot[W - off];           // block a = bn - off = j + 1 + (W - off)
        Wm[(size_t)(sj + pf_c[q]) * ld + col + pf_r[q]] = pre[q];
      }
      if (tid < 6) s_r[sj + tid] = pre_r;
    }
    sj += 6; if (sj >= NW) sj -= NW;
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) st.solve_failed = s_fail;
}

// Back substitution L' x = y of the banded factor, from the last block column up: one warp, lane = row below (up to
// four rows of six entries per lane, prefetched one step ahead), x lives in a circular shared array.  Then the
// candidate keyframe poses.
constexpr int kBandBackRows = (6 * kBandMaxW + 31) / 32;     // rows per lane
// TMA bulk copy + mbarrier helpers (descriptor-less cp.async.bulk; see orb.cu for why not the tensor-map form)
__device__ __forceinline__ uint32_t ba_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ba_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ba_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ba_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ba_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {   // 16-byte aligned
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ba_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ba_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ba_mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(ba_smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// Back substitution L' x = y over the band factor, one warp, block column by block column from the last.  The 1000 steps
// are a dependent chain, and with the column fetched from global memory only one step ahead every step waited out an L2
// round trip: the columns now arrive through a ring of kBackRing TMA bulk copies (one contiguous 6(W+1) x 6 panel each,
// completion on an mbarrier), issued kBackRing steps before they are used.
constexpr int kBackRing = 4;
__global__ void __launch_bounds__(32) k_band_backsub(BaDev d, BandArgs ba) { pdl_begin();
  __shared__ double s_x[6 * (kBandMaxW + 1)];
  __shared__ __align__(128) double s_ring[kBackRing][6 * (kBandMaxW + 1) * 6];
  __shared__ __align__(8) uint64_t s_bar[kBackRing];
  LmState& st = *d.st;
  if (st.done || st.solve_failed) return;
  const int lane = threadIdx.x, Kv = d.Kv, W = ba.W, NB = W + 1, NW = 6 * NB;
  const unsigned col_bytes = (unsigned)(NW * 6 * sizeof(double));
  if (lane == 0) {
    for (int r = 0; r < kBackRing; r++) ba_mbar_init(&s_bar[r], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = 0; r < kBackRing && Kv - 1 - r >= 0; r++) {
      ba_mbar_expect_tx(&s_bar[r], col_bytes);
      ba_bulk_load(s_ring[r], ba.Lcol + (size_t)(Kv - 1 - r) * NW * 6, col_bytes, &s_bar[r]);
    }
  }
  for (int i = lane; i < NW; i += 32) s_x[i] = 0.0;
  double ynx = (lane < 6) ? d.rhs[6 * (Kv - 1) + lane] : 0.0;
  __syncwarp();
  int bad = 0;
  int sj = ((Kv - 1) % NB) * 6;               // slot of block j
  int slot = 0;
  unsigned phase = 0;
  for (int j = Kv - 1; j >= 0; j--) {
    ba_mbar_wait(&s_bar[slot], phase);
    const double* Lc = s_ring[slot];
    const int rows = 6 * min(W, Kv - 1 - j);
    double2 cur[kBandBackRows][3];
    double dcur[21];
#pragma unroll
    for (int q = 0; q < kBandBackRows; q++) {
      const int t = lane + 32 * q;
      if (t < rows) {
        const double2* p = (const double2*)(Lc + (size_t)t * 6);
        cur[q][0] = p[0]; cur[q][1] = p[1]; cur[q][2] = p[2];
      } else {
        cur[q][0] = cur[q][1] = cur[q][2] = make_double2(0.0, 0.0);
      }
    }
    if (lane == 0) {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) dcur[k++] = Lc[(size_t)(6 * W + r) * 6 + c];
    }
    const double yj = ynx;
    if (j > 0 && lane < 6) ynx = d.rhs[6 * (j - 1) + lane];
    __syncwarp();                               // every lane has read its part of the ring slot
    if (lane == 0 && j - kBackRing >= 0) {      // refill the slot with the column kBackRing steps ahead
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      ba_mbar_expect_tx(&s_bar[slot], col_bytes);
      ba_bulk_load(s_ring[slot], ba.Lcol + (size_t)(j - kBackRing) * NW * 6, col_bytes, &s_bar[slot]);
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    int sx = sj + 6 + lane;                    // slot of row 6 (j + 1) + lane
    if (sx >= NW) sx -= NW;
#pragma unroll
    for (int q = 0; q < kBandBackRows; q++) {
      const double xi = s_x[sx];               // rows past the band hold zeros in cur[], any finite x will do
      acc[0] += cur[q][0].x * xi; acc[1] += cur[q][0].y * xi; acc[2] += cur[q][1].x * xi;
      acc[3] += cur[q][1].y * xi; acc[4] += cur[q][2].x * xi; acc[5] += cur[q][2].y * xi;
      sx += 32;
      while (sx >= NW) sx -= NW;
    }
#pragma unroll
    for (int c = 0; c < 6; c++)
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    double y[6];
#pragma unroll
    for (int r = 0; r < 6; r++) y[r] = __shfl_sync(0xffffffffu, yj, r);
    if (lane == 0) {
      double x[6];
#pragma unroll
      for (int r = 5; r >= 0; r--) {
        double v = y[r] - acc[r];
#pragma unroll
        for (int q = r + 1; q < 6; q++) v -= dcur[q * (q + 1) / 2 + r] * x[q];
        x[r] = v * dcur[r * (r + 1) / 2 + r];       // reciprocal diagonal
      }
#pragma unroll
      for (int r = 0; r < 6; r++) { s_x[sj + r] = x[r]; d.yc[6 * j + r] = x[r]; bad |= !isfinite(x[r]); }
    }
    sj -= 6; if (sj < 0) sj += NW;
    if (++slot == kBackRing) { slot = 0; phase ^= 1u; }
    __syncwarp();
  }
  bad = __any_sync(0xffffffffu, bad);
  if (bad) { if (lane == 0) st.solve_failed = 1; return; }
  __syncwarp();
  for (int cam = lane; cam < d.K; cam += 32) cam_candidate(d, st, cam);
}

}  // namespace cmos
#include "band_cr.cuh"
namespace cmos {

__global__ void __launch_bounds__(256) k_cam_candidates(BaDev d) { pdl_begin();
  LmState& st = *d.st;
  if (st.done) return;
  const int cam = blockIdx.x * 256 + threadIdx.x;
  if (cam == 0) {
    // validate the solution (finite), single thread: cheap relative to the factorisation
  }
  if (cam < d.K && !st.solve_failed) {
    const int a = d.cam_var[cam];
    if (a >= 0) {
      bool fin = true;
      for (int k = 0; k < 6; k++) fin = fin && isfinite(d.yc[6 * a + k]);
      if (!fin) { st.solve_failed = 1; return; }
    }
    cam_candidate(d, st, cam);
  }
}

}  // namespace cmos

#include "posegraph.cuh"

using namespace cmos;

// =================================================================================================
// Host side
// =================================================================================================
// NCCL is bound at run time (dlopen) so that the library loads on hosts without it; only the sharded
// global bundle adjustment needs it.
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy && GetErrorString;
  }
};
NcclApi g_nccl;
constexpr int kNcclDouble = 8, kNcclUint8 = 1, kNcclSum = 0, kNcclMax = 2;   // ncclFloat64, ncclUint8, ncclSum, ncclMax (nccl.h)
}  // namespace

struct cmos_ba {
  cmos_ba_params p{};
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  int* d_var_cam = nullptr;
  double* d_red = nullptr;
  double* d_HG = nullptr;          // [Kv][21] H_cc then [Kv][6] g_c, contiguous (one all-reduce)
  double* d_Sblk = nullptr;        // [n_blocks][36] then rhs [nc], contiguous (one all-reduce)
  cudaStream_t stream = nullptr;
  // set_problem staging: every array of a problem is packed into ONE page-locked buffer, uploaded with one copy and dealt
  // to its device array by one kernel (it was ~25 pageable copies of ~15 us each: a third of a LocalBundleAdjustment call)
  uint8_t *h_stage = nullptr, *d_stage = nullptr;
  size_t cap_stage = 0;
  size_t pose_stage[5] = {0, 0, 0, 0, 0};      // offsets of the pose-optimisation call's packed layout
  // topology of the last problem (keyframe flags + observation index arrays): an identical one skips the structure build
  std::vector<uint8_t> topo_flags;
  std::vector<int> topo_obs_cam, topo_obs_pt, topo_perm;
  bool topo_valid = false;
  // capacities
  size_t cap_pairs = 0, cap_blocks = 0, cap_S = 0;
  // device storage
  BaDev d{};
  double *d_cams0 = nullptr, *d_pts0 = nullptr;      // initial values (restore)
  double *d_cams_out = nullptr, *d_pts_out = nullptr;
  int *d_cam_var = nullptr, *d_o_cam = nullptr, *d_o_cv = nullptr, *d_o_pt = nullptr, *d_pt_start = nullptr;
  int *d_cam_start = nullptr, *d_cam_obs = nullptr, *d_blk_a = nullptr, *d_blk_b = nullptr, *d_blk_start = nullptr;
  int *d_pair_a = nullptr, *d_pair_b = nullptr, *d_perm = nullptr;
  float2* d_o_uv = nullptr;
  float* d_o_w = nullptr;
  uint8_t *d_o_mode = nullptr, *d_cam_flags = nullptr, *d_erase = nullptr;
  double* d_Linv = nullptr;
  // OptimizeSim3 staging (allocated on first use)
  size_t sim3_cap = 0;
  float *ds_obs = nullptr, *ds_sig = nullptr;   // [2][cap][2], [2][cap]
  double *ds_pts = nullptr, *ds_out = nullptr;   // [2][cap][3], [24]
  uint8_t* ds_bad = nullptr;
  // OptimizeEssentialGraph storage (allocated on first use, grown on demand)
  struct EgStore {
    size_t cap_kf = 0, cap_e = 0, cap_n = 0, cap_pts = 0, cap_tiles = 0, cap_S = 0;
    int plan_info[6] = {0, 0, 0, 0, 0, 0};   // last call: nested dissection used, Wb, node size, nodes, border keyframes, border size
    double *x0 = nullptr, *x1 = nullptr, *x_init = nullptr, *Scw = nullptr, *Snc = nullptr, *lie_out = nullptr, *Tiw = nullptr;
    uint8_t *flags = nullptr, *kind = nullptr;
    int *var = nullptr, *var_kf = nullptr, *ej = nullptr, *ei = nullptr, *inc_start = nullptr, *inc_edge = nullptr, *inc_other = nullptr,
        *inc_sign = nullptr, *tiles = nullptr, *ref = nullptr, *first_col = nullptr, *pos = nullptr;
    Sim3D *meas = nullptr, *Swc = nullptr;
    double *r = nullptr, *J = nullptr, *A = nullptr, *v = nullptr, *Hd = nullptr, *g = nullptr, *scale = nullptr, *delta = nullptr,
           *S = nullptr, *rhs = nullptr, *yc = nullptr, *Linv = nullptr, *pe = nullptr, *pk = nullptr, *Xw = nullptr, *Xo = nullptr;
    LmState* st = nullptr;
    int* done_host = nullptr;      // page-locked
    void release() {
      for (void* b : {(void*)x0, (void*)x1, (void*)x_init, (void*)Scw, (void*)Snc, (void*)lie_out, (void*)Tiw, (void*)flags, (void*)kind,
                      (void*)var, (void*)var_kf, (void*)ej, (void*)ei, (void*)inc_start, (void*)inc_edge, (void*)inc_other, (void*)inc_sign,
                      (void*)tiles, (void*)first_col, (void*)pos, (void*)ref, (void*)meas, (void*)Swc, (void*)r, (void*)J, (void*)A, (void*)v, (void*)Hd, (void*)g,
                      (void*)scale, (void*)delta, (void*)S, (void*)rhs, (void*)yc, (void*)Linv, (void*)pe, (void*)pk, (void*)Xw, (void*)Xo,
                      (void*)st})
        if (b) cudaFree(b);
      if (done_host) cudaFreeHost(done_host);
      *this = EgStore();
    }
  } eg;
  int* d_pan_tiles = nullptr;               // active row tiles of every panel of the blocked factorisation
  int* d_pan_first = nullptr;               // [n_panels] first column of every panel's row envelope (k_backsolve_all)
  std::vector<int> pan_start, pan_first_col; // [n_panels + 1], [n_panels]
  size_t cap_pan_tiles = 0;
  int band_W = 0;                            // > 0: banded reduced system, solved by k_solve_band
  bool schur_chunked = true;                 // k_schur_chunks + k_schur_combine (default) or the CTA-per-block k_schur
  int *d_chunk_start = nullptr, *d_chunk_blk = nullptr, *d_blk_chunk0 = nullptr;
  double* d_schur_part = nullptr;
  size_t cap_chunks = 0, cap_chunk_blocks = 0;
  int cr_N = 0, cr_n = 0, cr_Wb = 0, cr_levels = 0;   // cr_N > 0: banded system solved by block cyclic reduction (band_cr.cuh)
  int* d_band_blk = nullptr;                 // [Kv][band_W + 1]
  double* d_trace = nullptr;       // [2][trace_rows][8]
  int trace_rows = 0;
  cmos_ba_summary* d_summaries = nullptr;   // [2]
  bool has_problem = false, ran = false;
  int dbg_stop_pass = -1, dbg_stop_iter = -1;
  int launches = 0;
  // pose optimisation staging
  double *dp_pose = nullptr, *dp_xw = nullptr, *dp_trace = nullptr;
  float *dp_uv = nullptr, *dp_w = nullptr;
  int *dp_n = nullptr, *dp_inl = nullptr;
  uint8_t* dp_out = nullptr;
  cmos_ba_summary* dp_sum = nullptr;
  // stop flag mapping
  const void* stop_host_page = nullptr;
  uint8_t* stop_dev_page = nullptr;
  bool stop_registered_by_us = false;
  StageTimer timer;
};

namespace {

template <typename T>
bool alloc(T** p, size_t n) {
  return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)) == cudaSuccess;
}

int map_stop_flag(cmos_ba* h, const uint8_t* flag, const volatile uint8_t** dev) {
  *dev = nullptr;
  if (!flag) return CMOS_OK;
  const uintptr_t page = (uintptr_t)flag & ~(uintptr_t)4095;
  if (h->stop_host_page != (const void*)page) {
    if (h->stop_registered_by_us && h->stop_host_page) cudaHostUnregister((void*)h->stop_host_page);
    h->stop_registered_by_us = false;
    h->stop_host_page = nullptr;
    cudaError_t e = cudaHostRegister((void*)page, 4096, cudaHostRegisterMapped);
    if (e == cudaSuccess) h->stop_registered_by_us = true;
    else if (e != cudaErrorHostMemoryAlreadyRegistered) {
      cudaGetLastError();
      set_error("cannot map the stop flag into the device address space: %s", cudaGetErrorString(e));
      return CMOS_ERR_CUDA;
    }
    cudaGetLastError();
    void* dp = nullptr;
    CMOS_CUDA_OK(cudaHostGetDevicePointer(&dp, (void*)page, 0));
    h->stop_host_page = (const void*)page;
    h->stop_dev_page = (uint8_t*)dp;
  }
  *dev = h->stop_dev_page + ((uintptr_t)flag - page);
  return CMOS_OK;
}

// Enqueue one ceres::Solve: max_iterations LM rounds driven by the device-side state machine.
int enqueue_solve(cmos_ba* h, int max_iterations, int pass, cudaStream_t st) {
  BaDev d = h->d;
  d.trace = h->d_trace + (size_t)pass * h->trace_rows * kTraceCols;
  d.pass = pass; d.dbg_stop_pass = h->dbg_stop_pass; d.dbg_stop_iter = h->dbg_stop_iter;
  const int nlb = d.n_lin_blocks;
  const int g_par = (std::max(7 * d.K, 3 * d.M) + 255) / 256;
  const bool small = d.nc <= kSmallMaxN;
  const size_t small_smem = ((size_t)(d.nc + 1) * (d.nc + 2) / 2 + chol_scratch_doubles(d.nc)) * sizeof(double);
  const bool multi = d.multi != 0;
  auto allreduce = [&](double* buf, size_t count, int op) -> int {
    const int rc = g_nccl.AllReduce(buf, buf, count, kNcclDouble, op, h->comm, st);
    if (rc != 0) { set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString(rc)); return CMOS_ERR_CUDA; }
    return CMOS_OK;
  };
  const bool fuse_prep = !multi && d.Kv > 0 && getenv("CMOS_BA_NO_FUSED_PREP") == nullptr;
  auto linearize = [&](bool with_prep) -> int {
    launch_chain(k_linearize, nlb, kLinThreads, 0, st, d);
    if (with_prep && fuse_prep) launch_chain(k_cam_blocks_prep, d.Kv + (d.M + kCamThreads - 1) / kCamThreads, kCamThreads, 0, st, d);
    else if (d.Kv > 0) launch_chain(k_cam_blocks, d.Kv, kCamThreads, 0, st, d);
    h->launches += 2;
    if (!multi) {
      if (d.Kv == 0) { launch_chain(k_post_lin, 1, 256, 0, st, d, 0); h->launches++; }   // otherwise the last CTA of k_cam_blocks ran it
      return CMOS_OK;
    }
    int rc;
    launch_chain(k_post_lin, 1, 256, 0, st, d, 1);
    // ONE message: H_cc and g_c of every keyframe, then cost, |x|^2 and the per-rank gradient-max slots
    if ((rc = allreduce(h->d_HG, (size_t)27 * d.Kv + 2 + d.n_ranks, kNcclSum))) return rc;
    if (d.Kv > 0) launch_chain(k_cam_finish, (d.Kv + 127) / 128, 128, 0, st, d);
    launch_chain(k_post_lin, 1, 256, 0, st, d, 2);
    h->launches += 3;
    return CMOS_OK;
  };
  launch_chain(k_lm_init, 1, 1, 0, st, d, max_iterations);
  h->launches++;
  NvtxRange nvtx_solve(pass == 0 ? "cmos.ba.solve.pass0" : "cmos.ba.solve.pass1");
  for (int it = 0; it < max_iterations; it++) {
    NvtxRange nvtx_it("cmos.ba.lm_iteration");
    int rc;
    if ((rc = linearize(true))) return rc;
    if (!fuse_prep) {
      launch_chain(k_point_prep, (d.M + kLinThreads - 1) / kLinThreads, kLinThreads, 0, st, d);
      h->launches++;
    }
    if (d.Kv > 0) {
      if (h->schur_chunked) {
        launch_chain(k_schur_chunks, (d.n_chunks + kSchurWarps - 1) / kSchurWarps, 32 * kSchurWarps, 0, st, d);
        // (summing the chunks inside k_solve_small instead — one launch less — was measured: one CTA walks 8 k entries x ~5
        // dependent global loads, +13 us per iteration; the 190-CTA kernel takes 7)
        launch_chain(k_schur_combine, (d.n_blocks * 48 + 255) / 256, 256, 0, st, d);
        h->launches += 2;
      } else {
        launch_chain(k_schur, d.n_blocks, kSchurThreads, 0, st, d);
        h->launches++;
      }
      // the one exchange step of the sharded solve: partial reduced camera system + rhs summed over ranks
      if (multi && (rc = allreduce(d.Sblk, (size_t)d.n_blocks * 36 + d.nc, kNcclSum))) return rc;
      if (small) {
        launch_chain(k_solve_small, 1, kSolveThreads, small_smem, st, d);
        h->launches++;
      } else if (h->band_W > 0 && h->cr_N > 0) {
        // nested-dissection (block cyclic reduction) Cholesky of the banded system: log2(N) levels of dense node
        // factorisations, all nodes of a level in parallel
        const CrArgs ca{h->cr_n, h->cr_Wb, h->cr_N, h->band_W, h->cr_levels, h->d_band_blk, d.S};
        const int nt = ca.n / 24, tile_ctas = (nt * nt + kCrGemmWarps - 1) / kCrGemmWarps;
        launch_chain(k_cr_assemble, dim3(ca.N, kCrAsmSplit), 256, 0, st, d, ca);
        h->launches++;
        for (int l = 1; l <= ca.levels; l++) {
          const int cnt = ((ca.N >> (l - 1)) + 1) / 2;
          launch_chain(k_cr_factor, cnt, kSolveThreads, cr_factor_smem(ca.n), st, d, ca, l);
          h->launches++;
          if (l < ca.levels) {
            launch_chain(k_cr_spike, dim3(ca.n / kCrSlab / cr_spike_warps(ca.n), 2, cnt), 32 * cr_spike_warps(ca.n), cr_spike_smem(ca.n), st, d, ca, l, 0);
            launch_chain(k_cr_schur, dim3(tile_ctas, 4, cnt), 32 * kCrGemmWarps, 0, st, d, ca, l);
            h->launches += 2;
          }
        }
        for (int l = ca.levels; l >= 1; l--) {
          const int cnt = ((ca.N >> (l - 1)) + 1) / 2;
          launch_chain(k_cr_back, cnt, 32 * kCrBackWarps, cr_back_smem(ca.n), st, d, ca, l);
          h->launches++;
        }
        launch_chain(k_cam_candidates, (d.K + 255) / 256, 256, 0, st, d);
        h->launches++;
      } else if (h->band_W > 0) {
        BandArgs ba{h->band_W, h->d_band_blk, d.S};
        launch_chain(k_solve_band, 1, kBandThreads, band_smem_bytes(h->band_W), st, d, ba);
        launch_chain(k_band_backsub, 1, 32, 0, st, d, ba);
        h->launches += 2;
      } else {
        const int n = d.nc;
        CMOS_CUDA_OK(cudaMemsetAsync(d.S, 0, (size_t)n * n * sizeof(double), st));
        launch_chain(k_scatter_S, (d.n_blocks * 36 + 255) / 256, 256, 0, st, d);
        h->launches++;
        for (int k0 = 0, p = 0; k0 < n; k0 += kNB, p++) {
          const int kb = std::min(kNB, n - k0);
          launch_chain(k_potrf_diag, 1, kPotrfThreads, kPanelSmem, st, d, k0, kb, h->d_Linv);
          const int na = h->pan_start[p + 1] - h->pan_start[p];   // active tiles below this panel (rhs row included)
          if (na > 0) {
            const int* tiles = h->d_pan_tiles + h->pan_start[p];
            launch_chain(k_trsm_panel, 2 * na, 256, kPanelSmem, st, d, k0, kb, h->d_Linv, tiles);
            launch_chain(k_syrk_tile, na * (na + 1) / 2, 256, kPanelSmem, st, d, k0, kb, tiles);
            h->launches += 2;
          }
          h->launches++;
        }
        if (use_back_all(n)) {
          launch_chain(k_backsolve_all, 1, 1024, back_all_smem(n), st, d, h->d_Linv, h->d_pan_first);
          h->launches++;
        } else
          for (int k0 = ((n - 1) / kNB) * kNB; k0 >= 0; k0 -= kNB) {
            const int kb = std::min(kNB, n - k0);
            const int c0 = std::min(h->pan_first_col[k0 / kNB], k0);
            launch_chain(k_backsolve_panel, std::max(1, (k0 - c0 + 255) / 256), 256, 0, st, d, k0, kb, h->d_Linv, c0);
            h->launches++;
          }
        launch_chain(k_cam_candidates, (d.K + 255) / 256, 256, 0, st, d);
        h->launches++;
      }
    }
    launch_chain(k_backsub, nlb, kLinThreads, 0, st, d);
    h->launches++;
    if (!multi) {
      // the last CTA of k_backsub decides
    } else {
      launch_chain(k_decide, 1, 256, 0, st, d, 1);
      if ((rc = allreduce(d.red + 3, 4, kNcclSum))) return rc;                    // candidate cost, model change, |step|^2, failure
      launch_chain(k_decide, 1, 256, 0, st, d, 2);
      h->launches += 2;
    }
  }
  if (max_iterations == 0) {   // Ceres still evaluates iteration 0
    int rc;
    if ((rc = linearize(false))) return rc;
  }
  launch_chain(k_summary, 1, 1, 0, st, d, h->d_summaries + pass);
  h->launches++;
  CMOS_CUDA_OK(cudaGetLastError());
  return CMOS_OK;
}

}  // namespace

extern "C" {

int cmos_ba_create(const cmos_ba_params* params, cmos_ba_t* out) {
  CMOS_REQUIRE(params && out, "null argument");
  CMOS_REQUIRE(params->max_cams >= 1 && params->max_points >= 1 && params->max_obs >= 1, "bad sizes");
  CMOS_REQUIRE(params->max_pose_batch >= 0 && params->max_pose_corr >= 0, "bad pose sizes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(params->device >= 0 && params->device < ndev, "device %d out of range", params->device);
  CMOS_CUDA_OK(cudaSetDevice(params->device));
  cmos_ba* h = new cmos_ba();
  h->p = *params;
  const size_t K = params->max_cams, M = params->max_points, N = params->max_obs;
  const size_t pairs_per_obs = params->max_pairs_per_obs > 0 ? params->max_pairs_per_obs : 8;
  h->cap_pairs = N * pairs_per_obs;
  h->cap_blocks = std::min<size_t>(K * (K + 1) / 2, h->cap_pairs) + 1;
  h->cap_S = (6 * K) * (6 * K);
  { const size_t np = (6 * K + kNB) / kNB + 2; h->cap_pan_tiles = np * np; }
  h->trace_rows = 256;
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  BaDev& d = h->d;
  const size_t nlb = ((size_t)M * kPointLanes + kLinThreads - 1) / kLinThreads;
  ok = ok && alloc(&d.cams[0], 7 * K) && alloc(&d.cams[1], 7 * K) && alloc(&d.pts[0], 3 * M) && alloc(&d.pts[1], 3 * M);
  ok = ok && alloc(&h->d_cams0, 7 * K) && alloc(&h->d_pts0, 3 * M) && alloc(&h->d_cams_out, 7 * K) && alloc(&h->d_pts_out, 3 * M);
  ok = ok && alloc(&h->d_cam_var, K) && alloc(&h->d_o_cam, N) && alloc(&h->d_o_cv, N) && alloc(&h->d_o_pt, N) &&
       alloc(&h->d_pt_start, M + 1) && alloc(&h->d_cam_start, K + 1) && alloc(&h->d_cam_obs, N) &&
       alloc(&h->d_blk_a, h->cap_blocks) && alloc(&h->d_blk_b, h->cap_blocks) && alloc(&h->d_blk_start, h->cap_blocks + 1) &&
       alloc(&h->d_pair_a, h->cap_pairs) && alloc(&h->d_pair_b, h->cap_pairs) && alloc(&h->d_perm, N) &&
       alloc(&h->d_o_uv, N) && alloc(&h->d_o_w, N) && alloc(&h->d_o_mode, N) && alloc(&h->d_cam_flags, K) &&
       alloc(&h->d_erase, N);
  ok = ok && alloc(&d.Jc, 12 * N) && alloc(&d.Jp, 6 * N) && alloc(&d.res, 2 * N) && alloc(&d.Hpp, 6 * M) && alloc(&d.gp, 3 * M) &&
       alloc(&d.Hinv, 6 * M) && alloc(&d.tp, 3 * M) && alloc(&d.scale_p, 3 * M) && alloc(&h->d_HG, 27 * K + 2 + 64) &&
       alloc(&h->d_var_cam, K) && alloc(&h->d_red, 16) && alloc(&h->d_Sblk, h->cap_blocks * 36 + 6 * K + 8) &&
       alloc(&d.scale_c, 6 * K) && alloc(&d.S, h->cap_S) &&
       alloc(&d.yc, 6 * K + 8) && alloc(&d.part, 6 * nlb + 4 * K + 16) && alloc(&d.st, 1) &&
       alloc(&h->d_Linv, ((6 * K + kNB - 1) / kNB) * kNB * kNB) && alloc(&h->d_pan_tiles, h->cap_pan_tiles) && alloc(&h->d_pan_first, (6 * K + kNB) / kNB + 2) && alloc(&h->d_band_blk, (size_t)K * (kBandMaxW + 1)) && alloc(&h->d_trace, 2 * (size_t)h->trace_rows * kTraceCols) &&
       alloc(&h->d_summaries, 2);
  const size_t PB = std::max(params->max_pose_batch, 1), PC = std::max(params->max_pose_corr, 1);
  ok = ok && alloc(&h->dp_pose, 7 * PB) && alloc(&h->dp_xw, 3 * PB * PC) && alloc(&h->dp_uv, 2 * PB * PC) &&
       alloc(&h->dp_w, PB * PC) && alloc(&h->dp_n, PB) && alloc(&h->dp_inl, PB) && alloc(&h->dp_out, PB * PC) &&
       alloc(&h->dp_sum, PB) && alloc(&h->dp_trace, PB * (size_t)h->trace_rows * kTraceCols);
  if (!ok) {
    set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    cmos_ba_destroy(h);
    return CMOS_ERR_CUDA;
  }
  cudaMemset(d.st, 0, sizeof(LmState));
  cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  cudaFuncSetAttribute(k_solve_band, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  cudaFuncSetAttribute(k_cr_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cr_factor_smem(kCrMaxN));
  cudaFuncSetAttribute(k_cr_spike, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_crb_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)crb_solve_smem(kCrMaxN));
  cudaFuncSetAttribute(k_cr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cr_back_smem(kCrMaxN));
  cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmem);
  cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmem);
  cudaFuncSetAttribute(k_syrk_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmem);
  cudaFuncSetAttribute(k_backsolve_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_all_smem(kBackAllMaxN));
  CMOS_CUDA_OK(cudaMemset(h->d_red, 0, 16 * sizeof(double)));      // incl. the arrival tickets
  // cudaMemset runs asynchronously on the legacy stream, which is NOT ordered with the handle's non-blocking stream
  CMOS_CUDA_OK(cudaDeviceSynchronize());
  CMOS_CUDA_OK(cudaGetLastError());
  *out = h;
  return CMOS_OK;
}

int cmos_ba_destroy(cmos_ba_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->p.device);
  BaDev& d = h->d;
  void* bufs[] = {d.cams[0], d.cams[1], d.pts[0], d.pts[1], h->d_cams0, h->d_pts0, h->d_cams_out, h->d_pts_out, h->d_cam_var,
                  h->d_o_cam, h->d_o_cv, h->d_o_pt, h->d_pt_start, h->d_cam_start, h->d_cam_obs, h->d_blk_a, h->d_blk_b,
                  h->d_blk_start, h->d_pair_a, h->d_pair_b, h->d_perm, h->d_o_uv, h->d_o_w, h->d_o_mode, h->d_cam_flags,
                  h->d_erase, d.Jc, d.Jp, d.res, d.Hpp, d.gp, d.Hinv, d.tp, d.scale_p, h->d_HG, h->d_var_cam, h->d_red, h->d_Sblk, d.scale_c, d.S,
                  d.yc, d.part, d.st, h->d_Linv, h->d_pan_tiles, h->d_pan_first, h->d_band_blk, h->d_trace, h->d_summaries, h->dp_pose, h->dp_xw, h->dp_uv, h->dp_w,
                  h->dp_n, h->dp_inl, h->dp_out, h->dp_sum, h->dp_trace, h->d_chunk_start, h->d_chunk_blk, h->d_blk_chunk0,
                  h->d_schur_part};
  for (void* b : bufs)
    if (b) cudaFree(b);
  for (void* b : {(void*)h->ds_obs, (void*)h->ds_sig, (void*)h->ds_pts, (void*)h->ds_out, (void*)h->ds_bad})
    if (b) cudaFree(b);
  h->eg.release();
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->d_stage) cudaFree(h->d_stage);
  if (h->stop_registered_by_us && h->stop_host_page) cudaHostUnregister((void*)h->stop_host_page);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->stream) cudaStreamDestroy(h->stream);
  h->timer.destroy();
  delete h;
  return CMOS_OK;
}

int cmos_ba_pose_optimization(cmos_ba_t h, int32_t n_frames, double* pose7, const int32_t* n_corr, const double* xw,
                              const float* uv, const float* inv_sigma2, int32_t stride, const float* K4,
                              int32_t max_iterations, uint8_t* is_outlier, int32_t* n_inliers,
                              cmos_ba_summary* summaries, int32_t on_device, void* stream) {
  CMOS_REQUIRE(h && pose7 && n_corr && xw && uv && inv_sigma2 && K4 && is_outlier && n_inliers, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && stride >= 1 && max_iterations >= 0, "bad sizes");
  CMOS_REQUIRE(max_iterations + 1 < h->trace_rows, "max_iterations %d too large", max_iterations);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  PoseArgs a{};
  a.stride = stride;
  a.fx = (double)K4[0]; a.fy = (double)K4[1]; a.cx = (double)K4[2]; a.cy = (double)K4[3];
  a.max_iterations = max_iterations;
  a.trace_rows = h->trace_rows;
  const size_t B = n_frames, BC = B * stride;
  if (on_device) {
    a.pose7 = pose7; a.n_corr = n_corr; a.xw = xw; a.uv = uv; a.inv_sigma2 = inv_sigma2;
    a.is_outlier = is_outlier; a.n_inliers = n_inliers; a.summaries = summaries; a.trace = nullptr;
  } else {
    CMOS_REQUIRE(n_frames <= h->p.max_pose_batch && stride <= h->p.max_pose_corr,
                 "batch %d x %d exceeds the handle's pose capacity %d x %d", n_frames, stride, h->p.max_pose_batch,
                 h->p.max_pose_corr);
    // One page-locked buffer, laid out [xw | uv | w | n | pose | summaries | inliers | outlier flags]: ONE upload of
    // xw..pose, the kernel works in place, ONE download of pose..flags (it was five pageable copies in and four out:
    // 0.13 ms around a 0.06 ms solve, per tracked frame).
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_xw = 0, o_uv = o_xw + al(3 * BC * sizeof(double)), o_w = o_uv + al(2 * BC * sizeof(float)),
                 o_n = o_w + al(BC * sizeof(float)), o_pose = o_n + al(B * sizeof(int)), o_sum = o_pose + al(7 * B * sizeof(double)),
                 o_inl = o_sum + al(B * sizeof(cmos_ba_summary)), o_out = o_inl + al(B * sizeof(int)), total = o_out + al(BC);
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
    if (st != h->stream) CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));     // the buffer is shared with set_problem / get_results
    if (total > h->cap_stage) {
      if (h->h_stage) cudaFreeHost(h->h_stage);
      if (h->d_stage) cudaFree(h->d_stage);
      h->h_stage = h->d_stage = nullptr; h->cap_stage = 0;
      const size_t want = total + total / 4 + 4096;
      CMOS_CUDA_OK(cudaMallocHost((void**)&h->h_stage, want));
      CMOS_CUDA_OK(cudaMalloc((void**)&h->d_stage, want));
      h->cap_stage = want;
    }
    std::memcpy(h->h_stage + o_xw, xw, 3 * BC * sizeof(double));
    std::memcpy(h->h_stage + o_uv, uv, 2 * BC * sizeof(float));
    std::memcpy(h->h_stage + o_w, inv_sigma2, BC * sizeof(float));
    std::memcpy(h->h_stage + o_n, n_corr, B * sizeof(int));
    std::memcpy(h->h_stage + o_pose, pose7, 7 * B * sizeof(double));
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_stage, h->h_stage, o_sum, cudaMemcpyHostToDevice, st));
    uint8_t* ds = h->d_stage;
    a.pose7 = (double*)(ds + o_pose); a.n_corr = (const int*)(ds + o_n); a.xw = (const double*)(ds + o_xw);
    a.uv = (const float*)(ds + o_uv); a.inv_sigma2 = (const float*)(ds + o_w);
    a.is_outlier = ds + o_out; a.n_inliers = (int*)(ds + o_inl); a.summaries = (cmos_ba_summary*)(ds + o_sum); a.trace = h->dp_trace;
    h->pose_stage[0] = o_pose; h->pose_stage[1] = o_sum; h->pose_stage[2] = o_inl; h->pose_stage[3] = o_out; h->pose_stage[4] = total;
  }
  h->timer.begin(st);
  launch_chain(k_pose_opt, n_frames, kPoseThreads, 0, st, a);
  h->timer.mark(st);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  if (!on_device) {
    const size_t* o = h->pose_stage;
    CMOS_CUDA_OK(cudaMemcpyAsync(h->h_stage + o[0], h->d_stage + o[0], o[4] - o[0], cudaMemcpyDeviceToHost, st));
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
    std::memcpy(pose7, h->h_stage + o[0], 7 * B * sizeof(double));
    if (summaries) std::memcpy(summaries, h->h_stage + o[1], B * sizeof(cmos_ba_summary));
    std::memcpy(n_inliers, h->h_stage + o[2], B * sizeof(int));
    std::memcpy(is_outlier, h->h_stage + o[3], BC);
  }
  return CMOS_OK;
}

int cmos_ba_debug_pose_trace(cmos_ba_t h, int32_t frame, double* trace, int32_t rows) {
  CMOS_REQUIRE(h && trace && frame >= 0 && frame < h->p.max_pose_batch && rows >= 1 && rows <= h->trace_rows, "bad argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  CMOS_CUDA_OK(cudaDeviceSynchronize());
  CMOS_CUDA_OK(cudaMemcpy(trace, h->dp_trace + (size_t)frame * h->trace_rows * kTraceCols, (size_t)rows * kTraceCols * sizeof(double),
                          cudaMemcpyDeviceToHost));
  return CMOS_OK;
}

}  // extern "C"

namespace {
constexpr int kStageMax = 28;
struct StageTable { void* dst[kStageMax]; size_t off[kStageMax]; size_t bytes[kStageMax]; int n; };
// one launch deals the packed upload to its destination arrays (offsets are 16-byte aligned, sizes multiples of 4 except
// the keyframe flags)
// (to_stage != 0: the other direction — results gathered into the packed buffer for one download)
__global__ void __launch_bounds__(256) k_stage_scatter(StageTable t, uint8_t* __restrict__ stage, int to_stage) { pdl_begin();
  const size_t gt = (size_t)blockIdx.x * 256 + threadIdx.x, gs = (size_t)gridDim.x * 256;
  for (int e = 0; e < t.n; e++) {
    const uint8_t* s8 = to_stage ? (const uint8_t*)t.dst[e] : stage + t.off[e];
    uint8_t* d8 = to_stage ? stage + t.off[e] : (uint8_t*)t.dst[e];
    const size_t words = t.bytes[e] >> 2;
    if ((((uintptr_t)d8 | (uintptr_t)s8) & 3) == 0) {
      for (size_t i = gt; i < words; i += gs) ((uint32_t*)d8)[i] = ((const uint32_t*)s8)[i];
      for (size_t i = 4 * words + gt; i < t.bytes[e]; i += gs) d8[i] = s8[i];
    } else {
      for (size_t i = gt; i < t.bytes[e]; i += gs) d8[i] = s8[i];
    }
  }
}
struct Stager {
  cmos_ba* h;
  StageTable t{};
  size_t used = 0;
  const void* srcs[kStageMax];
  explicit Stager(cmos_ba* hh) : h(hh) { t.n = 0; }
  void add(void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    t.dst[t.n] = dst; t.off[t.n] = used; t.bytes[t.n] = bytes; srcs[t.n] = src; t.n++;
    used += (bytes + 15) & ~(size_t)15;
  }
  // copies the sources into the page-locked buffer, enqueues the upload and the scatter; host sources may die afterwards
  int flush(cudaStream_t st) {
    if (!t.n) return CMOS_OK;
    if (used > h->cap_stage) {
      CMOS_CUDA_OK(cudaStreamSynchronize(st));
      if (h->h_stage) cudaFreeHost(h->h_stage);
      if (h->d_stage) cudaFree(h->d_stage);
      h->h_stage = h->d_stage = nullptr; h->cap_stage = 0;
      const size_t want = used + used / 4 + 4096;
      CMOS_CUDA_OK(cudaMallocHost((void**)&h->h_stage, want));
      CMOS_CUDA_OK(cudaMalloc((void**)&h->d_stage, want));
      h->cap_stage = want;
    } else {
      CMOS_CUDA_OK(cudaStreamSynchronize(st));      // the previous upload may still be reading the buffer
    }
    for (int e = 0; e < t.n; e++) std::memcpy(h->h_stage + t.off[e], srcs[e], t.bytes[e]);
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_stage, h->h_stage, used, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<size_t>(296, (used / 4 + 255) / 256 + 1);
    launch_chain(k_stage_scatter, grid, 256, 0, st, t, h->d_stage, 0);
    CMOS_CUDA_OK(cudaGetLastError());
    return CMOS_OK;
  }
  // the other direction: add(device array, host destination, bytes) ..., then download() gathers, copies once, and
  // hands every piece to its host destination (synchronises the stream)
  int download(cudaStream_t st) {
    if (!t.n) return CMOS_OK;
    if (used > h->cap_stage) {
      CMOS_CUDA_OK(cudaStreamSynchronize(st));
      if (h->h_stage) cudaFreeHost(h->h_stage);
      if (h->d_stage) cudaFree(h->d_stage);
      h->h_stage = h->d_stage = nullptr; h->cap_stage = 0;
      const size_t want = used + used / 4 + 4096;
      CMOS_CUDA_OK(cudaMallocHost((void**)&h->h_stage, want));
      CMOS_CUDA_OK(cudaMalloc((void**)&h->d_stage, want));
      h->cap_stage = want;
    }
    const int grid = (int)std::min<size_t>(296, (used / 4 + 255) / 256 + 1);
    launch_chain(k_stage_scatter, grid, 256, 0, st, t, h->d_stage, 1);
    CMOS_CUDA_OK(cudaGetLastError());
    CMOS_CUDA_OK(cudaMemcpyAsync(h->h_stage, h->d_stage, used, cudaMemcpyDeviceToHost, st));
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
    for (int e = 0; e < t.n; e++) std::memcpy(const_cast<void*>(srcs[e]), h->h_stage + t.off[e], t.bytes[e]);
    return CMOS_OK;
  }
};
}  // namespace

extern "C" {

int cmos_debug_solve_spd(const double* A, const double* b, int32_t n, int32_t variant, double* x, int32_t* failed, int64_t* cycles2) {
  CMOS_REQUIRE(A && b && x && failed && n >= 6 && n % 6 == 0 && n <= kSmallMaxN, "bad argument (n a multiple of 6, <= %d)", kSmallMaxN);
  double *dA = nullptr, *db = nullptr, *dx = nullptr; int* df = nullptr; long long* dc = nullptr;
  CMOS_CUDA_OK(cudaMalloc(&dA, (size_t)n * n * 8)); CMOS_CUDA_OK(cudaMalloc(&db, n * 8)); CMOS_CUDA_OK(cudaMalloc(&dx, n * 8));
  CMOS_CUDA_OK(cudaMalloc(&df, 4)); CMOS_CUDA_OK(cudaMalloc(&dc, 16));
  CMOS_CUDA_OK(cudaMemcpy(dA, A, (size_t)n * n * 8, cudaMemcpyHostToDevice));
  CMOS_CUDA_OK(cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice));
  const size_t smem = ((size_t)(n + 1) * (n + 2) / 2 + chol_scratch_doubles(n)) * sizeof(double);
  CMOS_CUDA_OK(cudaFuncSetAttribute(k_debug_solve_spd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  for (int rep = 0; rep < 2; rep++)       // the second run reports warm instruction-cache cycles
    k_debug_solve_spd<<<1, kSolveThreads, smem>>>(dA, db, n, dx, df, dc, variant);
  CMOS_CUDA_OK(cudaDeviceSynchronize());
  long long hc[2];
  CMOS_CUDA_OK(cudaMemcpy(x, dx, n * 8, cudaMemcpyDeviceToHost));
  CMOS_CUDA_OK(cudaMemcpy(failed, df, 4, cudaMemcpyDeviceToHost));
  CMOS_CUDA_OK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
  if (cycles2) { cycles2[0] = hc[0]; cycles2[1] = hc[1]; }
  cudaFree(dA); cudaFree(db); cudaFree(dx); cudaFree(df); cudaFree(dc);
  return CMOS_OK;
}

int cmos_ba_set_problem(cmos_ba_t h, int32_t n_cams, const double* cams, const uint8_t* cam_flags, int32_t n_points,
                        const double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                        const float* uv, const float* inv_sigma2, const float* K4) {
  CMOS_REQUIRE(h && cams && cam_flags && points && obs_cam && obs_pt && uv && inv_sigma2 && K4, "null argument");
  CMOS_REQUIRE(n_cams >= 1 && n_cams <= h->p.max_cams && n_points >= 1 && n_points <= h->p.max_points && n_obs >= 1 &&
               n_obs <= h->p.max_obs, "problem %d keyframes / %d points / %d observations exceeds the handle (%d / %d / %d)",
               n_cams, n_points, n_obs, h->p.max_cams, h->p.max_points, h->p.max_obs);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  const int K = n_cams, M = n_points, N = n_obs;
  // ---- same topology as the last problem (keyframe flags and observation index arrays): new values only ----
  // (not for sharded solves: the structure build contains a collective, every rank must take the same path)
  if (h->topo_valid && h->n_ranks == 1 && h->has_problem && h->d.K == K && h->d.M == M && h->d.N == N &&
      std::memcmp(h->topo_flags.data(), cam_flags, K) == 0 &&
      std::memcmp(h->topo_obs_cam.data(), obs_cam, (size_t)N * sizeof(int)) == 0 &&
      std::memcmp(h->topo_obs_pt.data(), obs_pt, (size_t)N * sizeof(int)) == 0 && !std::getenv("CMOS_BA_NO_TOPO_CACHE")) {
    std::vector<float2> o_uv(N);
    std::vector<float> o_w(N);
    for (int p = 0; p < N; p++) { const int i = h->topo_perm[p]; o_uv[p] = make_float2(uv[2 * i], uv[2 * i + 1]); o_w[p] = inv_sigma2[i]; }
    Stager sg(h);
    sg.add(h->d_cams0, cams, 7 * (size_t)K * sizeof(double));
    sg.add(h->d_pts0, points, 3 * (size_t)M * sizeof(double));
    sg.add(h->d_o_uv, o_uv.data(), (size_t)N * sizeof(float2));
    sg.add(h->d_o_w, o_w.data(), (size_t)N * sizeof(float));
    int rc = sg.flush(h->stream);
    if (rc) return rc;
    BaDev& d = h->d;
    d.fx = (double)K4[0]; d.fy = (double)K4[1]; d.cx = (double)K4[2]; d.cy = (double)K4[3];
    d.trace = nullptr; d.stop_flag = nullptr;
    h->ran = false;
    return CMOS_OK;
  }
  // ---- structure (host, counting sorts only) ----
  h->topo_valid = false;             // (stays false if this call fails half way)
  std::vector<int> cam_var(K, -1);
  int Kv = 0;
  for (int k = 0; k < K; k++)
    if (!(cam_flags[k] & 1)) cam_var[k] = Kv++;
  std::vector<int> pt_start(M + 1, 0);
  for (int i = 0; i < N; i++) {
    CMOS_REQUIRE(obs_cam[i] >= 0 && obs_cam[i] < K && obs_pt[i] >= 0 && obs_pt[i] < M, "observation %d out of range", i);
    pt_start[obs_pt[i] + 1]++;
  }
  for (int j = 0; j < M; j++) pt_start[j + 1] += pt_start[j];
  std::vector<int> perm(N), fill(pt_start.begin(), pt_start.end() - 1);
  for (int i = 0; i < N; i++) perm[fill[obs_pt[i]]++] = i;
  std::vector<int> o_cam(N), o_cv(N), o_pt(N);
  std::vector<float2> o_uv(N);
  std::vector<float> o_w(N);
  std::vector<int> cam_start(Kv + 1, 0);
  for (int p = 0; p < N; p++) {
    const int i = perm[p];
    o_cam[p] = obs_cam[i]; o_cv[p] = cam_var[obs_cam[i]]; o_pt[p] = obs_pt[i];
    o_uv[p] = make_float2(uv[2 * i], uv[2 * i + 1]); o_w[p] = inv_sigma2[i];
    if (o_cv[p] >= 0) cam_start[o_cv[p] + 1]++;
  }
  for (int a = 0; a < Kv; a++) cam_start[a + 1] += cam_start[a];
  std::vector<int> var_cam(std::max(Kv, 1), 0);
  for (int k = 0; k < K; k++)
    if (cam_var[k] >= 0) var_cam[cam_var[k]] = k;
  std::vector<int> cam_obs(std::max(cam_start[Kv], 1));
  {
    std::vector<int> f2(cam_start.begin(), cam_start.end() - 1);
    for (int p = 0; p < N; p++)
      if (o_cv[p] >= 0) cam_obs[f2[o_cv[p]]++] = p;
  }
  // pair lists per non-zero block (a <= b)
  CMOS_REQUIRE((size_t)Kv * Kv <= (size_t)64 << 20, "too many variable keyframes (%d) for the dense block table", Kv);
  std::vector<int> table((size_t)Kv * Kv, 0);
  size_t n_pairs = 0;
  for (int j = 0; j < M; j++)
    for (int pa = pt_start[j]; pa < pt_start[j + 1]; pa++) {
      if (o_cv[pa] < 0) continue;
      for (int pb = pt_start[j]; pb < pt_start[j + 1]; pb++) {
        if (o_cv[pb] < 0) continue;
        if (o_cv[pa] < o_cv[pb] || (o_cv[pa] == o_cv[pb])) { table[(size_t)o_cv[pa] * Kv + o_cv[pb]]++; n_pairs++; }
      }
    }
  // Long tracks (sum over points of L (L + 1) / 2 pairs) can exceed what create() sized from max_pairs_per_obs: the
  // pair lists grow on demand instead of rejecting the graph (dense local windows, global BA after many revisits).
  if (n_pairs > h->cap_pairs) {
    const size_t want = n_pairs + n_pairs / 4;
    CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_pair_a); cudaFree(h->d_pair_b);
    h->d_pair_a = h->d_pair_b = nullptr; h->cap_pairs = 0;
    CMOS_REQUIRE(alloc(&h->d_pair_a, want) && alloc(&h->d_pair_b, want),
                 "cannot grow the co-observation pair lists to %zu entries", want);
    h->cap_pairs = want;
  }
  // Sharded solve: the reduced camera system is summed block by block over the ranks, so every rank needs the
  // SAME block list — the union of the co-visibility patterns (a rank's contiguous point range only covers part of
  // the trajectory).  One byte-mask all-reduce (max); blocks without local pairs simply contribute zero.
  std::vector<uint8_t> present((size_t)Kv * Kv, 0);
  for (size_t i = 0; i < present.size(); i++) present[i] = table[i] > 0;
  if (h->n_ranks > 1 && Kv > 0) {
    uint8_t* d_mask = nullptr;
    CMOS_CUDA_OK(cudaMalloc(&d_mask, present.size()));
    CMOS_CUDA_OK(cudaMemcpyAsync(d_mask, present.data(), present.size(), cudaMemcpyHostToDevice, h->stream));
    const int nrc = g_nccl.AllReduce(d_mask, d_mask, present.size(), kNcclUint8, kNcclMax, h->comm, h->stream);
    if (nrc != 0) { cudaFree(d_mask); set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString(nrc)); return CMOS_ERR_CUDA; }
    CMOS_CUDA_OK(cudaMemcpyAsync(present.data(), d_mask, present.size(), cudaMemcpyDeviceToHost, h->stream));
    CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d_mask);
  }
  std::vector<int> blk_a, blk_b, blk_start;
  blk_start.push_back(0);
  for (int a = 0; a < Kv; a++)
    for (int b = a; b < Kv; b++) {
      int& t = table[(size_t)a * Kv + b];
      if (t > 0 || (a != b && present[(size_t)a * Kv + b])) {
        blk_a.push_back(a); blk_b.push_back(b);
        blk_start.push_back(blk_start.back() + t);
        t = (int)blk_a.size();      // 1-based block id
      } else if (a == b) {          // every variable keyframe gets its diagonal block
        blk_a.push_back(a); blk_b.push_back(b);
        blk_start.push_back(blk_start.back());
        t = (int)blk_a.size();
      }
    }
  const int nb = (int)blk_a.size();
  // Row envelope of the reduced system for the blocked factorisation (systems too large for one CTA): per panel of
  // kNB columns the 64-row tiles below it (relative to the first row under the panel) that have an entry in or left
  // of the panel.  Fill stays inside the row envelope, so TRSM / SYRK only visit these tiles.  The rhs row (row n) is
  // dense.
  h->pan_start.assign(1, 0);
  h->pan_first_col.clear();
  h->band_W = 0;
  if (6 * Kv > kSmallMaxN) {
    int Wmax = 0;
    for (int i = 0; i < nb; i++) Wmax = std::max(Wmax, blk_b[i] - blk_a[i]);
    const char* no_band = std::getenv("CMOS_BA_NO_BAND");
    const size_t NWb = 6 * (size_t)(Wmax + 1);
    if (Wmax >= 1 && Wmax <= kBandMaxW && !(no_band && no_band[0] == '1') &&
        band_smem_bytes(Wmax) <= 227 * 1024 - 2048 &&
        (size_t)Kv * NWb * 6 <= h->cap_S) {
      std::vector<int> band((size_t)Kv * (Wmax + 1), -1);
      for (int i = 0; i < nb; i++) band[(size_t)blk_b[i] * (Wmax + 1) + (blk_b[i] - blk_a[i])] = i;
      CMOS_CUDA_OK(cudaMemcpyAsync(h->d_band_blk, band.data(), band.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
      h->band_W = Wmax;
      // Long bands: nested dissection over nodes of Wb >= W blocks (band_cr.cuh).  Wb is a multiple of 4 so the node size
      // is a multiple of the 24 x 24 tensor-core tile; short systems (< 8 nodes) stay on the one-CTA band walk.
      // CMOS_BA_BAND_SERIAL=1 forces the serial walk (A/B runs, tests that compare the two).
      const char* serial = std::getenv("CMOS_BA_BAND_SERIAL");
      const int Wb = (std::max(Wmax, 8) + 3) / 4 * 4, N = (Kv + Wb - 1) / Wb;
      h->cr_N = 0;
      if (!(serial && serial[0] == '1') && 6 * Wb <= kCrMaxN && N >= 8 && cr_doubles(N, 6 * Wb) <= h->cap_S) {
        h->cr_N = N; h->cr_n = 6 * Wb; h->cr_Wb = Wb;
        h->cr_levels = 1;
        while ((1 << h->cr_levels) <= N) h->cr_levels++;
      }
    }
  }
  if (6 * Kv > kSmallMaxN && h->band_W == 0) {
    const int n = 6 * Kv;
    std::vector<int> first_blk(Kv);
    for (int b = 0; b < Kv; b++) first_blk[b] = b;
    for (int i = 0; i < nb; i++) first_blk[blk_b[i]] = std::min(first_blk[blk_b[i]], blk_a[i]);
    std::vector<int> tiles;
    for (int k0 = 0; k0 < n; k0 += kNB) {
      const int kb = std::min(kNB, n - k0), t0 = k0 + kb;
      int fc = k0;
      for (int r = k0; r < k0 + kb; r++) fc = std::min(fc, 6 * first_blk[r / 6]);
      h->pan_first_col.push_back(fc);
      for (int t = 0; t0 + 64 * t <= n; t++) {
        const int r0 = t0 + 64 * t, r1 = std::min(r0 + 63, n);
        bool active = r1 == n;                                   // the rhs row
        for (int r = r0; r <= std::min(r1, n - 1) && !active; r += 6 - r % 6) active = 6 * first_blk[r / 6] < t0;
        if (active) tiles.push_back(t);
      }
      h->pan_start.push_back((int)tiles.size());
    }
    CMOS_REQUIRE(tiles.size() <= h->cap_pan_tiles, "panel tile list %zu exceeds capacity %zu", tiles.size(), h->cap_pan_tiles);
    if (!tiles.empty())
      CMOS_CUDA_OK(cudaMemcpyAsync(h->d_pan_tiles, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_pan_first, h->pan_first_col.data(), h->pan_first_col.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  if ((size_t)nb > h->cap_blocks) {   // same policy for the block list and the block-sparse reduced system
    const size_t want = (size_t)nb + nb / 4 + 1;
    CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_blk_a); cudaFree(h->d_blk_b); cudaFree(h->d_blk_start); cudaFree(h->d_Sblk);
    h->d_blk_a = h->d_blk_b = h->d_blk_start = nullptr; h->d_Sblk = nullptr; h->cap_blocks = 0;
    CMOS_REQUIRE(alloc(&h->d_blk_a, want) && alloc(&h->d_blk_b, want) && alloc(&h->d_blk_start, want + 1) &&
                 alloc(&h->d_Sblk, want * 36 + 6 * (size_t)h->p.max_cams + 8),
                 "cannot grow the reduced-system block list to %zu blocks", want);
    h->cap_blocks = want;
  }
  std::vector<int> pair_a(std::max<size_t>(n_pairs, 1)), pair_b(std::max<size_t>(n_pairs, 1));
  {
    std::vector<int> f3(blk_start.begin(), blk_start.end() - 1);
    for (int j = 0; j < M; j++)
      for (int pa = pt_start[j]; pa < pt_start[j + 1]; pa++) {
        if (o_cv[pa] < 0) continue;
        for (int pb = pt_start[j]; pb < pt_start[j + 1]; pb++) {
          if (o_cv[pb] < 0 || o_cv[pa] > o_cv[pb]) continue;
          const int id = table[(size_t)o_cv[pa] * Kv + o_cv[pb]] - 1;
          pair_a[f3[id]] = pa; pair_b[f3[id]] = pb; f3[id]++;
        }
      }
  }
  // ---- upload: one page-locked buffer, one copy, one scatter kernel ----
  cudaStream_t st = h->stream;
  BaDev& d = h->d;
  Stager sg(h);
  sg.add(h->d_cams0, cams, 7 * (size_t)K * sizeof(double));
  sg.add(h->d_pts0, points, 3 * (size_t)M * sizeof(double));
  sg.add(h->d_cam_var, cam_var.data(), K * sizeof(int));
  sg.add(h->d_var_cam, var_cam.data(), var_cam.size() * sizeof(int));
  sg.add(h->d_cam_flags, cam_flags, K);
  sg.add(h->d_o_cam, o_cam.data(), N * sizeof(int));
  sg.add(h->d_o_cv, o_cv.data(), N * sizeof(int));
  sg.add(h->d_o_pt, o_pt.data(), N * sizeof(int));
  sg.add(h->d_o_uv, o_uv.data(), N * sizeof(float2));
  sg.add(h->d_o_w, o_w.data(), N * sizeof(float));
  sg.add(h->d_perm, perm.data(), N * sizeof(int));
  sg.add(h->d_pt_start, pt_start.data(), (M + 1) * sizeof(int));
  sg.add(h->d_cam_start, cam_start.data(), (Kv + 1) * sizeof(int));
  sg.add(h->d_cam_obs, cam_obs.data(), cam_obs.size() * sizeof(int));
  sg.add(h->d_blk_a, blk_a.data(), nb * sizeof(int));
  sg.add(h->d_blk_b, blk_b.data(), nb * sizeof(int));
  sg.add(h->d_blk_start, blk_start.data(), (nb + 1) * sizeof(int));
  sg.add(h->d_pair_a, pair_a.data(), pair_a.size() * sizeof(int));
  sg.add(h->d_pair_b, pair_b.data(), pair_b.size() * sizeof(int));
  d.K = K; d.Kv = Kv; d.M = M; d.N = N; d.n_blocks = nb; d.nc = 6 * Kv;
  d.fx = (double)K4[0]; d.fy = (double)K4[1]; d.cx = (double)K4[2]; d.cy = (double)K4[3];
  d.cam_var = h->d_cam_var; d.o_cam = h->d_o_cam; d.o_cv = h->d_o_cv; d.o_pt = h->d_o_pt; d.o_uv = h->d_o_uv;
  d.o_w = h->d_o_w; d.o_mode = h->d_o_mode; d.pt_start = h->d_pt_start; d.cam_start = h->d_cam_start;
  d.cam_obs = h->d_cam_obs; d.blk_a = h->d_blk_a; d.blk_b = h->d_blk_b; d.blk_start = h->d_blk_start;
  d.pair_a = h->d_pair_a; d.pair_b = h->d_pair_b;
  d.var_cam = h->d_var_cam; d.red = h->d_red;
  d.ticket = (unsigned int*)(h->d_red + 8);
  d.Hcc = h->d_HG; d.gc = h->d_HG + 21 * (size_t)Kv;
  d.Sblk = h->d_Sblk; d.rhs = h->d_Sblk + (size_t)nb * 36;
  {
    // chunks of <= chunk_pairs pairs, never across blocks.  One round (32 pairs) per warp unless that makes more than
    // ~64 k warps; CMOS_BA_SCHUR=cta selects the CTA-per-block kernel (A/B runs).
    const char* force = std::getenv("CMOS_BA_SCHUR");
    h->schur_chunked = !(force && force[0] == 'c');
    int chunk_pairs = 32;
    while (n_pairs / chunk_pairs > 65536) chunk_pairs *= 2;
    std::vector<int> chunk_start, chunk_blk, blk_chunk0(nb + 1, 0);
    for (int i = 0; i < nb; i++) {
      blk_chunk0[i] = (int)chunk_blk.size();
      for (int e = blk_start[i]; e < blk_start[i + 1]; e += chunk_pairs) { chunk_start.push_back(e); chunk_blk.push_back(i); }
    }
    blk_chunk0[nb] = (int)chunk_blk.size();
    chunk_start.push_back(nb > 0 ? blk_start[nb] : 0);
    const size_t nch = chunk_blk.size();
    if (nch > h->cap_chunks || (size_t)nb > h->cap_chunk_blocks) {
      CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
      cudaFree(h->d_chunk_start); cudaFree(h->d_chunk_blk); cudaFree(h->d_blk_chunk0); cudaFree(h->d_schur_part);
      h->d_chunk_start = h->d_chunk_blk = h->d_blk_chunk0 = nullptr; h->d_schur_part = nullptr;
      h->cap_chunks = h->cap_chunk_blocks = 0;
      const size_t wc = nch + nch / 4 + 16, wb = (size_t)nb + nb / 4 + 16;
      CMOS_REQUIRE(alloc(&h->d_chunk_start, wc + 1) && alloc(&h->d_chunk_blk, wc) && alloc(&h->d_blk_chunk0, wb + 1) &&
                   alloc(&h->d_schur_part, wc * kSchurPart), "cannot allocate the Schur chunk lists (%zu chunks)", nch);
      h->cap_chunks = wc; h->cap_chunk_blocks = wb;
    }
    sg.add(h->d_chunk_start, chunk_start.data(), chunk_start.size() * sizeof(int));
    if (nch) sg.add(h->d_chunk_blk, chunk_blk.data(), nch * sizeof(int));
    sg.add(h->d_blk_chunk0, blk_chunk0.data(), blk_chunk0.size() * sizeof(int));
    {
      const int rc = sg.flush(st);                   // copies every source into the page-locked buffer: the host vectors may die
      if (rc) return rc;
    }
    d.chunk_start = h->d_chunk_start; d.chunk_blk = h->d_chunk_blk; d.blk_chunk0 = h->d_blk_chunk0;
    d.schur_part = h->d_schur_part; d.n_chunks = (int)nch;
  }
  d.multi = h->n_ranks > 1; d.is_root = h->rank == 0;
  d.rank = h->rank; d.n_ranks = h->n_ranks; d.lin_tail = h->d_HG + 27 * (size_t)Kv;
  const int nlb = (int)(((size_t)M * kPointLanes + kLinThreads - 1) / kLinThreads);
  d.n_lin_blocks = nlb;
  int o = 0;
  d.o_lin_cost = o; o += nlb; d.o_lin_gmax = o; o += nlb; d.o_lin_xn2 = o; o += nlb;
  d.o_bs_cost = o; o += nlb; d.o_bs_mcc = o; o += nlb; d.o_bs_sn2 = o; o += nlb;
  d.o_cam_gmax = o; o += Kv; d.o_cam_xn2 = o; o += Kv; d.o_cam_mcc = o; o += Kv; d.o_cam_sn2 = o; o += Kv;
  d.trace = nullptr; d.stop_flag = nullptr;
  CMOS_CUDA_OK(cudaMemsetAsync(d.part, 0, (size_t)o * sizeof(double), st));
  CMOS_CUDA_OK(cudaMemsetAsync(d.red, 0, 8 * sizeof(double), st));
  h->topo_flags.assign(cam_flags, cam_flags + K);
  h->topo_obs_cam.assign(obs_cam, obs_cam + N);
  h->topo_obs_pt.assign(obs_pt, obs_pt + N);
  h->topo_perm = perm;
  h->topo_valid = true;
  h->has_problem = true;
  h->ran = false;
  return CMOS_OK;
}

static int restore_and_prepare(cmos_ba* h, cudaStream_t st) {
  BaDev& d = h->d;
  for (int b = 0; b < 2; b++) {
    CMOS_CUDA_OK(cudaMemcpyAsync(d.cams[b], h->d_cams0, 7 * (size_t)d.K * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CMOS_CUDA_OK(cudaMemcpyAsync(d.pts[b], h->d_pts0, 3 * (size_t)d.M * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  CMOS_CUDA_OK(cudaMemsetAsync(d.st, 0, sizeof(LmState), st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_trace, 0, 2 * (size_t)h->trace_rows * kTraceCols * sizeof(double), st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_erase, 0, d.N, st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_summaries, 0, 2 * sizeof(cmos_ba_summary), st));
  h->launches = 0;
  return CMOS_OK;
}

int cmos_ba_run_local(cmos_ba_t h, int32_t iterations_pass0, int32_t iterations_pass1, const uint8_t* stop_flag,
                      void* stream) {
  CMOS_REQUIRE(h, "null handle");
  if (!h->has_problem) { set_error("cmos_ba_set_problem must be called first"); return CMOS_ERR_STATE; }
  CMOS_REQUIRE(iterations_pass0 >= 0 && iterations_pass1 >= 0 && iterations_pass0 + 1 < h->trace_rows &&
               iterations_pass1 + 1 < h->trace_rows, "bad iteration counts");
  CMOS_REQUIRE(h->n_ranks == 1, "LocalBundleAdjustment runs on one GPU (replicas only, SURVEY.md 8e)");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  int rc = map_stop_flag(h, stop_flag, &h->d.stop_flag);
  if (rc) return rc;
  if ((rc = restore_and_prepare(h, st))) return rc;
  h->ran = true;
  BaDev& d = h->d;
  const int gN = (d.N + 255) / 256;
  h->timer.begin(st);
  launch_chain(k_set_mode, gN, 256, 0, st, d, 1);
  if ((rc = enqueue_solve(h, iterations_pass0, 0, st))) return rc;
  launch_chain(k_outlier_scan, gN, 256, 0, st, d, h->d_cam_flags, h->d_perm, h->d_erase, 1);
  if ((rc = enqueue_solve(h, iterations_pass1, 1, st))) return rc;
  launch_chain(k_outlier_scan, gN, 256, 0, st, d, h->d_cam_flags, h->d_perm, h->d_erase, 0);
  launch_chain(k_gather_result, (std::max(7 * d.K, 3 * d.M) + 255) / 256, 256, 0, st, d, h->d_cams0, h->d_pts0, h->d_cams_out, h->d_pts_out);
  h->timer.mark(st);
  h->launches += 4;
  CMOS_CUDA_OK(cudaGetLastError());
  return CMOS_OK;
}

int cmos_ba_run_global(cmos_ba_t h, int32_t n_iterations, int32_t robust, const uint8_t* stop_flag, void* stream) {
  CMOS_REQUIRE(h, "null handle");
  if (!h->has_problem) { set_error("cmos_ba_set_problem must be called first"); return CMOS_ERR_STATE; }
  CMOS_REQUIRE(n_iterations >= 0 && n_iterations + 1 < h->trace_rows, "bad iteration count");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  int rc = map_stop_flag(h, stop_flag, &h->d.stop_flag);
  if (rc) return rc;
  if ((rc = restore_and_prepare(h, st))) return rc;
  h->ran = true;
  BaDev& d = h->d;
  h->timer.begin(st);
  launch_chain(k_set_mode, (d.N + 255) / 256, 256, 0, st, d, robust ? 1 : 2);
  if ((rc = enqueue_solve(h, n_iterations, 0, st))) return rc;
  launch_chain(k_gather_result, (std::max(7 * d.K, 3 * d.M) + 255) / 256, 256, 0, st, d, h->d_cams0, h->d_pts0, h->d_cams_out, h->d_pts_out);
  h->timer.mark(st);
  h->launches += 2;
  CMOS_CUDA_OK(cudaGetLastError());
  return CMOS_OK;
}

int cmos_ba_get_results(cmos_ba_t h, double* cams, double* points, uint8_t* erase, cmos_ba_summary* summaries,
                        void* stream) {
  CMOS_REQUIRE(h, "null handle");
  if (!h->ran) { set_error("no solve has run"); return CMOS_ERR_STATE; }
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const BaDev& d = h->d;
  {
    // one gather kernel + one copy into the page-locked buffer instead of four pageable downloads.  The staging buffer is
    // shared with set_problem's upload on the handle's own stream: wait for that stream first when another one is used.
    if (st != h->stream) CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
    Stager sg(h);
    if (cams) sg.add(h->d_cams_out, cams, 7 * (size_t)d.K * sizeof(double));
    if (points) sg.add(h->d_pts_out, points, 3 * (size_t)d.M * sizeof(double));
    if (erase) sg.add(h->d_erase, erase, d.N);
    if (summaries) sg.add(h->d_summaries, summaries, 2 * sizeof(cmos_ba_summary));
    const int rc = sg.download(st);
    if (rc) return rc;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
  }
  // The caller's stop-flag page is page-locked only while a solve can read it: a stale registration would make later
  // copies from unrelated heap memory that shares the page fail ("invalid argument": a partly pinned source range).
  if (h->stop_registered_by_us && h->stop_host_page) {
    cudaHostUnregister((void*)h->stop_host_page);
    h->stop_registered_by_us = false;
    h->stop_host_page = nullptr;
    h->stop_dev_page = nullptr;
  }
  return CMOS_OK;
}

int cmos_ba_local_bundle_adjustment(cmos_ba_t h, int32_t n_cams, double* cams, const uint8_t* cam_flags, int32_t n_points,
                                    double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                                    const float* uv, const float* inv_sigma2, const float* K4, const uint8_t* stop_flag,
                                    uint8_t* erase, cmos_ba_summary* summaries) {
  int rc = cmos_ba_set_problem(h, n_cams, cams, cam_flags, n_points, points, n_obs, obs_cam, obs_pt, uv, inv_sigma2, K4);
  if (rc) return rc;
  if ((rc = cmos_ba_run_local(h, 5, 10, stop_flag, nullptr))) return rc;   // CeresOptimizer.cc:517-519
  return cmos_ba_get_results(h, cams, points, erase, summaries, nullptr);
}

int cmos_ba_bundle_adjustment(cmos_ba_t h, int32_t n_cams, double* cams, const uint8_t* cam_const, int32_t n_points,
                              double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                              const float* uv, const float* inv_sigma2, const float* K4, int32_t n_iterations,
                              int32_t robust, const uint8_t* stop_flag, cmos_ba_summary* summary) {
  int rc = cmos_ba_set_problem(h, n_cams, cams, cam_const, n_points, points, n_obs, obs_cam, obs_pt, uv, inv_sigma2, K4);
  if (rc) return rc;
  if ((rc = cmos_ba_run_global(h, n_iterations, robust, stop_flag, nullptr))) return rc;
  cmos_ba_summary s2[2];
  rc = cmos_ba_get_results(h, cams, points, nullptr, s2, nullptr);
  if (summary) *summary = s2[0];
  return rc;
}

int cmos_ba_comm_unique_id(uint8_t* id128) {
  CMOS_REQUIRE(id128, "null argument");
  if (!g_nccl.load()) { set_error("NCCL (libnccl.so.2) is not available: %s", dlerror()); return CMOS_ERR_STATE; }
  ncclUniqueId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) { set_error("ncclGetUniqueId failed: %s", g_nccl.GetErrorString(rc)); return CMOS_ERR_CUDA; }
  std::memcpy(id128, id.internal, 128);
  return CMOS_OK;
}

int cmos_ba_comm_init(cmos_ba_t h, const uint8_t* id128, int32_t n_ranks, int32_t rank) {
  CMOS_REQUIRE(h && id128 && n_ranks >= 1 && n_ranks <= 64 && rank >= 0 && rank < n_ranks, "bad argument (1 <= n_ranks <= 64)");
  if (!g_nccl.load()) { set_error("NCCL (libnccl.so.2) is not available: %s", dlerror()); return CMOS_ERR_STATE; }
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  const int rc = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
  if (rc != 0) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc)); h->comm = nullptr; return CMOS_ERR_CUDA; }
  h->n_ranks = n_ranks;
  h->rank = rank;
  h->has_problem = false;
  return CMOS_OK;
}

int cmos_ba_debug_trace(cmos_ba_t h, int32_t pass, double* trace, int32_t rows) {
  CMOS_REQUIRE(h && trace && (pass == 0 || pass == 1) && rows >= 1 && rows <= h->trace_rows, "bad argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  CMOS_CUDA_OK(cudaDeviceSynchronize());
  CMOS_CUDA_OK(cudaMemcpy(trace, h->d_trace + (size_t)pass * h->trace_rows * kTraceCols,
                          (size_t)rows * kTraceCols * sizeof(double), cudaMemcpyDeviceToHost));
  return CMOS_OK;
}

int cmos_ba_debug_stop_at(cmos_ba_t h, int32_t pass, int32_t iteration) {
  CMOS_REQUIRE(h, "null handle");
  h->dbg_stop_pass = pass; h->dbg_stop_iter = iteration;
  return CMOS_OK;
}

int cmos_ba_last_launch_count(cmos_ba_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

int cmos_ba_set_profiling(cmos_ba_t h, int32_t enable) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  h->timer.reset();
  h->timer.enabled = enable != 0;
  return CMOS_OK;
}

int cmos_ba_solve_time(cmos_ba_t h, double* ms, int64_t* calls) {
  CMOS_REQUIRE(h && ms && calls, "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  h->timer.fold();
  *ms = h->timer.total_ms[0];
  *calls = h->timer.calls;
  return CMOS_OK;
}

int cmos_ba_optimize_sim3(cmos_ba_t h, int32_t n, double* s12, double* R12, double* t12, const float* K1, const float* K2,
                          const float* obs1, const float* inv_sigma1, const double* P3D2c, const float* obs2,
                          const float* inv_sigma2, const double* P3D1c, float th2, int32_t max_iterations,
                          uint8_t* is_bad, double* lie7, int32_t* n_inliers, cmos_ba_summary* summary) {
  CMOS_REQUIRE(h && s12 && R12 && t12 && K1 && K2 && n_inliers, "null argument");
  CMOS_REQUIRE(n >= 0 && (n == 0 || (obs1 && inv_sigma1 && P3D2c && obs2 && inv_sigma2 && P3D1c && is_bad)), "null argument");
  CMOS_REQUIRE(max_iterations >= 0 && max_iterations + 1 < h->trace_rows, "max_iterations %d too large", max_iterations);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  const size_t cap = std::max<size_t>(n, 1);
  if (cap > h->sim3_cap) {
    for (void* b : {(void*)h->ds_obs, (void*)h->ds_sig, (void*)h->ds_pts, (void*)h->ds_out, (void*)h->ds_bad})
      if (b) cudaFree(b);
    h->ds_obs = nullptr; h->ds_sig = nullptr; h->ds_pts = nullptr; h->ds_out = nullptr; h->ds_bad = nullptr; h->sim3_cap = 0;
    if (!alloc(&h->ds_obs, 4 * cap) || !alloc(&h->ds_sig, 2 * cap) || !alloc(&h->ds_pts, 6 * cap) || !alloc(&h->ds_out, 24) ||
        !alloc(&h->ds_bad, cap)) {
      set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      return CMOS_ERR_CUDA;
    }
    h->sim3_cap = cap;
  }
  const size_t c = h->sim3_cap;
  auto up = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
  CMOS_CUDA_OK(up(h->ds_obs, obs1, (size_t)n * 2 * sizeof(float)));
  CMOS_CUDA_OK(up(h->ds_obs + 2 * c, obs2, (size_t)n * 2 * sizeof(float)));
  CMOS_CUDA_OK(up(h->ds_sig, inv_sigma1, (size_t)n * sizeof(float)));
  CMOS_CUDA_OK(up(h->ds_sig + c, inv_sigma2, (size_t)n * sizeof(float)));
  CMOS_CUDA_OK(up(h->ds_pts, P3D2c, (size_t)n * 3 * sizeof(double)));
  CMOS_CUDA_OK(up(h->ds_pts + 3 * c, P3D1c, (size_t)n * 3 * sizeof(double)));
  Sim3Args a{};
  a.n = n; a.s12 = *s12;
  for (int i = 0; i < 9; i++) a.R12[i] = R12[i];
  for (int i = 0; i < 3; i++) a.t12[i] = t12[i];
  for (int i = 0; i < 4; i++) { a.K1[i] = (double)K1[i]; a.K2[i] = (double)K2[i]; }
  a.obs1 = h->ds_obs; a.obs2 = h->ds_obs + 2 * c; a.inv_sigma1 = h->ds_sig; a.inv_sigma2 = h->ds_sig + c;
  a.P3D2c = h->ds_pts; a.P3D1c = h->ds_pts + 3 * c;
  a.huber_a = std::sqrt((double)th2);
  a.max_iterations = max_iterations;
  a.is_bad = h->ds_bad; a.out = h->ds_out; a.summary = h->dp_sum; a.trace = h->dp_trace;
  launch_chain(k_sim3_opt, 1, kPoseThreads, 0, st, a);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  double out[24];
  CMOS_CUDA_OK(cudaMemcpyAsync(out, h->ds_out, sizeof(out), cudaMemcpyDeviceToHost, st));
  if (n) CMOS_CUDA_OK(cudaMemcpyAsync(is_bad, h->ds_bad, n, cudaMemcpyDeviceToHost, st));
  if (summary) CMOS_CUDA_OK(cudaMemcpyAsync(summary, h->dp_sum, sizeof(cmos_ba_summary), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  if (lie7) for (int i = 0; i < 7; i++) lie7[i] = out[i];
  *s12 = out[7];
  for (int i = 0; i < 9; i++) R12[i] = out[8 + i];
  for (int i = 0; i < 3; i++) t12[i] = out[17 + i];
  *n_inliers = (int)out[20];
  return CMOS_OK;
}

}  // extern "C"

namespace {
// Band + border plan of a pose graph over Kv variable keyframes (edges between variable keyframes vi != vj).  A keyframe
// whose edges reach further than Wc keyframes away goes to the border (greedy cover of the long edges: loop-closure edges
// tie a handful of keyframes to far-away ones); the rest, in keyframe order, is banded with half-width <= Wc and is cut into
// nodes of Wc keyframes (7 Wc unknowns padded to a multiple of 24).  ok = false: no such structure within the solver's
// limits (node <= 144 unknowns, border <= 144 unknowns) -> the blocked Cholesky takes the system.
struct EgPlan { bool ok = false; int Wb = 0, n = 0, N = 0, levels = 0, nb = 0, nbp = 0, Ki = 0; std::vector<int> pos; };
EgPlan plan_eg_bcr(int Kv, const std::vector<std::pair<int, int>>& edges) {
  EgPlan best;
  if (Kv < 1) return best;
  static const int kWc[] = {3, 6, 10, 13, 17, 20};
  for (int pass = 0; pass < 2 && !best.ok; pass++) {          // pass 0: a border of <= 10 keyframes; pass 1: <= 20
    const int nb_max = pass == 0 ? 10 : 20;
    for (int Wc : kWc) {
      std::vector<char> border(Kv, 0);
      int nb = 0;
      bool fits = true;
      for (;;) {
        std::vector<int> cnt(Kv, 0);
        int any = 0;
        for (const auto& e : edges)
          if (!border[e.first] && !border[e.second] && std::abs(e.first - e.second) > Wc) { cnt[e.first]++; cnt[e.second]++; any++; }
        if (!any) break;
        int pick = 0;
        for (int k = 1; k < Kv; k++) if (cnt[k] >= cnt[pick]) pick = k;     // most long edges, ties -> the newer keyframe
        border[pick] = 1;
        if (++nb > nb_max) { fits = false; break; }
      }
      if (!fits || nb >= Kv) continue;
      EgPlan p;
      p.ok = true; p.Wb = Wc; p.n = (7 * Wc + 23) / 24 * 24; p.nb = nb; p.nbp = (7 * nb + 23) / 24 * 24; p.Ki = Kv - nb;
      p.N = (p.Ki + Wc - 1) / Wc;
      p.levels = 1;
      while ((1 << p.levels) <= p.N) p.levels++;
      p.pos.resize(Kv);
      for (int k = 0, pi = 0, bi = 0; k < Kv; k++) p.pos[k] = border[k] ? -1 - bi++ : pi++;
      best = p;
      break;
    }
  }
  return best;
}
}  // namespace

extern "C" {

int cmos_ba_debug_essential_graph_plan(cmos_ba_t h, int32_t* info6) {
  CMOS_REQUIRE(h && info6, "null argument");
  for (int i = 0; i < 6; i++) info6[i] = h->eg.plan_info[i];
  return CMOS_OK;
}

int cmos_ba_optimize_essential_graph(cmos_ba_t h, int32_t n_kf, const double* Scw, const uint8_t* kf_flags, const double* Snc,
                                     int32_t n_edges, const int32_t* edge_j, const int32_t* edge_i, const uint8_t* edge_kind,
                                     int32_t max_iterations, int32_t n_points, const double* Xw, const int32_t* ref_kf,
                                     double* lie_out, double* Tiw_out, double* Xw_out, cmos_ba_summary* summary) {
  CMOS_REQUIRE(h && Scw && kf_flags && Snc && n_kf > 0, "null argument");
  CMOS_REQUIRE(n_edges >= 0 && (n_edges == 0 || (edge_j && edge_i && edge_kind)), "null edge arrays");
  CMOS_REQUIRE(n_points >= 0 && (n_points == 0 || (Xw && ref_kf && Xw_out)), "null point arrays");
  CMOS_REQUIRE(max_iterations >= 0 && max_iterations + 1 < h->trace_rows, "max_iterations %d too large", max_iterations);
  for (int e = 0; e < n_edges; e++)
    CMOS_REQUIRE(edge_j[e] >= 0 && edge_j[e] < n_kf && edge_i[e] >= 0 && edge_i[e] < n_kf, "edge %d references keyframe out of range", e);
  for (int p = 0; p < n_points; p++) CMOS_REQUIRE(ref_kf[p] >= 0 && ref_kf[p] < n_kf, "point %d references keyframe out of range", p);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  auto& g = h->eg;
  // ---- structure: variable keyframes, incidence CSR, row envelope -----------------------------------------------------
  std::vector<int> var(n_kf, -1), var_kf;
  for (int k = 0; k < n_kf; k++) if (!(kf_flags[k] & 1)) { var[k] = (int)var_kf.size(); var_kf.push_back(k); }
  const int Kv = (int)var_kf.size(), n = 7 * Kv;
  std::vector<int> inc_start(Kv + 1, 0);
  for (int e = 0; e < n_edges; e++) {
    if (var[edge_i[e]] >= 0) inc_start[var[edge_i[e]] + 1]++;
    if (var[edge_j[e]] >= 0) inc_start[var[edge_j[e]] + 1]++;
  }
  for (int a = 0; a < Kv; a++) inc_start[a + 1] += inc_start[a];
  const int n_inc = inc_start[Kv];
  std::vector<int> inc_edge(std::max(n_inc, 1)), inc_other(std::max(n_inc, 1)), inc_sign(std::max(n_inc, 1)), fill(inc_start.begin(), inc_start.end() - 1);
  std::vector<int> first_blk(Kv);
  for (int a = 0; a < Kv; a++) first_blk[a] = a;
  for (int e = 0; e < n_edges; e++) {       // edge order is kept inside every keyframe's list: the sums follow the reference's insertion order
    const int vi = var[edge_i[e]], vj = var[edge_j[e]];
    if (vi >= 0) { const int q = fill[vi]++; inc_edge[q] = e; inc_other[q] = vj; inc_sign[q] = 1; }
    if (vj >= 0) { const int q = fill[vj]++; inc_edge[q] = e; inc_other[q] = vi; inc_sign[q] = -1; }
    if (vi >= 0 && vj >= 0) { first_blk[std::max(vi, vj)] = std::min(first_blk[std::max(vi, vj)], std::min(vi, vj)); }
  }
  // nested dissection (band + border) when the graph has that structure; CMOS_EG_BLOCKED=1 forces the blocked Cholesky
  EgPlan plan;
  {
    std::vector<std::pair<int, int>> vv;
    for (int e = 0; e < n_edges; e++) {
      const int vi = var[edge_i[e]], vj = var[edge_j[e]];
      if (vi >= 0 && vj >= 0 && vi != vj) vv.emplace_back(vi, vj);
    }
    const char* blocked = std::getenv("CMOS_EG_BLOCKED");
    if (!(blocked && blocked[0] == '1')) plan = plan_eg_bcr(Kv, vv);
  }
  std::vector<int> tiles, pan_start(1, 0), pan_first_col;
  for (int k0 = 0; k0 < n && !plan.ok; k0 += kNB) {
    const int kb = std::min(kNB, n - k0), t0 = k0 + kb;
    int fc = k0;
    for (int r = k0; r < k0 + kb; r++) fc = std::min(fc, 7 * first_blk[r / 7]);
    pan_first_col.push_back(fc);
    for (int t = 0; t0 + 64 * t <= n; t++) {
      const int r0 = t0 + 64 * t, r1 = std::min(r0 + 63, n);
      bool active = r1 == n;                                   // the rhs row
      for (int r = r0; r <= std::min(r1, n - 1) && !active; r++) active = 7 * first_blk[r / 7] < t0;
      if (active) tiles.push_back(t);
    }
    pan_start.push_back((int)tiles.size());
  }
  // ---- storage --------------------------------------------------------------------------------------------------------
  const size_t n_panels = (n + kNB - 1) / kNB;
  // node arrays of the nested-dissection solve live in the allocation of the dense S (+ rhs and solution by node)
  const size_t bcr_doubles = plan.ok ? cr_doubles(plan.N, plan.n) + crb_doubles(plan.N, plan.n, plan.nbp) + 2 * (size_t)plan.N * plan.n : 0;
  if ((size_t)n_kf > g.cap_kf || (size_t)n_edges > g.cap_e || (size_t)n > g.cap_n || (size_t)n_points > g.cap_pts || tiles.size() > g.cap_tiles ||
      bcr_doubles > g.cap_S) {
    const size_t ck = std::max<size_t>(n_kf, g.cap_kf), ce = std::max<size_t>(std::max(n_edges, 1), g.cap_e), cn = std::max<size_t>(std::max(n, 7), g.cap_n),
                 cp = std::max<size_t>(std::max(n_points, 1), g.cap_pts), ct = std::max<size_t>(std::max<size_t>(tiles.size(), 1), g.cap_tiles);
    const size_t cS = std::max(std::max(cn * cn, bcr_doubles), g.cap_S);
    g.release();
    const size_t cpan = (cn + kNB - 1) / kNB;
    bool ok = alloc(&g.x0, 7 * ck) && alloc(&g.x1, 7 * ck) && alloc(&g.x_init, 7 * ck) && alloc(&g.Scw, 13 * ck) && alloc(&g.Snc, 13 * ck) &&
              alloc(&g.lie_out, 7 * ck) && alloc(&g.Tiw, 16 * ck) && alloc(&g.flags, ck) && alloc(&g.kind, ce) && alloc(&g.var, ck) &&
              alloc(&g.var_kf, ck) && alloc(&g.ej, ce) && alloc(&g.ei, ce) && alloc(&g.inc_start, ck + 1) && alloc(&g.inc_edge, 2 * ce) &&
              alloc(&g.inc_other, 2 * ce) && alloc(&g.inc_sign, 2 * ce) && alloc(&g.tiles, ct) && alloc(&g.first_col, cpan + 1) && alloc(&g.pos, ck) && alloc(&g.ref, cp) && alloc(&g.meas, ce) &&
              alloc(&g.Swc, ck) && alloc(&g.r, 7 * ce) && alloc(&g.J, 49 * ce) && alloc(&g.A, 49 * ce) && alloc(&g.v, 7 * ce) &&
              alloc(&g.Hd, 49 * ck) && alloc(&g.g, cn) && alloc(&g.scale, cn) && alloc(&g.delta, cn) && alloc(&g.S, cS) &&
              alloc(&g.rhs, cn) && alloc(&g.yc, cn) && alloc(&g.Linv, cpan * kNB * kNB) && alloc(&g.pe, 3 * ce) && alloc(&g.pk, 3 * ck) &&
              alloc(&g.Xw, 3 * cp) && alloc(&g.Xo, 3 * cp) && alloc(&g.st, 1) &&
              cudaMallocHost((void**)&g.done_host, sizeof(int)) == cudaSuccess;
    if (!ok) {
      set_error("device allocation failed for a %d-keyframe essential graph: %s", n_kf, cudaGetErrorString(cudaGetLastError()));
      g.release();
      return CMOS_ERR_CUDA;
    }
    g.cap_kf = ck; g.cap_e = ce; g.cap_n = cn; g.cap_pts = cp; g.cap_tiles = ct; g.cap_S = cS;
  }
  auto up = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
  CMOS_CUDA_OK(up(g.Scw, Scw, (size_t)n_kf * 13 * sizeof(double)));
  CMOS_CUDA_OK(up(g.Snc, Snc, (size_t)n_kf * 13 * sizeof(double)));
  CMOS_CUDA_OK(up(g.flags, kf_flags, n_kf));
  CMOS_CUDA_OK(up(g.var, var.data(), (size_t)n_kf * sizeof(int)));
  CMOS_CUDA_OK(up(g.var_kf, var_kf.data(), (size_t)Kv * sizeof(int)));
  CMOS_CUDA_OK(up(g.ej, edge_j, (size_t)n_edges * sizeof(int)));
  CMOS_CUDA_OK(up(g.ei, edge_i, (size_t)n_edges * sizeof(int)));
  CMOS_CUDA_OK(up(g.kind, edge_kind, n_edges));
  CMOS_CUDA_OK(up(g.inc_start, inc_start.data(), (size_t)(Kv + 1) * sizeof(int)));
  CMOS_CUDA_OK(up(g.inc_edge, inc_edge.data(), (size_t)n_inc * sizeof(int)));
  CMOS_CUDA_OK(up(g.inc_other, inc_other.data(), (size_t)n_inc * sizeof(int)));
  CMOS_CUDA_OK(up(g.inc_sign, inc_sign.data(), (size_t)n_inc * sizeof(int)));
  CMOS_CUDA_OK(up(g.tiles, tiles.data(), tiles.size() * sizeof(int)));
  CMOS_CUDA_OK(up(g.first_col, pan_first_col.data(), pan_first_col.size() * sizeof(int)));
  if (plan.ok) CMOS_CUDA_OK(up(g.pos, plan.pos.data(), (size_t)Kv * sizeof(int)));
  CMOS_CUDA_OK(up(g.Xw, Xw, (size_t)n_points * 3 * sizeof(double)));
  CMOS_CUDA_OK(up(g.ref, ref_kf, (size_t)n_points * sizeof(int)));
  EgDev d{};
  d.n_kf = n_kf; d.Kv = Kv; d.E = n_edges; d.n = n;
  d.x[0] = g.x0; d.x[1] = g.x1; d.x0 = g.x_init; d.Scw = g.Scw; d.Snc = g.Snc; d.kf_flags = g.flags; d.var = g.var; d.var_kf = g.var_kf;
  d.edge_j = g.ej; d.edge_i = g.ei; d.edge_kind = g.kind; d.meas = g.meas; d.r = g.r; d.J = g.J; d.A = g.A; d.v = g.v;
  d.inc_start = g.inc_start; d.inc_edge = g.inc_edge; d.inc_other = g.inc_other; d.inc_sign = g.inc_sign;
  d.Hd = g.Hd; d.g = g.g; d.scale = g.scale; d.delta = g.delta; d.S = g.S; d.rhs = g.rhs; d.yc = g.yc;
  d.p_cost = g.pe; d.p_mcc = g.pe + g.cap_e; d.p_cand = g.pe + 2 * g.cap_e;
  d.p_gmax = g.pk; d.p_xn2 = g.pk + g.cap_kf; d.p_sn2 = g.pk + 2 * g.cap_kf;
  d.st = g.st; d.trace = h->dp_trace;
  BaDev dv{};                       // the view the blocked Cholesky kernels read: S, rhs, yc, nc, st
  dv.S = g.S; dv.rhs = g.rhs; dv.yc = g.yc; dv.nc = n; dv.st = g.st;
  CrArgs ca{};
  g.plan_info[0] = plan.ok; g.plan_info[1] = plan.Wb; g.plan_info[2] = plan.n; g.plan_info[3] = plan.N; g.plan_info[4] = plan.nb; g.plan_info[5] = plan.nbp;
  if (plan.ok) {
    ca.n = plan.n; ca.Wb = plan.Wb; ca.N = plan.N; ca.W = plan.Wb; ca.levels = plan.levels; ca.band_blk = nullptr; ca.base = g.S;
    ca.nbp = plan.nbp;
    ca.bbase = g.S + cr_doubles(plan.N, plan.n);
    double* rn = ca.bbase + crb_doubles(plan.N, plan.n, plan.nbp);
    ca.rhs_nodes = rn; ca.x_nodes = rn + (size_t)plan.N * plan.n;
    d.bcr = 1; d.n_interior = plan.Ki; d.n_border = plan.nb; d.pos = g.pos; d.ca = ca;
  }
  const int eff_iterations = Kv > 0 ? max_iterations : 0;
  h->launches = 0;
  const int gk = (n_kf + 127) / 128, ge = std::max(1, (n_edges + 63) / 64), gw = std::max(1, (Kv + 3) / 4);
  launch_chain(k_eg_logs, gk, 128, 0, st, d, eff_iterations);
  launch_chain(k_eg_meas, std::max(1, (n_edges + 127) / 128), 128, 0, st, d);
  h->launches += 2;
  auto linearize = [&]() {
    launch_chain(k_eg_linearize, ge, 64, 0, st, d);
    launch_chain(k_eg_assemble, gw, 128, 0, st, d);
    launch_chain(k_eg_post_lin, 1, 256, 0, st, d);
    h->launches += 3;
  };
  for (int it = 0; it < eff_iterations; it++) {
    linearize();
    if (plan.ok) {
      // nested dissection: log2(N) levels of node factorisations (all nodes of a level in parallel), the border last
      launch_chain(k_eg_cr_clear, ca.N + 1, 256, 0, st, d);
      launch_chain(k_eg_build, gw, 128, 0, st, d);
      h->launches += 2;
      const int nt = ca.n / 24, ntb = ca.nbp / 24, kw = cr_spike_warps(ca.n);
      const int tile_ctas = (nt * nt + kCrGemmWarps - 1) / kCrGemmWarps;
      const int tile_ctas_b = (std::max(nt * ntb, ntb * ntb) + kCrGemmWarps - 1) / kCrGemmWarps;
      for (int l = 1; l <= ca.levels; l++) {
        const int cnt = ((ca.N >> (l - 1)) + 1) / 2;
        launch_chain(k_cr_factor, cnt, kSolveThreads, cr_factor_smem(ca.n), st, dv, ca, l);
        h->launches++;
        // neighbour and border spikes in one launch, neighbour and border products in one launch; the last level has no neighbours
        const int sx = std::max(nt / kw, (ntb + kw - 1) / kw);
        if (l < ca.levels) launch_chain(k_cr_spike, dim3(ntb ? sx : nt / kw, ntb ? 3 : 2, cnt), 32 * kw, cr_spike_smem(ca.n), st, dv, ca, l, 0);
        else if (ntb) launch_chain(k_cr_spike, dim3((ntb + kw - 1) / kw, 1, cnt), 32 * kw, cr_spike_smem(ca.n), st, dv, ca, l, 2);
        if (l < ca.levels && ntb) launch_chain(k_cr_schur_all, dim3(std::max(tile_ctas, tile_ctas_b), 8, cnt), 32 * kCrGemmWarps, 0, st, dv, ca, l);
        else if (l < ca.levels) launch_chain(k_cr_schur, dim3(tile_ctas, 4, cnt), 32 * kCrGemmWarps, 0, st, dv, ca, l);
        else if (ntb) launch_chain(k_crb_schur, dim3(tile_ctas_b, 4, cnt), 32 * kCrGemmWarps, 0, st, dv, ca, l);
        h->launches += (l < ca.levels || ntb) ? 2 : 0;
      }
      if (ntb) { launch_chain(k_crb_solve, 1, kSolveThreads, crb_solve_smem(ca.nbp), st, dv, ca); h->launches++; }
      for (int l = ca.levels; l >= 1; l--) {
        const int cnt = ((ca.N >> (l - 1)) + 1) / 2;
        launch_chain(k_cr_back, cnt, 32 * kCrBackWarps, cr_back_smem(ca.n), st, dv, ca, l);
        h->launches++;
      }
      launch_chain(k_eg_cr_scatter, (n + 255) / 256, 256, 0, st, d);
      h->launches++;
    } else {
    CMOS_CUDA_OK(cudaMemsetAsync(g.S, 0, (size_t)n * n * sizeof(double), st));
    launch_chain(k_eg_build, gw, 128, 0, st, d);
    h->launches++;
    for (int k0 = 0, p = 0; k0 < n; k0 += kNB, p++) {
      const int kb = std::min(kNB, n - k0);
      launch_chain(k_potrf_diag, 1, kPotrfThreads, kPanelSmem, st, dv, k0, kb, g.Linv);
      const int na = pan_start[p + 1] - pan_start[p];
      if (na > 0) {
        const int* tl = g.tiles + pan_start[p];
        launch_chain(k_trsm_panel, 2 * na, 256, kPanelSmem, st, dv, k0, kb, g.Linv, tl);
        launch_chain(k_syrk_tile, na * (na + 1) / 2, 256, kPanelSmem, st, dv, k0, kb, tl);
        h->launches += 2;
      }
      h->launches++;
    }
    if (use_back_all(n)) {
      launch_chain(k_backsolve_all, 1, 1024, back_all_smem(n), st, dv, g.Linv, g.first_col);
      h->launches++;
    } else
      for (int k0 = ((n - 1) / kNB) * kNB; k0 >= 0; k0 -= kNB) {
        const int kb = std::min(kNB, n - k0);
        const int c0 = std::min(pan_first_col[k0 / kNB], k0);
        launch_chain(k_backsolve_panel, std::max(1, (k0 - c0 + 255) / 256), 256, 0, st, dv, k0, kb, g.Linv, c0);
        h->launches++;
      }
    }
    launch_chain(k_eg_step, gk, 128, 0, st, d);
    launch_chain(k_eg_eval, ge, 64, 0, st, d);
    launch_chain(k_eg_decide, 1, 256, 0, st, d, g.done_host);
    h->launches += 3;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));         // one word: has the device-side state machine terminated?
    if (*g.done_host) break;
  }
  if (eff_iterations == 0) linearize();              // Ceres still evaluates iteration 0
  launch_chain(k_eg_summary, 1, 1, 0, st, d, h->dp_sum);
  launch_chain(k_eg_finish, gk, 128, 0, st, d, g.lie_out, g.Tiw, g.Swc);
  h->launches += 2;
  if (n_points > 0) { launch_chain(k_eg_points, (n_points + 255) / 256, 256, 0, st, d, n_points, g.Xw, g.ref, g.Swc, g.Xo); h->launches++; }
  CMOS_CUDA_OK(cudaGetLastError());
  if (lie_out) CMOS_CUDA_OK(cudaMemcpyAsync(lie_out, g.lie_out, (size_t)n_kf * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (Tiw_out) CMOS_CUDA_OK(cudaMemcpyAsync(Tiw_out, g.Tiw, (size_t)n_kf * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (n_points > 0) CMOS_CUDA_OK(cudaMemcpyAsync(Xw_out, g.Xo, (size_t)n_points * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (summary) CMOS_CUDA_OK(cudaMemcpyAsync(summary, h->dp_sum, sizeof(cmos_ba_summary), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

}  // extern "C"
