// Device-side math shared by the bundle-adjustment kernels: reprojection residual with analytic tangent
// Jacobians, Huber corrector, quaternion manifold, Levenberg-Marquardt state machine, deterministic block
// reductions.  All fp64.  Behaviour: include/CeresOptimizer.h:56-166 + the Ceres semantics of SURVEY.md A.5.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmos {

constexpr double kHuberA = 2.447651936039926;    // sqrt(5.991), CeresOptimizer.cc:81,296,421
constexpr double kHuberB = kHuberA * kHuberA;
constexpr double kChi2 = 5.991;                  // CeresOptimizer.cc:253,548

// Solver::Options defaults used by every ceres::Solve call of the reference
constexpr double kInitialRadius = 1e4, kMaxRadius = 1e16, kMinRadius = 1e-32;
constexpr double kMinLmDiag = 1e-6, kMaxLmDiag = 1e32;
constexpr double kMinRelativeDecrease = 1e-3;
constexpr double kFunctionTol = 1e-6, kParameterTol = 1e-8, kGradientTol = 1e-10;
constexpr int kMaxInvalidSteps = 5;

enum { TERM_MAX_ITER = 0, TERM_FUNCTION_TOL = 1, TERM_PARAMETER_TOL = 2, TERM_GRADIENT_TOL = 3, TERM_USER = 4,
       TERM_FAILURE = 5, TERM_MIN_RADIUS = 6 };

struct LmState {
  double radius, decrease_factor;
  double x_cost, x_norm, gmax, initial_cost;
  int iteration, max_iterations, successful, invalid_consec;
  int termination, done, need_lin, cur;      // cur: which of the two parameter buffers holds x
  int jac_evals, solve_failed, first, pad;
};

__device__ __forceinline__ void lm_init(LmState& s, int max_iterations, int cur) {
  s.radius = kInitialRadius; s.decrease_factor = 2.0;
  s.x_cost = 0; s.x_norm = 0; s.gmax = 0; s.initial_cost = 0;
  s.iteration = 0; s.max_iterations = max_iterations; s.successful = 0; s.invalid_consec = 0;
  s.termination = TERM_MAX_ITER; s.done = 0; s.need_lin = 1; s.cur = cur;
  s.jac_evals = 0; s.solve_failed = 0; s.first = 1; s.pad = 0;
}

// After EvaluateGradientAndJacobian: record cost / norms, test the gradient tolerance
// (TrustRegionMinimizer::IterationZero / HandleSuccessfulStep + FinalizeIteration...).
__device__ __forceinline__ void lm_after_linearize(LmState& s, double cost, double gmax, double x_norm, double* trace) {
  s.x_cost = cost; s.gmax = gmax; s.x_norm = x_norm;
  s.jac_evals++;
  if (s.first) { s.initial_cost = cost; s.first = 0; }
  s.need_lin = 0;
  if (trace) {
    double* t = trace + 8 * s.iteration;
    if (s.iteration == 0) { t[1] = 0; t[3] = 0; t[4] = 0; t[5] = s.radius; t[6] = 0; t[7] = 0; }
    t[0] = cost; t[2] = gmax;
  }
  if (s.iteration >= s.max_iterations) { s.done = 1; s.termination = TERM_MAX_ITER; return; }
  if (gmax <= kGradientTol) { s.done = 1; s.termination = TERM_GRADIENT_TOL; }
}

// One trust-region iteration's bookkeeping once the candidate has been evaluated.
__device__ __forceinline__ void lm_decide(LmState& s, bool solve_ok, double mcc, double cand_cost, double step_norm,
                                          double* trace) {
  s.iteration++;
  double* t = trace ? trace + 8 * s.iteration : nullptr;
  const bool valid = solve_ok && (mcc > 0.0);
  if (!valid) {   // HandleInvalidStep -> LevenbergMarquardtStrategy::StepIsInvalid
    s.invalid_consec++;
    s.radius = s.radius / s.decrease_factor; s.decrease_factor *= 2.0;
    if (t) { t[0] = s.x_cost; t[1] = 0; t[2] = s.gmax; t[3] = 0; t[4] = 0; t[5] = s.radius; t[6] = 0; t[7] = 0; }
    if (s.invalid_consec >= kMaxInvalidSteps) { s.done = 1; s.termination = TERM_FAILURE; }
    else if (s.iteration >= s.max_iterations) { s.done = 1; s.termination = TERM_MAX_ITER; }
    else if (s.radius < kMinRadius) { s.done = 1; s.termination = TERM_MIN_RADIUS; }
    return;
  }
  s.invalid_consec = 0;
  if (step_norm <= kParameterTol * (s.x_norm + kParameterTol)) {
    s.done = 1; s.termination = TERM_PARAMETER_TOL;
    if (t) { t[0] = s.x_cost; t[1] = 0; t[2] = s.gmax; t[3] = step_norm; t[4] = 0; t[5] = s.radius; t[6] = 0; t[7] = 1; }
    return;
  }
  const double cost_change = s.x_cost - cand_cost;
  if (fabs(cost_change) <= kFunctionTol * s.x_cost) {
    s.done = 1; s.termination = TERM_FUNCTION_TOL;
    if (t) { t[0] = s.x_cost; t[1] = cost_change; t[2] = s.gmax; t[3] = step_norm; t[4] = 0; t[5] = s.radius; t[6] = 0; t[7] = 1; }
    return;
  }
  const double rd = cost_change / mcc;
  const bool accepted = rd > kMinRelativeDecrease;
  if (accepted) {
    s.cur ^= 1;
    s.need_lin = 1;
    s.x_cost = cand_cost;
    const double c = 2.0 * rd - 1.0;
    s.radius = s.radius / fmax(1.0 / 3.0, 1.0 - c * c * c);
    s.radius = fmin(kMaxRadius, s.radius);
    s.decrease_factor = 2.0;
    s.successful++;
  } else {
    s.radius = s.radius / s.decrease_factor; s.decrease_factor *= 2.0;
  }
  if (t) { t[0] = s.x_cost; t[1] = cost_change; t[2] = s.gmax; t[3] = step_norm; t[4] = rd; t[5] = s.radius; t[6] = accepted; t[7] = 1; }
  if (s.iteration >= s.max_iterations) {
    s.done = 1; s.termination = TERM_MAX_ITER;
    if (accepted) s.jac_evals++;   // Ceres re-evaluates the Jacobian after the last accepted step too
  } else if (s.radius < kMinRadius) { s.done = 1; s.termination = TERM_MIN_RADIUS; }
}

// ---- geometry ---------------------------------------------------------------------------------------
struct Proj {
  double r0, r1;         // residual (already multiplied by invSigma2: quirk Q1)
  double a0, a1, a2;     // q * X (rotated point, without t)
  double p0, p1, p2;     // camera-frame point
};

// residual exactly as the functor computes it (Eigen _transformVector, K * p order)
__device__ __forceinline__ Proj project_obs(const double* __restrict__ cam, const double* __restrict__ X,
                                            const double fx, const double fy, const double cx, const double cy,
                                            const double u, const double v, const double w) {
  const double qx = cam[3], qy = cam[4], qz = cam[5], qw = cam[6];
  double uv0 = qy * X[2] - qz * X[1], uv1 = qz * X[0] - qx * X[2], uv2 = qx * X[1] - qy * X[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  Proj P;
  P.a0 = X[0] + qw * uv0 + (qy * uv2 - qz * uv1);
  P.a1 = X[1] + qw * uv1 + (qz * uv0 - qx * uv2);
  P.a2 = X[2] + qw * uv2 + (qx * uv1 - qy * uv0);
  P.p0 = P.a0 + cam[0]; P.p1 = P.a1 + cam[1]; P.p2 = P.a2 + cam[2];
  const double px = fx * P.p0 + cx * P.p2, py = fy * P.p1 + cy * P.p2;
  P.r0 = w * (u - px / P.p2);
  P.r1 = w * (v - py / P.p2);
  return P;
}

// tangent Jacobian wrt (t, delta): 2x6 row-major.  d p/d t = I, d p/d delta = -2 [q*X]x  (A.5)
__device__ __forceinline__ void jac_cam(const Proj& P, double fx, double fy, double w, double* Jc) {
  const double iz = 1.0 / P.p2;
  const double A00 = fx * iz, A02 = -fx * P.p0 * iz * iz, A11 = fy * iz, A12 = -fy * P.p1 * iz * iz;
  const double m = -w, m2 = 2.0 * w;
  Jc[0] = m * A00; Jc[1] = 0.0; Jc[2] = m * A02;
  Jc[3] = m2 * (-A02 * P.a1); Jc[4] = m2 * (-A00 * P.a2 + A02 * P.a0); Jc[5] = m2 * (A00 * P.a1);
  Jc[6] = 0.0; Jc[7] = m * A11; Jc[8] = m * A12;
  Jc[9] = m2 * (A11 * P.a2 - A12 * P.a1); Jc[10] = m2 * (A12 * P.a0); Jc[11] = m2 * (-A11 * P.a0);
}

// Jacobian wrt the point: -w * A * M(q), M = d(q*X)/dX of the _transformVector polynomial
__device__ __forceinline__ void jac_point(const Proj& P, const double* __restrict__ cam, double fx, double fy, double w,
                                          double* Jp) {
  const double qx = cam[3], qy = cam[4], qz = cam[5], qw = cam[6];
  const double M00 = 1.0 - 2.0 * (qy * qy + qz * qz), M01 = 2.0 * (qx * qy - qw * qz), M02 = 2.0 * (qx * qz + qw * qy);
  const double M10 = 2.0 * (qx * qy + qw * qz), M11 = 1.0 - 2.0 * (qx * qx + qz * qz), M12 = 2.0 * (qy * qz - qw * qx);
  const double M20 = 2.0 * (qx * qz - qw * qy), M21 = 2.0 * (qy * qz + qw * qx), M22 = 1.0 - 2.0 * (qx * qx + qy * qy);
  const double iz = 1.0 / P.p2;
  const double A00 = fx * iz, A02 = -fx * P.p0 * iz * iz, A11 = fy * iz, A12 = -fy * P.p1 * iz * iz;
  const double m = -w;
  Jp[0] = m * (A00 * M00 + A02 * M20); Jp[1] = m * (A00 * M01 + A02 * M21); Jp[2] = m * (A00 * M02 + A02 * M22);
  Jp[3] = m * (A11 * M10 + A12 * M20); Jp[4] = m * (A11 * M11 + A12 * M21); Jp[5] = m * (A11 * M12 + A12 * M22);
}

// loss: mode bit0 = a Huber block exists, bit1 = a loss-free block exists (LocalBA pass 1 has both, quirk Q2).
// Returns the cost contribution; *wsum = sum of rho' over the blocks (row weight for J'J and J'r).
__device__ __forceinline__ double loss_eval(double s, int mode, double* wsum) {
  double cost = 0.0, w = 0.0;
  if (mode & 1) {
    if (s > kHuberB) {
      const double r = sqrt(s);
      cost += 0.5 * (2.0 * kHuberA * r - kHuberB);
      w += fmax(2.2250738585072014e-308, kHuberA / r);
    } else { cost += 0.5 * s; w += 1.0; }
  }
  if (mode & 2) { cost += 0.5 * s; w += 1.0; }
  *wsum = w;
  return cost;
}

// EigenQuaternionParameterization::Plus
__device__ __forceinline__ void quat_plus(const double* __restrict__ q, const double* __restrict__ d, double* out) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = sin(n) / n;
    const double ax = s * d[0], ay = s * d[1], az = s * d[2], aw = cos(n);
    const double bx = q[0], by = q[1], bz = q[2], bw = q[3];
    out[3] = aw * bw - ax * bx - ay * by - az * bz;
    out[0] = aw * bx + ax * bw + ay * bz - az * by;
    out[1] = aw * by + ay * bw + az * bx - ax * bz;
    out[2] = aw * bz + az * bw + ax * by - ay * bx;
  } else {
    out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  }
}

// || x - Plus(x, -g) ||_inf of one pose block (t: 3 euclidean, q: manifold)
__device__ __forceinline__ double pose_gradient_max(const double* __restrict__ cam, const double* __restrict__ g) {
  double m = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
  const double ng[3] = {-g[3], -g[4], -g[5]};
  double qp[4];
  quat_plus(cam + 3, ng, qp);
  for (int i = 0; i < 4; i++) m = fmax(m, fabs(cam[3 + i] - qp[i]));
  return m;
}

// symmetric 3x3 (00 01 02 11 12 22) inverse through its Cholesky factor; false when not positive definite
__device__ __forceinline__ bool invert3_sym(const double* H, double* inv) {
  const double a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5];
  if (!(a > 0.0)) return false;
  const double l00 = sqrt(a), l10 = b / l00, l20 = c / l00;
  const double t11 = d - l10 * l10;
  if (!(t11 > 0.0)) return false;
  const double l11 = sqrt(t11), l21 = (e - l20 * l10) / l11;
  const double t22 = f - l20 * l20 - l21 * l21;
  if (!(t22 > 0.0)) return false;
  const double l22 = sqrt(t22);
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  inv[0] = i00 * i00 + i10 * i10 + i20 * i20;
  inv[1] = i10 * i11 + i20 * i21;
  inv[2] = i20 * i22;
  inv[3] = i11 * i11 + i21 * i21;
  inv[4] = i21 * i22;
  inv[5] = i22 * i22;
  return true;
}

// ---- deterministic block reductions ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the block; result valid in every thread.  scratch: >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += scratch[i];
    scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nw; i++) t = fmax(t, scratch[i]);
    scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

}  // namespace cmos
