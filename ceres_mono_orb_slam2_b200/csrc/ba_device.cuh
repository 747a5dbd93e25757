// Device-side math shared by the bundle-adjustment kernels: reprojection residual with analytic tangent
// Jacobians, Huber corrector, quaternion manifold, Levenberg-Marquardt state machine, deterministic block
// reductions.  All fp64.  Behaviour: include/CeresOptimizer.h:56-166 + the Ceres semantics of SURVEY.md A.5.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

namespace cmos {

// ---- programmatic dependent launch ------------------------------------------------------------------------------------------
// A solve is a chain of small, dependent kernels on one stream (LocalBundleAdjustment: 98 launches of 3-30 us).  Every kernel
// of these files starts with pdl_begin(): it lets the NEXT kernel of the stream be scheduled at once (its CTAs become resident
// and park in griddepcontrol.wait) and then waits until the PREVIOUS kernel has completed and its writes are visible — so a
// kernel starts computing the moment its predecessor retires instead of paying the launch latency after it.  launch_chain
// launches with the attribute that enables this; kernels launched without it see the two instructions as no-ops.
// CMOS_BA_NO_PDL=1 launches the ordinary way (A/B).
__device__ __forceinline__ void pdl_begin() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
  static const bool on = getenv("CMOS_BA_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);     // errors surface through cudaGetLastError like <<< >>>
}

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

// ---- Sim3 (OptimizeSim3: include/CeresOptimizer.h:168-268, src/CeresOptimizer.cc:24-47) -----------------------------
// Lie algebra vector [upsilon(3), omega(3), sigma]; group element as scale, rotation matrix, translation.  exp / log
// are Sophus' closed forms (sim3.hpp, sim_details.hpp: calcW / calcWInv with the 1e-10 small-angle branches); the
// rotation goes through Rodrigues' formula here (the CPU oracle goes through the quaternion exponential).
namespace cmos {

struct Sim3D { double s, R[9], t[3]; };

__device__ __forceinline__ void hat3(const double* w, double* O) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1]; O[3] = w[2]; O[4] = 0; O[5] = -w[0]; O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
__device__ __forceinline__ void mul33(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mul3v(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

__device__ inline void sim3_calc_w(const double* O, const double* O2, double theta, double sigma, double scale, double* W) {
  const double eps = 1e-10;
  double A, B, C;
  if (fabs(sigma) < eps) {
    C = 1.0;
    if (fabs(theta) < eps) { A = 0.5; B = 1.0 / 6.0; }
    else { const double t2 = theta * theta; A = (1.0 - cos(theta)) / t2; B = (theta - sin(theta)) / (t2 * theta); }
  } else {
    C = (scale - 1.0) / sigma;
    if (fabs(theta) < eps) {
      const double s2 = sigma * sigma;
      A = ((sigma - 1.0) * scale + 1.0) / s2;
      B = (scale * 0.5 * s2 + scale - 1.0 - sigma * scale) / (s2 * sigma);
    } else {
      const double t2 = theta * theta, a = scale * sin(theta), b = scale * cos(theta), c = t2 + sigma * sigma;
      A = (a * sigma + (1.0 - b) * theta) / (theta * c);
      B = (C - ((b - 1.0) * sigma + a * theta) / c) * 1.0 / t2;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) W[i] = A * O[i] + B * O2[i] + ((i % 4 == 0) ? C : 0.0);
}

__device__ inline void sim3_calc_w_inv(const double* O, const double* O2, double theta, double sigma, double scale, double* W) {
  const double eps = 1e-10;
  const double scale_sq = scale * scale, t2 = theta * theta, st = sin(theta), ct = cos(theta);
  double a, b, c;
  if (fabs(sigma * sigma) < eps) {
    c = 1.0 - 0.5 * sigma;
    a = -0.5;
    if (fabs(t2) < eps) b = 1.0 / 12.0;
    else b = (theta * st + 2.0 * ct - 2.0) / (2.0 * t2 * (ct - 1.0));
  } else {
    const double scale_cu = scale_sq * scale;
    c = sigma / (scale - 1.0);
    if (fabs(t2) < eps) {
      a = (-sigma * scale + scale - 1.0) / ((scale - 1.0) * (scale - 1.0));
      b = (scale_sq * sigma - 2.0 * scale_sq + scale * sigma + 2.0 * scale) / (2.0 * scale_cu - 6.0 * scale_sq + 6.0 * scale - 2.0);
    } else {
      const double ss = scale * st, sc = scale * ct;
      a = (theta * sc - theta - sigma * ss) / (theta * (scale_sq - 2.0 * sc + 1.0));
      b = -scale * (theta * ss - theta * st + sigma * sc - scale * sigma + sigma * ct - sigma) /
          (t2 * (scale_cu - 2.0 * scale * sc - scale_sq + 2.0 * sc + scale - 1.0));
    }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) W[i] = a * O[i] + b * O2[i] + ((i % 4 == 0) ? c : 0.0);
}

__device__ inline void sim3_exp(const double* v, Sim3D& S) {
  const double* w = v + 3;
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], theta = sqrt(t2);
  double O[9], O2[9], A, B;
  hat3(w, O);
  mul33(O, O, O2);
  if (t2 < 1e-20) { A = 1.0 - t2 / 6.0; B = 0.5 - t2 / 24.0; }
  else { A = sin(theta) / theta; B = (1.0 - cos(theta)) / t2; }
#pragma unroll
  for (int i = 0; i < 9; i++) S.R[i] = A * O[i] + B * O2[i] + ((i % 4 == 0) ? 1.0 : 0.0);
  S.s = exp(v[6]);
  double W[9];
  sim3_calc_w(O, O2, theta, v[6], S.s, W);
  mul3v(W, v, S.t);
}

__device__ inline void sim3_log(const Sim3D& S, double* v) {
  // rotation -> unit quaternion (Shepperd) -> rotation vector (2 atan2(|q_v|, q_w) / |q_v|) q_v
  const double* R = S.R;
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double t = sqrt(tr + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t; q[j] = (R[3 * j + i] + R[3 * i + j]) * t; q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= qn;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], qw = q[3];
  double two_atan, theta;
  if (n2 < 1e-20) { two_atan = 2.0 / qw - 2.0 / 3.0 * n2 / (qw * qw * qw); theta = 2.0 * n2 / qw; }
  else {
    const double n = sqrt(n2);
    two_atan = 2.0 * (qw < 0 ? -atan2(n, -qw) : atan2(n, qw)) / n;
    if (fabs(qw) < 1e-10) two_atan = (qw >= 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    theta = two_atan * n;
  }
  v[3] = two_atan * q[0]; v[4] = two_atan * q[1]; v[5] = two_atan * q[2];
  v[6] = log(S.s);
  double O[9], O2[9], W[9];
  hat3(v + 3, O);
  mul33(O, O, O2);
  sim3_calc_w_inv(O, O2, theta, v[6], S.s, W);
  mul3v(W, S.t, v);
}

__device__ inline void sim3_mul(const Sim3D& a, const Sim3D& b, Sim3D& o) {
  o.s = a.s * b.s;
  mul33(a.R, b.R, o.R);
  double r[3];
  mul3v(a.R, b.t, r);
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = a.s * r[i] + a.t[i];
}

__device__ inline void sim3_inverse(const Sim3D& a, Sim3D& o) {
  o.s = 1.0 / a.s;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) o.R[3 * i + j] = a.R[3 * j + i];
  double r[3];
  mul3v(o.R, a.t, r);
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = -o.s * r[i];
}

// Sim3Parameterization::Plus (CeresOptimizer.cc:24-41)
__device__ inline void sim3_plus(const double* x, const double* delta, double* out) {
  double d[7];
#pragma unroll
  for (int i = 0; i < 7; i++) d[i] = delta[i];
  d[6] = fmax(d[6], -20.0);
  Sim3D a, b, c;
  sim3_exp(x, a);
  sim3_exp(d, b);
  sim3_mul(a, b, c);
  sim3_log(c, out);
}

// Sim3ErrorTerm::Evaluate for p = S * P (or S^-1 * P): residual and, if J != null, the 2x7 Jacobian, times inv_sigma
__device__ __forceinline__ void sim3_error_term(const Sim3D& T, const double* __restrict__ P, double fx, double fy, double cx,
                                                double cy, double u, double v, double inv_sigma, double* r, double* J) {
  double rp[3];
  mul3v(T.R, P, rp);
  const double X = T.s * rp[0] + T.t[0], Y = T.s * rp[1] + T.t[1], Z = T.s * rp[2] + T.t[2];
  const double pr0 = fx * X + cx * Z, pr1 = fy * Y + cy * Z;
  r[0] = inv_sigma * (pr0 / Z - u);
  r[1] = inv_sigma * (pr1 / Z - v);
  if (!J) return;
  const double Z2 = Z * Z;
  const double a00 = fx / Z, a02 = -X * fx / Z2, a11 = fy / Z, a12 = -fy * Y / Z2;
  // J_camera * [I | -hat(p) | p]
  J[0] = inv_sigma * a00; J[1] = 0.0; J[2] = inv_sigma * a02;
  J[3] = inv_sigma * (a02 * Y); J[4] = inv_sigma * (a00 * Z - a02 * X); J[5] = inv_sigma * (-a00 * Y);
  J[6] = inv_sigma * (a00 * X + a02 * Z);
  J[7] = 0.0; J[8] = inv_sigma * a11; J[9] = inv_sigma * a12;
  J[10] = inv_sigma * (-a11 * Z + a12 * Y); J[11] = inv_sigma * (-a12 * X); J[12] = inv_sigma * (a11 * X);
  J[13] = inv_sigma * (a11 * Y + a12 * Z);
}

}  // namespace cmos
