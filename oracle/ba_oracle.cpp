// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (ceres_mono_orb_slam2_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
// CPU fp64 restatement of the optimisation behind CeresOptimizer::{PoseOptimization, LocalBundleAdjustment,
// BundleAdjustment} of the reference (src/CeresOptimizer.cc:59-225,275-342,344-599, cost functors
// include/CeresOptimizer.h:56-166).
//
// PARITY UNPINNED: the arithmetic lives in Ceres Solver (un-vendored, unpinned, needs <= 2.1 because of
// ceres::LocalParameterization) and the reference ships no tests or golden vectors for this path.  Ceres is not
// in this image, so this file restates its documented algorithm:
//   * cost functors evaluated with forward-mode dual numbers exactly as AutoDiffCostFunction does, including
//     Eigen's Quaternion::_transformVector formula and the K * p_c product order;
//   * HuberLoss(sqrt(5.991)) through the Triggs corrector (rho'' <= 0 -> alpha = 0: rows scaled by sqrt(rho'));
//   * EigenQuaternionParameterization (q stored x,y,z,w; Plus = delta_q * q; its 4x3 Jacobian);
//   * TrustRegionMinimizer + LevenbergMarquardtStrategy with Solver::Options defaults: Jacobi scaling
//     1/(1+||col||) fixed at iteration 0, D^2 = clamp(diag(J'J),1e-6,1e32)/radius reused after a rejected step,
//     step quality rho, radius schedule, function / parameter / gradient tolerances 1e-6 / 1e-8 / 1e-10,
//     max 5 consecutive invalid steps;
//   * SPARSE_NORMAL_CHOLESKY = exact solve of (J'J + D'D) y = J'r; done here by eliminating the 3x3 point blocks
//     first (algebraically the same linear solve as CHOLMOD's).
// and the reference's quirks Q1 (residual = e * invSigma2), Q2 (LocalBA pass 1 keeps the pass-0 Huber blocks and
// adds a loss-free copy of the inliers) and Q5 (PoseOptimization = one robust solve + one classification).
// It is sanity-checked in tests/test_oracle_ba.py against finite differences and scipy.optimize.least_squares.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------
// forward-mode dual number (ceres::Jet<double, N>)
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }   // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; v[k] = 1; }
};
template <int N> Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
template <int N> Jet<N> operator*(double s, const Jet<N>& f) { Jet<N> h; h.a = s * f.a; for (int i = 0; i < N; i++) h.v[i] = s * f.v[i]; return h; }
template <int N> Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h; h.a = s - f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
inline double operator_val(double x) { return x; }

// PoseGraph3dErrorTerm::operator() / PoseErrorTerm::operator()  (CeresOptimizer.h:63-88,122-143)
template <typename T>
void reprojection_residual(const T* t, const T* q, const T* X, const double K4[4], double u, double v, double w,
                           T* r) {
  // Eigen::Quaternion * vector (QuaternionBase::_transformVector): uv = 2 * vec x v; v + w*uv + vec x uv
  T uv0 = q[1] * X[2] - q[2] * X[1];
  T uv1 = q[2] * X[0] - q[0] * X[2];
  T uv2 = q[0] * X[1] - q[1] * X[0];
  uv0 = uv0 + uv0; uv1 = uv1 + uv1; uv2 = uv2 + uv2;
  T p0 = X[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1) + t[0];
  T p1 = X[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2) + t[1];
  T p2 = X[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0) + t[2];
  // projected = K * p_cp
  T px = K4[0] * p0 + K4[2] * p2;
  T py = K4[1] * p1 + K4[3] * p2;
  T pz = p2;
  T r0 = u - px / pz;
  T r1 = v - py / pz;
  r[0] = w * r0;   // applyOnTheLeft(diag(invSigma2, invSigma2))  — quirk Q1
  r[1] = w * r1;
}

// EigenQuaternionParameterization::Plus, q = (x, y, z, w)
void quat_plus(const double* q, const double* d, double* out) {
  const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    const double ax = s * d[0], ay = s * d[1], az = s * d[2], aw = std::cos(n);
    const double bx = q[0], by = q[1], bz = q[2], bw = q[3];
    out[3] = aw * bw - ax * bx - ay * by - az * bz;
    out[0] = aw * bx + ax * bw + ay * bz - az * by;
    out[1] = aw * by + ay * bw + az * bx - ax * bz;
    out[2] = aw * bz + az * bw + ax * by - ay * bx;
  } else {
    for (int i = 0; i < 4; i++) out[i] = q[i];
  }
}

// EigenQuaternionParameterization::ComputeJacobian, 4x3 row-major
void quat_plus_jacobian(const double* x, double* J) {
  J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
  J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
  J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
  J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}

constexpr double kHuberA = 2.447651936039926;   // sqrt(5.991)

struct Block {   // one ceres residual block
  int cam, pt;
  double u, v, w;
  int huber;     // loss_function != nullptr
};

struct Lin {     // one linearised residual block (already corrected by the loss)
  double r[2];
  double Jc[12];   // 2x6 tangent: t(3), delta(3)
  double Jp[6];    // 2x3
};

// Ceres evaluates residuals and Jacobians with options.num_threads threads (LocalBundleAdjustment sets 4,
// CeresOptimizer.cc:516).  g_eval_threads > 1 splits the residual blocks into contiguous chunks; the per-block costs are
// then summed in block order, so the result is bit-identical to the one-thread evaluation.  Only the evaluation is threaded
// (Schur elimination and the reduced solve stay on one thread).
int g_eval_threads = 1;
template <typename F>
void parallel_blocks(size_t n, F f) {
  const int T = (int)std::min<size_t>((size_t)std::max(1, g_eval_threads), std::max<size_t>(n / 256, 1));
  if (T <= 1) { f(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back([=] { f(n * t / T, n * (t + 1) / T); });
  for (auto& x : th) x.join();
}

struct Solver {
  int K = 0, M = 0;
  std::vector<double> cams, pts;   // current x
  std::vector<uint8_t> cam_const;
  bool pts_const = false;
  double K4[4];
  std::vector<Block> blocks;
  std::vector<int> cam_var;        // cam -> variable index or -1
  int Kv = 0;

  // ---- evaluation -------------------------------------------------------------------------------
  double cost_only(const std::vector<double>& c, const std::vector<double>& p) const {
    std::vector<double> bc(blocks.size());
    parallel_blocks(blocks.size(), [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++) {
        const Block& b = blocks[i];
        double r[2];
        reprojection_residual<double>(&c[7 * b.cam], &c[7 * b.cam + 3], &p[3 * b.pt], K4, b.u, b.v, b.w, r);
        const double s = r[0] * r[0] + r[1] * r[1];
        double rho0 = s;
        if (b.huber && s > kHuberA * kHuberA) rho0 = 2.0 * kHuberA * std::sqrt(s) - kHuberA * kHuberA;
        bc[i] = 0.5 * rho0;
      }
    });
    double cost = 0.0;
    for (double v : bc) cost += v;
    return cost;
  }

  double linearize(std::vector<Lin>& lin) const {
    lin.resize(blocks.size());
    std::vector<double> bc(blocks.size());
    parallel_blocks(blocks.size(), [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      const Block& b = blocks[i];
      typedef Jet<10> J10;
      J10 t[3], q[4], X[3], r[2];
      for (int k = 0; k < 3; k++) t[k] = J10(cams[7 * b.cam + k], k);
      for (int k = 0; k < 4; k++) q[k] = J10(cams[7 * b.cam + 3 + k], 3 + k);
      for (int k = 0; k < 3; k++) X[k] = J10(pts[3 * b.pt + k], 7 + k);
      reprojection_residual<J10>(t, q, X, K4, b.u, b.v, b.w, r);
      double PJ[12];
      quat_plus_jacobian(&cams[7 * b.cam + 3], PJ);
      Lin& L = lin[i];
      const double s = r[0].a * r[0].a + r[1].a * r[1].a;
      double rho0 = s, rho1 = 1.0;
      if (b.huber && s > kHuberA * kHuberA) {
        const double rr = std::sqrt(s);
        rho0 = 2.0 * kHuberA * rr - kHuberA * kHuberA;
        rho1 = std::max(std::numeric_limits<double>::min(), kHuberA / rr);
      }
      bc[i] = 0.5 * rho0;
      const double sc = b.huber ? std::sqrt(rho1) : 1.0;   // corrector with alpha = 0
      for (int row = 0; row < 2; row++) {
        L.r[row] = sc * r[row].a;
        for (int k = 0; k < 3; k++) L.Jc[row * 6 + k] = sc * r[row].v[k];
        for (int k = 0; k < 3; k++) {
          double acc = 0.0;
          for (int g = 0; g < 4; g++) acc += r[row].v[3 + g] * PJ[g * 3 + k];
          L.Jc[row * 6 + 3 + k] = sc * acc;
        }
        for (int k = 0; k < 3; k++) L.Jp[row * 3 + k] = sc * r[row].v[7 + k];
      }
    }
    });
    double cost = 0.0;
    for (double v : bc) cost += v;
    return cost;
  }

  int n_cols() const { return 6 * Kv + (pts_const ? 0 : 3 * M); }
  int cam_col(int cam) const { return 6 * cam_var[cam]; }
  int pt_col(int pt) const { return 6 * Kv + 3 * pt; }

  void plus(const std::vector<double>& c, const std::vector<double>& p, const double* delta, std::vector<double>& co,
            std::vector<double>& po) const {
    co = c; po = p;
    for (int k = 0; k < K; k++) {
      if (cam_var[k] < 0) continue;
      const double* d = delta + 6 * cam_var[k];
      for (int i = 0; i < 3; i++) co[7 * k + i] = c[7 * k + i] + d[i];
      quat_plus(&c[7 * k + 3], d + 3, &co[7 * k + 3]);
    }
    if (!pts_const)
      for (int j = 0; j < M; j++)
        for (int i = 0; i < 3; i++) po[3 * j + i] = p[3 * j + i] + delta[6 * Kv + 3 * j + i];
  }

  double ambient_norm(const std::vector<double>& c, const std::vector<double>& p) const {
    double s = 0.0;
    for (int k = 0; k < K; k++)
      if (cam_var[k] >= 0)
        for (int i = 0; i < 7; i++) s += c[7 * k + i] * c[7 * k + i];
    if (!pts_const)
      for (double x : p) s += x * x;
    return std::sqrt(s);
  }
  double ambient_diff(const std::vector<double>& c0, const std::vector<double>& p0, const std::vector<double>& c1,
                      const std::vector<double>& p1, bool inf_norm) const {
    double s = 0.0;
    auto acc = [&](double d) { if (inf_norm) s = std::max(s, std::fabs(d)); else s += d * d; };
    for (int k = 0; k < K; k++)
      if (cam_var[k] >= 0)
        for (int i = 0; i < 7; i++) acc(c0[7 * k + i] - c1[7 * k + i]);
    if (!pts_const)
      for (size_t i = 0; i < p0.size(); i++) acc(p0[i] - p1[i]);
    return inf_norm ? s : std::sqrt(s);
  }
};

// symmetric positive definite solve, in place (lower Cholesky, dense row-major storage); returns false if not PD.
// The loops are restricted to the ROW ENVELOPE of A (fc[i] = first non-zero column of row i): Cholesky fill stays inside
// it, so this is the exact factorisation CHOLMOD produces for Ceres on the reduced camera system (CeresOptimizer.cc:178-187
// selects SPARSE_SCHUR), and the products it skips are products with exact zeros — the result equals the dense loop's bit
// for bit.  It makes configs[4] (6000 unknowns, half-bandwidth 126) cost O(n w^2) instead of O(n^3).
bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  std::vector<int> fc(n), last(n);
  for (int i = 0; i < n; i++) {
    int c = 0;
    const double* ai = &A[(size_t)i * n];
    while (c < i && ai[c] == 0.0) c++;
    fc[i] = c;
  }
  for (int i = 0; i < n; i++) last[i] = i;
  for (int i = 0; i < n; i++) last[fc[i]] = std::max(last[fc[i]], i);
  for (int i = 1; i < n; i++) last[i] = std::max(last[i], last[i - 1]);   // rows > last[i] have fc > i
  for (int i = 0; i < n; i++) {
    double* ai = &A[(size_t)i * n];
    for (int j = fc[i]; j < i; j++) {
      const double* aj = &A[(size_t)j * n];
      double s = ai[j];
      for (int k = std::max(fc[i], fc[j]); k < j; k++) s -= ai[k] * aj[k];
      ai[j] = s / aj[j];
    }
    double d = ai[i];
    for (int k = fc[i]; k < i; k++) d -= ai[k] * ai[k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    ai[i] = std::sqrt(d);
  }
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = fc[i]; k < i; k++) s -= A[(size_t)i * n + k] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k <= last[i]; k++)
      if (fc[k] <= i) s -= A[(size_t)k * n + i] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

bool invert3_sym(const double* H /*6: 00 01 02 11 12 22*/, double* inv) {
  const double a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5];
  // Cholesky of the 3x3 (same PD test a sparse Cholesky would make)
  if (!(a > 0)) return false;
  const double l00 = std::sqrt(a), l10 = b / l00, l20 = c / l00;
  const double t11 = d - l10 * l10;
  if (!(t11 > 0)) return false;
  const double l11 = std::sqrt(t11), l21 = (e - l20 * l10) / l11;
  const double t22 = f - l20 * l20 - l21 * l21;
  if (!(t22 > 0)) return false;
  const double l22 = std::sqrt(t22);
  // inverse of L
  const double i00 = 1 / l00, i11 = 1 / l11, i22 = 1 / l22;
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

}  // namespace

extern "C" {

struct ba_oracle_summary {
  int32_t iterations;            // LM iterations performed (successful + unsuccessful + invalid)
  int32_t successful_steps;
  int32_t termination;           // 0 max iterations, 1 function tol, 2 parameter tol, 3 gradient tol, 4 user stop,
                                 // 5 failure, 6 min trust-region radius
  int32_t jacobian_evaluations;
  double initial_cost, final_cost;
};

// trace: [iterations+1][8] = cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius, accepted, valid
void ba_oracle_set_threads(int n) { g_eval_threads = n < 1 ? 1 : n; }

int ba_oracle_solve(int K, double* cams, const uint8_t* cam_const, int M, double* pts, int pts_const, int N,
                    const int32_t* obs_cam, const int32_t* obs_pt, const float* uv, const float* inv_sigma2,
                    const uint8_t* mode, const double* K4, int max_iterations, ba_oracle_summary* out, double* trace,
                    int trace_cap) {
  Solver S;
  S.K = K; S.M = M;
  S.cams.assign(cams, cams + 7 * (size_t)K);
  S.pts.assign(pts, pts + 3 * (size_t)M);
  S.cam_const.assign(cam_const, cam_const + K);
  S.pts_const = pts_const != 0;
  for (int i = 0; i < 4; i++) S.K4[i] = K4[i];
  S.cam_var.assign(K, -1);
  for (int k = 0; k < K; k++)
    if (!cam_const[k]) S.cam_var[k] = S.Kv++;
  for (int i = 0; i < N; i++) {
    const int m = mode ? mode[i] : 1;
    Block b{obs_cam[i], obs_pt[i], (double)uv[2 * i], (double)uv[2 * i + 1], (double)inv_sigma2[i], 1};
    if (m & 1) { b.huber = 1; S.blocks.push_back(b); }
    if (m & 2) { b.huber = 0; S.blocks.push_back(b); }
  }
  const int n = S.n_cols(), nc = 6 * S.Kv;
  const bool have_pts = !S.pts_const;
  ba_oracle_summary sum{};
  std::vector<Lin> lin;
  std::vector<double> scale(n, 1.0), grad(n), diag(n), lm(n), step(n), delta(n);
  std::vector<double> cand_c, cand_p;
  double x_cost = 0, radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int consecutive_invalid = 0;
  int tr = 0;
  auto put_trace = [&](double cost, double dc, double gmax, double sn, double rd, double rad, double acc, double valid) {
    if (trace && tr < trace_cap) {
      double* t = trace + 8 * tr;
      t[0] = cost; t[1] = dc; t[2] = gmax; t[3] = sn; t[4] = rd; t[5] = rad; t[6] = acc; t[7] = valid;
    }
    tr++;
  };

  // EvaluateGradientAndJacobian
  double gmax = 0.0;
  auto evaluate = [&](bool first) {
    x_cost = S.linearize(lin);
    sum.jacobian_evaluations++;
    std::fill(grad.begin(), grad.end(), 0.0);
    std::vector<double> col2(first ? n : 0, 0.0);
    for (size_t i = 0; i < lin.size(); i++) {
      const Block& b = S.blocks[i];
      const Lin& L = lin[i];
      const int cv = S.cam_var[b.cam];
      if (cv >= 0)
        for (int k = 0; k < 6; k++) {
          grad[6 * cv + k] += L.Jc[k] * L.r[0] + L.Jc[6 + k] * L.r[1];
          if (first) col2[6 * cv + k] += L.Jc[k] * L.Jc[k] + L.Jc[6 + k] * L.Jc[6 + k];
        }
      if (have_pts)
        for (int k = 0; k < 3; k++) {
          grad[nc + 3 * b.pt + k] += L.Jp[k] * L.r[0] + L.Jp[3 + k] * L.r[1];
          if (first) col2[nc + 3 * b.pt + k] += L.Jp[k] * L.Jp[k] + L.Jp[3 + k] * L.Jp[3 + k];
        }
    }
    if (first)
      for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + std::sqrt(col2[i]));
    // gradient_max_norm = || x - Plus(x, -g) ||_inf
    std::vector<double> ng(n);
    for (int i = 0; i < n; i++) ng[i] = -grad[i];
    std::vector<double> pc, pp;
    S.plus(S.cams, S.pts, ng.data(), pc, pp);
    gmax = S.ambient_diff(S.cams, S.pts, pc, pp, true);
  };

  double x_norm = S.ambient_norm(S.cams, S.pts);
  evaluate(true);
  sum.initial_cost = x_cost;
  put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
  int iteration = 0;
  sum.termination = 0;
  bool done = false;
  if (gmax <= 1e-10) { sum.termination = 3; done = true; }
  if (!done && max_iterations <= 0) done = true;

  std::vector<double> Hpp, Hinv, gp, Sm, rhs;
  while (!done) {
    iteration++;
    // ---- ComputeTrustRegionStep: LevenbergMarquardtStrategy::ComputeStep on the scaled Jacobian ----
    if (!reuse_diagonal) {
      std::fill(diag.begin(), diag.end(), 0.0);
      for (size_t i = 0; i < lin.size(); i++) {
        const Block& b = S.blocks[i];
        const Lin& L = lin[i];
        const int cv = S.cam_var[b.cam];
        if (cv >= 0)
          for (int k = 0; k < 6; k++) {
            const double s = scale[6 * cv + k];
            diag[6 * cv + k] += (L.Jc[k] * s) * (L.Jc[k] * s) + (L.Jc[6 + k] * s) * (L.Jc[6 + k] * s);
          }
        if (have_pts)
          for (int k = 0; k < 3; k++) {
            const double s = scale[nc + 3 * b.pt + k];
            diag[nc + 3 * b.pt + k] += (L.Jp[k] * s) * (L.Jp[k] * s) + (L.Jp[3 + k] * s) * (L.Jp[3 + k] * s);
          }
      }
      for (int i = 0; i < n; i++) diag[i] = std::min(std::max(diag[i], 1e-6), 1e32);
    }
    for (int i = 0; i < n; i++) lm[i] = std::sqrt(diag[i] / radius);
    // normal equations (scaled): (J'J + D'D) y = J'r, points eliminated first
    Sm.assign((size_t)nc * nc, 0.0);
    rhs.assign(nc, 0.0);
    bool ok = true;
    if (have_pts) { Hpp.assign(6 * (size_t)S.M, 0.0); gp.assign(3 * (size_t)S.M, 0.0); Hinv.assign(6 * (size_t)S.M, 0.0); }
    for (size_t i = 0; i < lin.size(); i++) {
      const Block& b = S.blocks[i];
      const Lin& L = lin[i];
      const int cv = S.cam_var[b.cam];
      if (cv >= 0) {
        double Js[12];
        for (int k = 0; k < 6; k++) { Js[k] = L.Jc[k] * scale[6 * cv + k]; Js[6 + k] = L.Jc[6 + k] * scale[6 * cv + k]; }
        for (int a = 0; a < 6; a++) {
          rhs[6 * cv + a] += Js[a] * L.r[0] + Js[6 + a] * L.r[1];
          for (int c = 0; c < 6; c++) Sm[(size_t)(6 * cv + a) * nc + 6 * cv + c] += Js[a] * Js[c] + Js[6 + a] * Js[6 + c];
        }
      }
      if (have_pts) {
        double Js[6];
        for (int k = 0; k < 3; k++) { Js[k] = L.Jp[k] * scale[nc + 3 * b.pt + k]; Js[3 + k] = L.Jp[3 + k] * scale[nc + 3 * b.pt + k]; }
        double* H = &Hpp[6 * (size_t)b.pt];
        H[0] += Js[0] * Js[0] + Js[3] * Js[3]; H[1] += Js[0] * Js[1] + Js[3] * Js[4]; H[2] += Js[0] * Js[2] + Js[3] * Js[5];
        H[3] += Js[1] * Js[1] + Js[4] * Js[4]; H[4] += Js[1] * Js[2] + Js[4] * Js[5]; H[5] += Js[2] * Js[2] + Js[5] * Js[5];
        for (int k = 0; k < 3; k++) gp[3 * (size_t)b.pt + k] += Js[k] * L.r[0] + Js[3 + k] * L.r[1];
      }
    }
    for (int i = 0; i < nc; i++) Sm[(size_t)i * nc + i] += lm[i] * lm[i];
    if (have_pts) {
      for (int j = 0; j < S.M && ok; j++) {
        double H[6];
        for (int k = 0; k < 6; k++) H[k] = Hpp[6 * (size_t)j + k];
        H[0] += lm[nc + 3 * j] * lm[nc + 3 * j]; H[3] += lm[nc + 3 * j + 1] * lm[nc + 3 * j + 1]; H[5] += lm[nc + 3 * j + 2] * lm[nc + 3 * j + 2];
        ok = invert3_sym(H, &Hinv[6 * (size_t)j]);
      }
      if (ok) {
        // group residual blocks by point
        std::vector<std::vector<int>> by_pt(S.M);
        for (size_t i = 0; i < lin.size(); i++)
          if (S.cam_var[S.blocks[i].cam] >= 0) by_pt[S.blocks[i].pt].push_back((int)i);
        std::vector<double> Wb;   // per block of the point: W = Jc_s' Jp_s (6x3)
        for (int j = 0; j < S.M; j++) {
          const std::vector<int>& ids = by_pt[j];
          if (ids.empty()) continue;
          const double* Hi = &Hinv[6 * (size_t)j];
          const double Hf[9] = {Hi[0], Hi[1], Hi[2], Hi[1], Hi[3], Hi[4], Hi[2], Hi[4], Hi[5]};
          Wb.assign(ids.size() * 18, 0.0);
          std::vector<double> WH(ids.size() * 18, 0.0);
          for (size_t a = 0; a < ids.size(); a++) {
            const Lin& L = lin[ids[a]];
            const int cv = S.cam_var[S.blocks[ids[a]].cam];
            for (int r6 = 0; r6 < 6; r6++)
              for (int c3 = 0; c3 < 3; c3++)
                Wb[a * 18 + r6 * 3 + c3] = (L.Jc[r6] * L.Jp[c3] + L.Jc[6 + r6] * L.Jp[3 + c3]) * scale[6 * cv + r6] * scale[nc + 3 * j + c3];
            for (int r6 = 0; r6 < 6; r6++)
              for (int c3 = 0; c3 < 3; c3++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += Wb[a * 18 + r6 * 3 + k] * Hf[k * 3 + c3];
                WH[a * 18 + r6 * 3 + c3] = acc;
              }
          }
          for (size_t a = 0; a < ids.size(); a++) {
            const int ca = S.cam_var[S.blocks[ids[a]].cam];
            for (int r6 = 0; r6 < 6; r6++) {
              double acc = 0;
              for (int k = 0; k < 3; k++) acc += WH[a * 18 + r6 * 3 + k] * gp[3 * (size_t)j + k];
              rhs[6 * ca + r6] -= acc;
            }
            for (size_t bq = 0; bq < ids.size(); bq++) {
              const int cb = S.cam_var[S.blocks[ids[bq]].cam];
              for (int r6 = 0; r6 < 6; r6++)
                for (int c6 = 0; c6 < 6; c6++) {
                  double acc = 0;
                  for (int k = 0; k < 3; k++) acc += WH[a * 18 + r6 * 3 + k] * Wb[bq * 18 + c6 * 3 + k];
                  Sm[(size_t)(6 * ca + r6) * nc + 6 * cb + c6] -= acc;
                }
            }
          }
        }
      }
    }
    std::vector<double> yc = rhs;
    if (ok && nc > 0) ok = cholesky_solve(Sm, nc, yc);
    bool step_valid = ok;
    double model_cost_change = 0.0;
    if (ok) {
      for (int i = 0; i < nc; i++) step[i] = -yc[i];
      if (have_pts) {
        // y_p = Hinv (g_p - W' y_c)
        std::vector<double> acc(3 * (size_t)S.M, 0.0);
        for (size_t i = 0; i < lin.size(); i++) {
          const Block& b = S.blocks[i];
          const int cv = S.cam_var[b.cam];
          if (cv < 0) continue;
          const Lin& L = lin[i];
          for (int c3 = 0; c3 < 3; c3++) {
            double a = 0;
            for (int r6 = 0; r6 < 6; r6++)
              a += (L.Jc[r6] * L.Jp[c3] + L.Jc[6 + r6] * L.Jp[3 + c3]) * scale[6 * cv + r6] * scale[nc + 3 * b.pt + c3] * yc[6 * cv + r6];
            acc[3 * (size_t)b.pt + c3] += a;
          }
        }
        for (int j = 0; j < S.M; j++) {
          const double* Hi = &Hinv[6 * (size_t)j];
          const double b0 = gp[3 * (size_t)j] - acc[3 * (size_t)j], b1 = gp[3 * (size_t)j + 1] - acc[3 * (size_t)j + 1],
                       b2 = gp[3 * (size_t)j + 2] - acc[3 * (size_t)j + 2];
          step[nc + 3 * j] = -(Hi[0] * b0 + Hi[1] * b1 + Hi[2] * b2);
          step[nc + 3 * j + 1] = -(Hi[1] * b0 + Hi[3] * b1 + Hi[4] * b2);
          step[nc + 3 * j + 2] = -(Hi[2] * b0 + Hi[4] * b1 + Hi[5] * b2);
        }
      }
      for (int i = 0; i < n; i++)
        if (!std::isfinite(step[i])) step_valid = false;
    }
    reuse_diagonal = true;
    if (step_valid) {
      // model_cost_change = -(J s)'(r + J s / 2) with the scaled Jacobian
      for (size_t i = 0; i < lin.size(); i++) {
        const Block& b = S.blocks[i];
        const Lin& L = lin[i];
        const int cv = S.cam_var[b.cam];
        for (int row = 0; row < 2; row++) {
          double m = 0.0;
          if (cv >= 0)
            for (int k = 0; k < 6; k++) m += L.Jc[row * 6 + k] * scale[6 * cv + k] * step[6 * cv + k];
          if (have_pts)
            for (int k = 0; k < 3; k++) m += L.Jp[row * 3 + k] * scale[nc + 3 * b.pt + k] * step[nc + 3 * b.pt + k];
          model_cost_change -= m * (L.r[row] + m / 2.0);
        }
      }
      if (!(model_cost_change > 0.0)) step_valid = false;
    }
    if (!step_valid) {
      // HandleInvalidStep
      consecutive_invalid++;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
      if (consecutive_invalid >= 5) { sum.termination = 5; break; }
      if (iteration >= max_iterations) { sum.termination = 0; break; }
      if (radius < 1e-32) { sum.termination = 6; break; }
      continue;
    }
    consecutive_invalid = 0;
    for (int i = 0; i < n; i++) delta[i] = step[i] * scale[i];
    S.plus(S.cams, S.pts, delta.data(), cand_c, cand_p);
    const double cand_cost = S.cost_only(cand_c, cand_p);
    const double step_norm = S.ambient_diff(S.cams, S.pts, cand_c, cand_p, false);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) {   // ParameterToleranceReached
      sum.termination = 2;
      put_trace(x_cost, 0, gmax, step_norm, 0, radius, 0, 1);
      break;
    }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) {   // FunctionToleranceReached
      sum.termination = 1;
      put_trace(x_cost, cost_change, gmax, step_norm, 0, radius, 0, 1);
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    bool accepted = relative_decrease > 1e-3;
    if (accepted) {
      S.cams = cand_c; S.pts = cand_p;
      x_norm = S.ambient_norm(S.cams, S.pts);
      evaluate(false);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      sum.successful_steps++;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
    put_trace(x_cost, cost_change, gmax, step_norm, relative_decrease, radius, accepted ? 1 : 0, 1);
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= max_iterations) { sum.termination = 0; break; }
    if (gmax <= 1e-10) { sum.termination = 3; break; }
    if (radius < 1e-32) { sum.termination = 6; break; }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  std::memcpy(cams, S.cams.data(), sizeof(double) * 7 * (size_t)K);
  if (!S.pts_const) std::memcpy(pts, S.pts.data(), sizeof(double) * 3 * (size_t)M);
  if (out) *out = sum;
  return tr;
}

// residuals + tangent Jacobians of one observation (tests: finite differences, scipy cross-check)
void ba_oracle_residual(const double* cam7, const double* X, const double* K4, float u, float v, float inv_sigma2,
                        double* r, double* Jc /*2x6*/, double* Jp /*2x3*/) {
  Solver S;
  S.K = 1; S.M = 1; S.cams.assign(cam7, cam7 + 7); S.pts.assign(X, X + 3);
  for (int i = 0; i < 4; i++) S.K4[i] = K4[i];
  S.cam_var.assign(1, 0); S.Kv = 1;
  S.blocks.push_back(Block{0, 0, (double)u, (double)v, (double)inv_sigma2, 0});
  std::vector<Lin> lin;
  S.linearize(lin);
  for (int i = 0; i < 2; i++) r[i] = lin[0].r[i];
  for (int i = 0; i < 12; i++) Jc[i] = lin[0].Jc[i];
  for (int i = 0; i < 6; i++) Jp[i] = lin[0].Jp[i];
}

void ba_oracle_quat_plus(const double* q, const double* d, double* out) { quat_plus(q, d, out); }

// CeresOptimizer::CheckOutlier (CeresOptimizer.cc:227-241) + the z <= 0 test of LocalBundleAdjustment (:558-564)
static void check_obs(const double* cam7, const double* X, const double* K4, float u, float v, float inv_sigma,
                      double* chi2, double* z) {
  double r[2];
  // pixel = K * (q * X + t); error = obs - pixel.xy / pixel.z; error^2 * inv_sigma
  reprojection_residual<double>(cam7, cam7 + 3, X, K4, (double)u, (double)v, 1.0, r);
  *chi2 = (r[0] * r[0] + r[1] * r[1]) * (double)inv_sigma;
  double one[4] = {0, 0, 0, 0};
  (void)one;
  // z of q * X + t
  const double* q = cam7 + 3;
  double uv0 = q[1] * X[2] - q[2] * X[1], uv1 = q[2] * X[0] - q[0] * X[2], uv2 = q[0] * X[1] - q[1] * X[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  *z = X[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0) + cam7[2];
}

// CeresOptimizer::PoseOptimization (CeresOptimizer.cc:275-342) for one frame.
// Returns n_initial_correspondences - n_bad; pose7 in/out (q normalised on output, :335); is_outlier out.
int ba_oracle_pose_optimization(double* pose7, int n, const double* xw, const float* uv, const float* inv_sigma2,
                                const double* K4, int max_iterations, uint8_t* is_outlier, ba_oracle_summary* out,
                                double* trace, int trace_cap) {
  if (n < 3) { if (out) std::memset(out, 0, sizeof(*out)); return 0; }
  std::vector<int32_t> oc(n, 0), op(n);
  for (int i = 0; i < n; i++) op[i] = i;
  std::vector<double> pts(xw, xw + 3 * (size_t)n);
  const uint8_t cc = 0;
  ba_oracle_solve(1, pose7, &cc, n, pts.data(), 1, n, oc.data(), op.data(), uv, inv_sigma2, nullptr, K4,
                  max_iterations, out, trace, trace_cap);
  int n_bad = 0;
  for (int i = 0; i < n; i++) {
    double chi2, z;
    check_obs(pose7, xw + 3 * (size_t)i, K4, uv[2 * i], uv[2 * i + 1], inv_sigma2[i], &chi2, &z);
    is_outlier[i] = chi2 > 5.991;
    n_bad += is_outlier[i];
  }
  const double nq = std::sqrt(pose7[3] * pose7[3] + pose7[4] * pose7[4] + pose7[5] * pose7[5] + pose7[6] * pose7[6]);
  for (int i = 3; i < 7; i++) pose7[i] /= nq;
  return n - n_bad;
}

// CeresOptimizer::LocalBundleAdjustment (CeresOptimizer.cc:344-599) on a flattened graph.
// cam_flags: bit0 = constant (fixed keyframe, or keyframe id 0), bit1 = not a local keyframe (no outlier scan).
// erase[n_obs] out: observations the reference would erase from the map (:573-581).  summaries[2].
void ba_oracle_local(int K, double* cams, const uint8_t* cam_flags, int M, double* pts, int N, const int32_t* obs_cam,
                     const int32_t* obs_pt, const float* uv, const float* inv_sigma2, const double* K4,
                     int iters_pass0, int iters_pass1, uint8_t* erase, ba_oracle_summary* summaries) {
  std::vector<uint8_t> cc(K), mode(N, 1);
  for (int k = 0; k < K; k++) cc[k] = cam_flags[k] & 1;
  auto scan = [&]() {
    for (int i = 0; i < N; i++) {
      erase[i] = 0;
      if (cam_flags[obs_cam[i]] & 2) continue;
      double chi2, z;
      check_obs(cams + 7 * (size_t)obs_cam[i], pts + 3 * (size_t)obs_pt[i], K4, uv[2 * i], uv[2 * i + 1], inv_sigma2[i],
                &chi2, &z);
      erase[i] = (chi2 > 5.991) || (z <= 0);
    }
  };
  ba_oracle_solve(K, cams, cc.data(), M, pts, 0, N, obs_cam, obs_pt, uv, inv_sigma2, mode.data(), K4, iters_pass0,
                  summaries, nullptr, 0);
  scan();
  for (int i = 0; i < N; i++) mode[i] = erase[i] ? 1 : 3;   // quirk Q2: Huber blocks stay, inliers added again without loss
  ba_oracle_solve(K, cams, cc.data(), M, pts, 0, N, obs_cam, obs_pt, uv, inv_sigma2, mode.data(), K4, iters_pass1,
                  summaries + 1, nullptr, 0);
  scan();
}

// CeresOptimizer::BundleAdjustment (CeresOptimizer.cc:59-225)
void ba_oracle_global(int K, double* cams, const uint8_t* cam_const, int M, double* pts, int N, const int32_t* obs_cam,
                      const int32_t* obs_pt, const float* uv, const float* inv_sigma2, const double* K4, int n_iterations,
                      int robust, ba_oracle_summary* summary, double* trace, int trace_cap) {
  std::vector<uint8_t> mode(N, robust ? 1 : 2);
  ba_oracle_solve(K, cams, cam_const, M, pts, 0, N, obs_cam, obs_pt, uv, inv_sigma2, mode.data(), K4, n_iterations,
                  summary, trace, trace_cap);
}

// ---------------------------------------------------------------------------------------------------
// CeresOptimizer::OptimizeSim3 (CeresOptimizer.cc:601-735): one 7-vector (Sim3 Lie algebra, [upsilon, omega, sigma]) is
// optimised over 2 x n reprojection residuals — keyframe-2 points seen in keyframe 1 through S12, keyframe-1 points seen
// in keyframe 2 through S12^-1 (Sim3ErrorTerm, include/CeresOptimizer.h:168-255) — with HuberLoss(sqrt(th2)),
// Sim3Parameterization (Plus = log(exp(x) * exp(delta)), sigma step clamped at -20, identity local Jacobian,
// CeresOptimizer.cc:24-47), max 100 iterations, Ceres defaults otherwise.  The cost functor's Jacobian is the LEFT
// perturbation formula for both directions while Plus multiplies on the right; that is the reference's behaviour and is
// reproduced.  sqrt_information = inv_sigma * I with inv_sigma = inv_level_sigma2s_ (quirk Q1 again).
// Sophus (un-vendored, unpinned) provides Sim3d::exp / log / inverse; they are restated from Sophus' published
// closed forms (so3.hpp, rxso3.hpp, sim3.hpp, sim_details.hpp: quaternion exponential, calcW / calcWInv with the
// 1e-10 small-angle branches).  PARITY UNPINNED.
}  // extern "C" (the helpers below have C++ linkage: their names — exp, log — must not collide with libm's)

namespace sim3o {
constexpr double kEps = 1e-10;
struct Sim3 { double s, q[4] /*x y z w, unit*/, t[3]; };

void hat(const double* w, double* O) { O[0] = 0; O[1] = -w[2]; O[2] = w[1]; O[3] = w[2]; O[4] = 0; O[5] = -w[0]; O[6] = -w[1]; O[7] = w[0]; O[8] = 0; }
void mat3mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k = 0; k < 3; k++) a += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = a; }
}
void quat_mul(const double* a, const double* b, double* o) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
}
void quat_rot(const double* q, const double* v, double* o) {        // unit quaternion: q v q*
  const double uv0 = 2 * (q[1] * v[2] - q[2] * v[1]), uv1 = 2 * (q[2] * v[0] - q[0] * v[2]), uv2 = 2 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1);
  o[1] = v[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2);
  o[2] = v[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0);
}
void so3_exp(const double* w, double* q, double* theta) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  *theta = std::sqrt(t2);
  if (t2 < kEps * kEps) { const double t4 = t2 * t2; imag = 0.5 - t2 / 48.0 + t4 / 3840.0; real = 1.0 - t2 / 8.0 + t4 / 384.0; }
  else { const double h = 0.5 * *theta; imag = std::sin(h) / *theta; real = std::cos(h); }
  q[0] = imag * w[0]; q[1] = imag * w[1]; q[2] = imag * w[2]; q[3] = real;
}
void so3_log(const double* q, double* w, double* theta) {
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], qw = q[3];
  double two_atan;
  if (n2 < kEps * kEps) { two_atan = 2.0 / qw - 2.0 / 3.0 * n2 / (qw * qw * qw); *theta = 2.0 * n2 / qw; }
  else {
    const double n = std::sqrt(n2);
    // atan(n / w) chosen so that the rotation angle lies in (-pi, pi]
    two_atan = 2.0 * (qw < 0 ? -std::atan2(n, -qw) : std::atan2(n, qw)) / n;
    if (std::fabs(qw) < kEps) two_atan = (qw >= 0 ? M_PI : -M_PI) / n;
    *theta = two_atan * n;
  }
  w[0] = two_atan * q[0]; w[1] = two_atan * q[1]; w[2] = two_atan * q[2];
}
void calcW(const double* w, double theta, double sigma, double scale, double* W) {
  double O[9], O2[9];
  hat(w, O); mat3mul(O, O, O2);
  double A, B, C;
  if (std::fabs(sigma) < kEps) {
    C = 1.0;
    if (std::fabs(theta) < kEps) { A = 0.5; B = 1.0 / 6.0; }
    else { const double t2 = theta * theta; A = (1.0 - std::cos(theta)) / t2; B = (theta - std::sin(theta)) / (t2 * theta); }
  } else {
    C = (scale - 1.0) / sigma;
    if (std::fabs(theta) < kEps) {
      const double s2 = sigma * sigma;
      A = ((sigma - 1.0) * scale + 1.0) / s2;
      B = (scale * 0.5 * s2 + scale - 1.0 - sigma * scale) / (s2 * sigma);
    } else {
      const double t2 = theta * theta, a = scale * std::sin(theta), b = scale * std::cos(theta), c = t2 + sigma * sigma;
      A = (a * sigma + (1.0 - b) * theta) / (theta * c);
      B = (C - ((b - 1.0) * sigma + a * theta) / c) * 1.0 / t2;
    }
  }
  for (int i = 0; i < 9; i++) W[i] = A * O[i] + B * O2[i] + ((i % 4 == 0) ? C : 0.0);
}
void calcWInv(const double* w, double theta, double sigma, double scale, double* W) {
  double O[9], O2[9];
  hat(w, O); mat3mul(O, O, O2);
  const double scale_sq = scale * scale, t2 = theta * theta, st = std::sin(theta), ct = std::cos(theta);
  double a, b, c;
  if (std::fabs(sigma * sigma) < kEps) {
    c = 1.0 - 0.5 * sigma;
    a = -0.5;
    if (std::fabs(t2) < kEps) b = 1.0 / 12.0;
    else b = (theta * st + 2.0 * ct - 2.0) / (2.0 * t2 * (ct - 1.0));
  } else {
    const double scale_cu = scale_sq * scale;
    c = sigma / (scale - 1.0);
    if (std::fabs(t2) < kEps) {
      a = (-sigma * scale + scale - 1.0) / ((scale - 1.0) * (scale - 1.0));
      b = (scale_sq * sigma - 2.0 * scale_sq + scale * sigma + 2.0 * scale) / (2.0 * scale_cu - 6.0 * scale_sq + 6.0 * scale - 2.0);
    } else {
      const double ss = scale * st, sc = scale * ct;
      a = (theta * sc - theta - sigma * ss) / (theta * (scale_sq - 2.0 * sc + 1.0));
      b = -scale * (theta * ss - theta * st + sigma * sc - scale * sigma + sigma * ct - sigma) /
          (t2 * (scale_cu - 2.0 * scale * sc - scale_sq + 2.0 * sc + scale - 1.0));
    }
  }
  for (int i = 0; i < 9; i++) W[i] = a * O[i] + b * O2[i] + ((i % 4 == 0) ? c : 0.0);
}
Sim3 exp(const double* v) {
  Sim3 S;
  double theta;
  so3_exp(v + 3, S.q, &theta);
  S.s = std::exp(v[6]);
  double W[9];
  calcW(v + 3, theta, v[6], S.s, W);
  for (int i = 0; i < 3; i++) S.t[i] = W[3 * i] * v[0] + W[3 * i + 1] * v[1] + W[3 * i + 2] * v[2];
  return S;
}
void log(const Sim3& S, double* v) {
  double theta;
  so3_log(S.q, v + 3, &theta);
  v[6] = std::log(S.s);
  double W[9];
  calcWInv(v + 3, theta, v[6], S.s, W);
  for (int i = 0; i < 3; i++) v[i] = W[3 * i] * S.t[0] + W[3 * i + 1] * S.t[1] + W[3 * i + 2] * S.t[2];
}
Sim3 mul(const Sim3& a, const Sim3& b) {
  Sim3 o;
  o.s = a.s * b.s;
  quat_mul(a.q, b.q, o.q);
  const double n = std::sqrt(o.q[0] * o.q[0] + o.q[1] * o.q[1] + o.q[2] * o.q[2] + o.q[3] * o.q[3]);
  for (int i = 0; i < 4; i++) o.q[i] /= n;
  double r[3];
  quat_rot(a.q, b.t, r);
  for (int i = 0; i < 3; i++) o.t[i] = a.s * r[i] + a.t[i];
  return o;
}
Sim3 inverse(const Sim3& a) {
  Sim3 o;
  o.s = 1.0 / a.s;
  o.q[0] = -a.q[0]; o.q[1] = -a.q[1]; o.q[2] = -a.q[2]; o.q[3] = a.q[3];
  double r[3];
  quat_rot(o.q, a.t, r);
  for (int i = 0; i < 3; i++) o.t[i] = -o.s * r[i];
  return o;
}
void act(const Sim3& S, const double* P, double* o) {
  double r[3];
  quat_rot(S.q, P, r);
  for (int i = 0; i < 3; i++) o[i] = S.s * r[i] + S.t[i];
}
void rotation_matrix(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
void quat_from_matrix(const double* R, double* q) {   // unit quaternion of a rotation matrix (Shepperd)
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) { double t = std::sqrt(tr + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t; }
  else {
    int i = 0; if (R[4] > R[0]) i = 1; if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t; q[j] = (R[3 * j + i] + R[3 * i + j]) * t; q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
void plus(const double* x, const double* delta, double* out) {     // Sim3Parameterization::Plus
  double d[7];
  for (int i = 0; i < 7; i++) d[i] = delta[i];
  d[6] = std::max(d[6], -20.);
  log(mul(exp(x), exp(d)), out);
}
// Sim3ErrorTerm::Evaluate: residual (2) and Jacobian (2x7 row-major), both already multiplied by inv_sigma
void error_term(const double* lie, const double* K4, const double* obs, const double* P, double inv_sigma, bool do_inverse,
                double* r, double* J) {
  const Sim3 S = exp(lie);
  double p[3];
  if (!do_inverse) act(S, P, p); else act(inverse(S), P, p);
  const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  const double pr0 = fx * p[0] + cx * p[2], pr1 = fy * p[1] + cy * p[2], pr2 = p[2];
  r[0] = inv_sigma * (pr0 / pr2 - obs[0]);
  r[1] = inv_sigma * (pr1 / pr2 - obs[1]);
  if (!J) return;
  const double X = p[0], Y = p[1], Z = p[2], Z2 = Z * Z;
  const double Jc[6] = {fx / Z, 0., -X * fx / Z2, 0, fy / Z, -fy * Y / Z2};
  const double L[21] = {1, 0, 0, 0, Z, -Y, X,      // [I | -hat(p) | p]
                        0, 1, 0, -Z, 0, X, Y,
                        0, 0, 1, Y, -X, 0, Z};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 7; j++) {
      double a = 0;
      for (int k = 0; k < 3; k++) a += Jc[3 * i + k] * L[7 * k + j];
      J[7 * i + j] = inv_sigma * a;
    }
}
}  // namespace sim3o

extern "C" {

void ba_oracle_sim3_exp(const double* lie, double* s, double* R9, double* t3) {
  const sim3o::Sim3 S = sim3o::exp(lie);
  *s = S.s; sim3o::rotation_matrix(S.q, R9); std::memcpy(t3, S.t, 24);
}
void ba_oracle_sim3_log(double s, const double* R9, const double* t3, double* lie) {
  sim3o::Sim3 S; S.s = s; sim3o::quat_from_matrix(R9, S.q); std::memcpy(S.t, t3, 24);
  sim3o::log(S, lie);
}
void ba_oracle_sim3_plus(const double* x, const double* delta, double* out) { sim3o::plus(x, delta, out); }
void ba_oracle_sim3_error_term(const double* lie, const double* K4, const double* obs, const double* P, double inv_sigma,
                               int do_inverse, double* r, double* J) {
  sim3o::error_term(lie, K4, obs, P, inv_sigma, do_inverse != 0, r, J);
}

// OptimizeSim3.  S12 in/out as scale, rotation (row-major), translation; per correspondence: obs1 (keypoint of keyframe 1),
// inv_sigma1, P3D2c (keyframe 2's map point in camera-2 coordinates), obs2, inv_sigma2, P3D1c.  is_bad[n] out.
// Returns the reference's return value (0 when fewer than 10 inliers remain).  lie7 out = the optimised log.
int ba_oracle_optimize_sim3(int n, double* s12, double* R12, double* t12, const double* K1, const double* K2,
                            const float* obs1, const float* inv_sigma1, const double* P3D2c, const float* obs2,
                            const float* inv_sigma2, const double* P3D1c, float th2, int max_iterations, uint8_t* is_bad,
                            double* lie7, ba_oracle_summary* out, double* trace, int trace_cap) {
  using namespace sim3o;
  double x[7];
  { Sim3 S; S.s = *s12; quat_from_matrix(R12, S.q); std::memcpy(S.t, t12, 24); log(S, x); }
  const double huber_a = std::sqrt((double)th2), huber_b = huber_a * huber_a;
  ba_oracle_summary sum{};
  const int m = 2 * n;      // residual blocks
  std::vector<double> r(2 * (size_t)m), J(14 * (size_t)m);
  double scale[7], grad[7], H[49];
  double x_cost = 0, x_norm = 0, gmax = 0, radius = 1e4, decrease_factor = 2.0;
  int consecutive_invalid = 0, tr = 0;
  auto put_trace = [&](double cost, double dc, double gm, double sn, double rd, double rad, double acc, double valid) {
    if (trace && tr < trace_cap) { double* t = trace + 8 * tr; t[0] = cost; t[1] = dc; t[2] = gm; t[3] = sn; t[4] = rd; t[5] = rad; t[6] = acc; t[7] = valid; }
    tr++;
  };
  auto block = [&](const double* lie, int b, double* rr, double* JJ) {
    const int i = b >> 1;
    const double o1[2] = {obs1[2 * i], obs1[2 * i + 1]}, o2[2] = {obs2[2 * i], obs2[2 * i + 1]};
    if ((b & 1) == 0) error_term(lie, K1, o1, P3D2c + 3 * i, inv_sigma1[i], false, rr, JJ);
    else error_term(lie, K2, o2, P3D1c + 3 * i, inv_sigma2[i], true, rr, JJ);
  };
  auto rho = [&](double s, double* w) {
    if (s > huber_b) { const double rt = std::sqrt(s); *w = std::max(std::numeric_limits<double>::min(), huber_a / rt); return 2.0 * huber_a * rt - huber_b; }
    *w = 1.0; return s;
  };
  auto cost_only = [&](const double* lie) {
    double c = 0;
    for (int b = 0; b < m; b++) { double rr[2], w; block(lie, b, rr, nullptr); c += 0.5 * rho(rr[0] * rr[0] + rr[1] * rr[1], &w); }
    return c;
  };
  auto evaluate = [&](bool first) {
    double c = 0;
    std::fill(grad, grad + 7, 0.0); std::fill(H, H + 49, 0.0);
    for (int b = 0; b < m; b++) {
      double* rr = &r[2 * (size_t)b]; double* JJ = &J[14 * (size_t)b];
      block(x, b, rr, JJ);
      double w;
      c += 0.5 * rho(rr[0] * rr[0] + rr[1] * rr[1], &w);
      const double sw = std::sqrt(w);             // Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
      rr[0] *= sw; rr[1] *= sw;
      for (int k = 0; k < 14; k++) JJ[k] *= sw;
      for (int a = 0; a < 7; a++) {
        grad[a] += JJ[a] * rr[0] + JJ[7 + a] * rr[1];
        for (int d = 0; d < 7; d++) H[7 * a + d] += JJ[a] * JJ[d] + JJ[7 + a] * JJ[7 + d];
      }
    }
    x_cost = c;
    sum.jacobian_evaluations++;
    if (first) for (int a = 0; a < 7; a++) scale[a] = 1.0 / (1.0 + std::sqrt(H[8 * a]));
    double ng[7], xp[7];
    for (int a = 0; a < 7; a++) ng[a] = -grad[a];
    plus(x, ng, xp);
    gmax = 0; x_norm = 0;
    for (int a = 0; a < 7; a++) { gmax = std::max(gmax, std::fabs(x[a] - xp[a])); x_norm += x[a] * x[a]; }
    x_norm = std::sqrt(x_norm);
  };
  evaluate(true);
  sum.initial_cost = x_cost;
  put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
  int iteration = 0;
  sum.termination = 0;
  if (max_iterations == 0) {}
  else if (gmax <= 1e-10) sum.termination = 3;
  else for (;;) {
    iteration++;
    // LevenbergMarquardtStrategy::ComputeStep on the Jacobi-scaled system
    std::vector<double> A(49), b(7);
    double D2[7], gs[7];
    for (int a = 0; a < 7; a++) {
      D2[a] = std::min(std::max(scale[a] * scale[a] * H[8 * a], 1e-6), 1e32) / radius;
      gs[a] = scale[a] * grad[a];
      b[a] = gs[a];
      for (int d = 0; d < 7; d++) A[7 * a + d] = scale[a] * scale[d] * H[7 * a + d] + (a == d ? D2[a] : 0.0);
    }
    bool ok = cholesky_solve(A, 7, b);
    double step[7], delta[7], mcc = 0.0;
    for (int a = 0; a < 7; a++) { step[a] = -b[a]; if (!std::isfinite(step[a])) ok = false; }
    if (ok) {
      for (int bb = 0; bb < m; bb++)
        for (int row = 0; row < 2; row++) {
          double mm = 0;
          for (int a = 0; a < 7; a++) mm += J[14 * (size_t)bb + 7 * row + a] * scale[a] * step[a];
          mcc -= mm * (r[2 * (size_t)bb + row] + mm / 2.0);
        }
      if (!(mcc > 0.0)) ok = false;
    }
    if (!ok) {
      consecutive_invalid++;
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
      if (consecutive_invalid >= 5) { sum.termination = 5; break; }
      if (iteration >= max_iterations) { sum.termination = 0; break; }
      if (radius < 1e-32) { sum.termination = 6; break; }
      continue;
    }
    consecutive_invalid = 0;
    for (int a = 0; a < 7; a++) delta[a] = step[a] * scale[a];
    double cand[7];
    plus(x, delta, cand);
    const double cand_cost = cost_only(cand);
    double sn = 0;
    for (int a = 0; a < 7; a++) sn += (x[a] - cand[a]) * (x[a] - cand[a]);
    const double step_norm = std::sqrt(sn);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { sum.termination = 2; put_trace(x_cost, 0, gmax, step_norm, 0, radius, 0, 1); break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) { sum.termination = 1; put_trace(x_cost, cost_change, gmax, step_norm, 0, radius, 0, 1); break; }
    const double rd = cost_change / mcc;
    const bool accepted = rd > 1e-3;
    if (accepted) {
      std::memcpy(x, cand, sizeof(x));
      evaluate(false);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rd - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      sum.successful_steps++;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0;
    }
    put_trace(x_cost, cost_change, gmax, step_norm, rd, radius, accepted ? 1 : 0, 1);
    if (iteration >= max_iterations) { sum.termination = 0; break; }
    if (gmax <= 1e-10) { sum.termination = 3; break; }
    if (radius < 1e-32) { sum.termination = 6; break; }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  if (out) *out = sum;
  if (lie7) std::memcpy(lie7, x, sizeof(x));
  // S12 = exp(sim12); outlier scan with Eigen::Quaterniond(s * R) fed to CheckOutlier (:702-726)
  const Sim3 S = exp(x), Si = inverse(S);
  *s12 = S.s; rotation_matrix(S.q, R12); std::memcpy(t12, S.t, 24);
  auto check = [&](const Sim3& T, const double* K4, const float* obs, float inv_sigma, const double* P) {
    double R[9], sR[9], q[4], p[3];
    rotation_matrix(T.q, R);
    for (int i = 0; i < 9; i++) sR[i] = T.s * R[i];
    quat_from_matrix(sR, q);                       // Eigen's conversion applied to a scaled rotation: not a unit quaternion
    quat_rot(q, P, p);                             // Eigen's _transformVector polynomial, unit-norm not enforced
    for (int i = 0; i < 3; i++) p[i] += T.t[i];
    const double px = K4[0] * p[0] + K4[2] * p[2], py = K4[1] * p[1] + K4[3] * p[2], pz = p[2];
    const double eu = obs[0] - px / pz, ev = obs[1] - py / pz;
    return (eu * eu + ev * ev) * inv_sigma > huber_a * huber_a;
  };
  int n_bad = 0;
  for (int i = 0; i < n; i++) {
    const bool b12 = check(S, K1, obs1 + 2 * i, inv_sigma1[i], P3D2c + 3 * i);
    const bool b21 = check(Si, K2, obs2 + 2 * i, inv_sigma2[i], P3D1c + 3 * i);
    is_bad[i] = b12 || b21;
    n_bad += is_bad[i];
  }
  if (n - n_bad < 10) return 0;
  return n - n_bad;
}

}  // extern "C"

// ---- OptimizeEssentialGraph (src/CeresOptimizer.cc:736-957, include/CeresOptimizer.h:270-328) -------------------------
// 7-DoF pose graph over the keyframes' Sim3 logs after a loop closure.  Flattened view (the pointer-graph traversal of
// :793-895 that picks the edges is the adapter's job, include/orb_slam2/CeresOptimizer.h):
//   Scw[n_kf][13]   scale, rotation (row-major), translation of the INITIAL Scw of every keyframe: the corrected Sim3 where
//                   keyframes_corrected_sim3 holds the keyframe (:767-769), else (1, Rcw, tcw) (:770-774)
//   kf_flags[n_kf]  bit 0: constant block (the loop keyframe, :780-783); bit 1: Snc holds keyframes_non_corrected_sim3's entry
//   Snc[n_kf][13]   non-corrected Siw
//   edges           residual block (parameter 0 = keyframe j, parameter 1 = keyframe i) in the reference's insertion order;
//                   kind 0 = loop-connection edge: Sji = exp(lie_j) * exp(lie_i)^-1 (:797-809);
//                   kind 1 = spanning-tree / loop / co-visibility edge: each side's non-corrected Sim3 where present, else
//                   exp(lie) (:822-895)
// Solve: trust-region LM as everywhere else (no loss, identity local Jacobian, Sim3Parameterization::Plus), 100 iterations.
// Afterwards (:903-956): Tiw = [R | t / s] per keyframe (4x4 row-major) and every map point moved through its reference
// keyframe: X' = exp(lie_r)^-1 * (exp(lie0_r) * X).
namespace sim3o {
void adjoint(const Sim3& S, double* A /*7x7 row-major*/) {            // Sophus Sim3::Adj()
  double R[9], T[9], TR[9];
  rotation_matrix(S.q, R); hat(S.t, T); mat3mul(T, R, TR);
  std::fill(A, A + 49, 0.0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[7 * i + j] = S.s * R[3 * i + j]; A[7 * i + 3 + j] = TR[3 * i + j]; A[7 * (3 + i) + 3 + j] = R[3 * i + j]; }
  for (int i = 0; i < 3; i++) A[7 * i + 6] = -S.t[i];
  A[48] = 1.0;
}
// EssentialGraphErrorTerm::Evaluate with sqrt_information = I: residual (7) and Jacobian wrt keyframe i (7x7 row-major);
// the Jacobian wrt keyframe j is its negative (:301-302)
void edge_term(const Sim3& Sji, const double* lie_j, const double* lie_i, double* r, double* Ji) {
  const Sim3 Si = exp(lie_i), Sj = exp(lie_j);
  log(mul(mul(Sji, Si), inverse(Sj)), r);
  if (!Ji) return;
  double A[49] = {0}, A2[49], Jr[49], Adj[49];
  double Ow[9], Ou[9];
  hat(r + 3, Ow); hat(r, Ou);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[7 * i + j] = Ow[3 * i + j] + (i == j ? r[6] : 0.0);       // RxSO3::hat(omega, sigma)
      A[7 * i + 3 + j] = Ou[3 * i + j];
      A[7 * (3 + i) + 3 + j] = Ow[3 * i + j];
    }
  for (int i = 0; i < 3; i++) A[7 * i + 6] = -r[i];
  for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) { double a = 0; for (int k = 0; k < 7; k++) a += A[7 * i + k] * A[7 * k + j]; A2[7 * i + j] = a; }
  for (int i = 0; i < 49; i++) Jr[i] = ((i % 8 == 0) ? 1.0 : 0.0) + 0.5 * A[i] + 1.0 / 12. * A2[i];
  adjoint(Sj, Adj);
  for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) { double a = 0; for (int k = 0; k < 7; k++) a += Jr[7 * i + k] * Adj[7 * k + j]; Ji[7 * i + j] = a; }
}
Sim3 from_srt(const double* v13) { Sim3 S; S.s = v13[0]; quat_from_matrix(v13 + 1, S.q); std::memcpy(S.t, v13 + 10, 24); return S; }
}  // namespace sim3o

extern "C" {

void ba_oracle_sim3_adjoint(const double* lie, double* A49) { sim3o::adjoint(sim3o::exp(lie), A49); }
void ba_oracle_essential_edge(const double* Sji13, const double* lie_j, const double* lie_i, double* r7, double* Ji49) {
  sim3o::edge_term(sim3o::from_srt(Sji13), lie_j, lie_i, r7, Ji49);
}

int ba_oracle_essential_graph(int n_kf, const double* Scw, const uint8_t* kf_flags, const double* Snc, int n_edges,
                              const int32_t* edge_j, const int32_t* edge_i, const uint8_t* edge_kind, int max_iterations,
                              int n_points, const double* Xw, const int32_t* ref_kf, double* lie_out, double* Tiw_out,
                              double* Xw_out, ba_oracle_summary* out, double* trace, int trace_cap) {
  using namespace sim3o;
  std::vector<double> x(7 * (size_t)n_kf), x0;
  for (int k = 0; k < n_kf; k++) log(from_srt(Scw + 13 * k), &x[7 * k]);
  x0 = x;
  std::vector<int> var(n_kf, -1);
  int Kv = 0;
  for (int k = 0; k < n_kf; k++) if (!(kf_flags[k] & 1)) var[k] = Kv++;
  const int n = 7 * Kv;
  // the measurements, from the initial values (:797-809, :822-895)
  std::vector<Sim3> meas(n_edges);
  for (int e = 0; e < n_edges; e++) {
    const int j = edge_j[e], i = edge_i[e];
    Sim3 Sjw, Swi;
    if (edge_kind[e] == 0) { Sjw = exp(&x[7 * j]); Swi = inverse(exp(&x[7 * i])); }
    else {
      Swi = (kf_flags[i] & 2) ? inverse(from_srt(Snc + 13 * i)) : inverse(exp(&x[7 * i]));
      Sjw = (kf_flags[j] & 2) ? from_srt(Snc + 13 * j) : exp(&x[7 * j]);
    }
    meas[e] = mul(Sjw, Swi);
  }
  ba_oracle_summary sum{};
  // Normal equations in row-envelope (skyline) storage: row i holds columns fc[i]..i.  Cholesky fill stays inside the row
  // envelope, so this is an exact sparse factorisation for the essential graph (a band from the spanning tree and the
  // co-visibility edges plus a few full rows from the loop edges) — the role CHOLMOD plays under SPARSE_NORMAL_CHOLESKY.
  // BA_ORACLE_EG_DENSE=1 widens every row to column 0 (a dense factorisation, used by the tests as a cross-check).
  std::vector<int> fc(std::max(n, 1), 0);
  {
    std::vector<int> first_blk(std::max(Kv, 1));
    for (int a = 0; a < Kv; a++) first_blk[a] = a;
    for (int e = 0; e < n_edges; e++) {
      const int vi = var[edge_i[e]], vj = var[edge_j[e]];
      if (vi >= 0 && vj >= 0) first_blk[std::max(vi, vj)] = std::min(first_blk[std::max(vi, vj)], std::min(vi, vj));
    }
    const char* dense = std::getenv("BA_ORACLE_EG_DENSE");
    for (int i = 0; i < n; i++) fc[i] = (dense && dense[0] == '1') ? 0 : 7 * first_blk[i / 7];
  }
  std::vector<size_t> rp(n + 1, 0);
  for (int i = 0; i < n; i++) rp[i + 1] = rp[i] + (size_t)(i - fc[i] + 1);
  std::vector<double> r(7 * (size_t)n_edges), J(49 * (size_t)n_edges), H(std::max<size_t>(rp[n], 1)), grad(n), scale(n, 1.0);
  auto Hat = [&](std::vector<double>& M, int i, int j) -> double& { return M[rp[i] + (size_t)(j - fc[i])]; };   // fc[i] <= j <= i
  double x_cost = 0, x_norm = 0, gmax = 0, radius = 1e4, decrease_factor = 2.0;
  int consecutive_invalid = 0, tr = 0;
  auto put_trace = [&](double cost, double dc, double gm, double sn, double rd, double rad, double acc, double valid) {
    if (trace && tr < trace_cap) { double* t = trace + 8 * tr; t[0] = cost; t[1] = dc; t[2] = gm; t[3] = sn; t[4] = rd; t[5] = rad; t[6] = acc; t[7] = valid; }
    tr++;
  };
  auto cost_only = [&](const std::vector<double>& y) {
    double c = 0;
    for (int e = 0; e < n_edges; e++) {
      double rr[7];
      edge_term(meas[e], &y[7 * edge_j[e]], &y[7 * edge_i[e]], rr, nullptr);
      double s = 0; for (int a = 0; a < 7; a++) s += rr[a] * rr[a];
      c += 0.5 * s;
    }
    return c;
  };
  auto evaluate = [&](bool first) {
    double c = 0;
    std::fill(H.begin(), H.end(), 0.0); std::fill(grad.begin(), grad.end(), 0.0);
    for (int e = 0; e < n_edges; e++) {
      double* rr = &r[7 * (size_t)e]; double* JJ = &J[49 * (size_t)e];
      edge_term(meas[e], &x[7 * edge_j[e]], &x[7 * edge_i[e]], rr, JJ);
      double s = 0; for (int a = 0; a < 7; a++) s += rr[a] * rr[a];
      c += 0.5 * s;
      const int vi = var[edge_i[e]], vj = var[edge_j[e]];
      double A[49], v[7];
      for (int a = 0; a < 7; a++) {
        double g = 0; for (int k = 0; k < 7; k++) g += JJ[7 * k + a] * rr[k];
        v[a] = g;
        for (int b = 0; b < 7; b++) { double h = 0; for (int k = 0; k < 7; k++) h += JJ[7 * k + a] * JJ[7 * k + b]; A[7 * a + b] = h; }
      }
      for (int a = 0; a < 7; a++) {
        if (vi >= 0) grad[7 * vi + a] += v[a];
        if (vj >= 0) grad[7 * vj + a] -= v[a];
        for (int b = 0; b < 7; b++) {
          if (vi >= 0 && b <= a) Hat(H, 7 * vi + a, 7 * vi + b) += A[7 * a + b];
          if (vj >= 0 && b <= a) Hat(H, 7 * vj + a, 7 * vj + b) += A[7 * a + b];
          if (vi >= 0 && vj >= 0 && vi != vj) { const int hi = std::max(vi, vj), lo = std::min(vi, vj); Hat(H, 7 * hi + a, 7 * lo + b) -= A[7 * a + b]; }
        }
      }
    }
    x_cost = c;
    sum.jacobian_evaluations++;
    if (first) for (int a = 0; a < n; a++) scale[a] = 1.0 / (1.0 + std::sqrt(Hat(H, a, a)));
    gmax = 0; double xn = 0;
    for (int k = 0; k < n_kf; k++) {
      if (var[k] < 0) continue;
      double ng[7], xp[7];
      for (int a = 0; a < 7; a++) ng[a] = -grad[7 * var[k] + a];
      plus(&x[7 * k], ng, xp);
      for (int a = 0; a < 7; a++) { gmax = std::max(gmax, std::fabs(x[7 * k + a] - xp[a])); xn += x[7 * k + a] * x[7 * k + a]; }
    }
    x_norm = std::sqrt(xn);
  };
  evaluate(true);
  sum.initial_cost = x_cost;
  put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
  int iteration = 0;
  sum.termination = 0;
  if (max_iterations == 0 || n == 0) {}
  else if (gmax <= 1e-10) sum.termination = 3;
  else for (;;) {
    iteration++;
    std::vector<double> A(H.size()), b(n), D2(n);
    for (int a = 0; a < n; a++) {
      D2[a] = std::min(std::max(scale[a] * scale[a] * Hat(H, a, a), 1e-6), 1e32) / radius;
      b[a] = scale[a] * grad[a];
      for (int d = fc[a]; d <= a; d++) Hat(A, a, d) = scale[a] * scale[d] * Hat(H, a, d) + (a == d ? D2[a] : 0.0);
    }
    // envelope Cholesky A = L L' in place, then L y = b, L' x = y
    bool ok = true;
    for (int i = 0; i < n && ok; i++) {
      for (int j = fc[i]; j <= i; j++) {
        double sacc = Hat(A, i, j);
        for (int k = std::max(fc[i], fc[j]); k < j; k++) sacc -= Hat(A, i, k) * Hat(A, j, k);
        if (j < i) Hat(A, i, j) = sacc / Hat(A, j, j);
        else { if (!(sacc > 0.0) || !std::isfinite(sacc)) { ok = false; break; } Hat(A, i, i) = std::sqrt(sacc); }
      }
    }
    if (ok) {
      for (int i = 0; i < n; i++) { double sacc = b[i]; for (int k = fc[i]; k < i; k++) sacc -= Hat(A, i, k) * b[k]; b[i] = sacc / Hat(A, i, i); }
      for (int i = n - 1; i >= 0; i--) { b[i] /= Hat(A, i, i); for (int k = fc[i]; k < i; k++) b[k] -= Hat(A, i, k) * b[i]; }
    }
    std::vector<double> delta(n);
    double mcc = 0.0;
    for (int a = 0; a < n; a++) { delta[a] = -b[a] * scale[a]; if (!std::isfinite(delta[a])) ok = false; }
    if (ok) {
      for (int e = 0; e < n_edges; e++) {
        const int vi = var[edge_i[e]], vj = var[edge_j[e]];
        for (int row = 0; row < 7; row++) {
          double mm = 0;
          for (int a = 0; a < 7; a++) {
            const double d = (vi >= 0 ? delta[7 * vi + a] : 0.0) - (vj >= 0 ? delta[7 * vj + a] : 0.0);
            mm += J[49 * (size_t)e + 7 * row + a] * d;
          }
          mcc -= mm * (r[7 * (size_t)e + row] + mm / 2.0);
        }
      }
      if (!(mcc > 0.0)) ok = false;
    }
    if (!ok) {
      consecutive_invalid++;
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      put_trace(x_cost, 0, gmax, 0, 0, radius, 0, 0);
      if (consecutive_invalid >= 5) { sum.termination = 5; break; }
      if (iteration >= max_iterations) { sum.termination = 0; break; }
      if (radius < 1e-32) { sum.termination = 6; break; }
      continue;
    }
    consecutive_invalid = 0;
    std::vector<double> cand(x);
    double sn = 0;
    for (int k = 0; k < n_kf; k++) {
      if (var[k] < 0) continue;
      plus(&x[7 * k], &delta[7 * var[k]], &cand[7 * k]);
      for (int a = 0; a < 7; a++) sn += (x[7 * k + a] - cand[7 * k + a]) * (x[7 * k + a] - cand[7 * k + a]);
    }
    const double cand_cost = cost_only(cand);
    const double step_norm = std::sqrt(sn);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { sum.termination = 2; put_trace(x_cost, 0, gmax, step_norm, 0, radius, 0, 1); break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) { sum.termination = 1; put_trace(x_cost, cost_change, gmax, step_norm, 0, radius, 0, 1); break; }
    const double rd = cost_change / mcc;
    const bool accepted = rd > 1e-3;
    if (accepted) {
      x = cand;
      evaluate(false);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rd - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      sum.successful_steps++;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0;
    }
    put_trace(x_cost, cost_change, gmax, step_norm, rd, radius, accepted ? 1 : 0, 1);
    if (iteration >= max_iterations) { sum.termination = 0; break; }
    if (gmax <= 1e-10) { sum.termination = 3; break; }
    if (radius < 1e-32) { sum.termination = 6; break; }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  if (out) *out = sum;
  if (lie_out) std::memcpy(lie_out, x.data(), x.size() * sizeof(double));
  std::vector<Sim3> Swc(n_kf);
  for (int k = 0; k < n_kf; k++) {
    const Sim3 S = exp(&x[7 * k]);
    Swc[k] = inverse(S);
    if (Tiw_out) {
      double R[9]; rotation_matrix(S.q, R);
      double* T = Tiw_out + 16 * k;
      for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = R[3 * i + j]; T[4 * i + 3] = (1. / S.s) * S.t[i]; }
      T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    }
  }
  for (int p = 0; p < n_points; p++) {
    const int rk = ref_kf[p];
    const Sim3 Srw = exp(&x0[7 * rk]);
    double Pc[3];
    act(Srw, Xw + 3 * p, Pc);
    act(Swc[rk], Pc, Xw_out + 3 * p);
  }
  return 0;
}

}  // extern "C"
