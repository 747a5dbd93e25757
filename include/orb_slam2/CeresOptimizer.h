// Drop-in for ORB_SLAM2::CeresOptimizer::{PoseOptimization, LocalBundleAdjustment, BundleAdjustment,
// GlobalBundleAdjustemnt, CheckOutlier} (include/CeresOptimizer.h:353-387) over POD views.  Static methods like
// the reference; each calling thread gets its own engine context (SURVEY.md §8b "Threading").
// OptimizeSim3 / OptimizeEssentialGraph are "next" rows and stay on the reference's Ceres code.
#ifndef ORB_SLAM2_CMOS_CERESOPTIMIZER_H
#define ORB_SLAM2_CMOS_CERESOPTIMIZER_H

#include "views.h"

namespace ORB_SLAM2 {

class CeresOptimizer {
 public:
  // sizes of the per-thread engine; call once before the first solve if the defaults are too small
  static void Configure(int max_keyframes, int max_points, int max_obs, int max_correspondences = 4096, int device = 0) {
    cmos_ba_params& p = params();
    p.max_cams = max_keyframes; p.max_points = max_points; p.max_obs = max_obs; p.max_pairs_per_obs = 0;
    p.max_pose_batch = 1; p.max_pose_corr = max_correspondences; p.device = device;
    release();
  }

  static void BundleAdjustment(GraphView& g, int n_iterations = 200, bool* stop_flag = nullptr,
                               const unsigned long /*n_loop_keyframe*/ = 0, const bool is_robust = true) {
    cmos_ba_summary s;
    cmos_throw_if(cmos_ba_bundle_adjustment(ctx(), g.n_keyframes, g.keyframe_pose7, g.keyframe_flags, g.n_points,
                                            g.point_pos, g.n_obs, g.obs_keyframe, g.obs_point, g.obs_uv,
                                            g.obs_inv_sigma2, g.K4, n_iterations, is_robust ? 1 : 0,
                                            (const uint8_t*)stop_flag, &s), "CeresOptimizer::BundleAdjustment");
  }

  // [sic] the reference's spelling (CeresOptimizer.h:360)
  static void GlobalBundleAdjustemnt(GraphView& map, int n_iterations = 200, bool* stop_flag = nullptr,
                                     const unsigned long n_loop_keyframe = 0, const bool is_robust = true) {
    BundleAdjustment(map, n_iterations, stop_flag, n_loop_keyframe, is_robust);
  }

  static void LocalBundleAdjustment(GraphView& g, bool* stop_flag) {
    g.erase.assign(g.n_obs, 0);
    cmos_ba_summary s[2];
    cmos_throw_if(cmos_ba_local_bundle_adjustment(ctx(), g.n_keyframes, g.keyframe_pose7, g.keyframe_flags, g.n_points,
                                                  g.point_pos, g.n_obs, g.obs_keyframe, g.obs_point, g.obs_uv,
                                                  g.obs_inv_sigma2, g.K4, (const uint8_t*)stop_flag, g.erase.data(), s),
                  "CeresOptimizer::LocalBundleAdjustment");
  }

  // returns n_initial_correspondences - n_bad (CeresOptimizer.cc:341); fills frame.is_outlier, updates frame.pose7
  static int PoseOptimization(FramePoseView* frame) {
    frame->is_outlier.assign(frame->n, 0);
    if (frame->n < 3) return 0;
    int32_t n = frame->n, inliers = 0;
    cmos_throw_if(cmos_ba_pose_optimization(ctx(), 1, frame->pose7, &n, frame->world_pos, frame->uv, frame->inv_sigma2,
                                            frame->n, frame->K4, 100, frame->is_outlier.data(), &inliers, nullptr, 0,
                                            nullptr), "CeresOptimizer::PoseOptimization");
    return inliers;
  }

  // chi2 test of one observation (CeresOptimizer.cc:227-241); host arithmetic, it is 20 flops
  static bool CheckOutlier(const float K4[4], const float uv[2], float inv_sigma, const double X[3], const double tcw[3],
                           const double q_xyzw[4], double thres) {
    const double qx = q_xyzw[0], qy = q_xyzw[1], qz = q_xyzw[2], qw = q_xyzw[3];
    double u0 = qy * X[2] - qz * X[1], u1 = qz * X[0] - qx * X[2], u2 = qx * X[1] - qy * X[0];
    u0 += u0; u1 += u1; u2 += u2;
    const double p0 = X[0] + qw * u0 + (qy * u2 - qz * u1) + tcw[0], p1 = X[1] + qw * u1 + (qz * u0 - qx * u2) + tcw[1],
                 p2 = X[2] + qw * u2 + (qx * u1 - qy * u0) + tcw[2];
    const double eu = uv[0] - ((double)K4[0] * p0 + (double)K4[2] * p2) / p2;
    const double ev = uv[1] - ((double)K4[1] * p1 + (double)K4[3] * p2) / p2;
    return (eu * eu + ev * ev) * inv_sigma > thres;
  }

  static void release() {
    cmos_ba_t& h = slot();
    if (h) { cmos_ba_destroy(h); h = nullptr; }
  }

 private:
  static cmos_ba_params& params() {
    static cmos_ba_params p = {128, 16384, 131072, 0, 1, 4096, 0};
    return p;
  }
  static cmos_ba_t& slot() {
    static thread_local cmos_ba_t h = nullptr;
    return h;
  }
  static cmos_ba_t ctx() {
    cmos_ba_t& h = slot();
    if (!h) cmos_throw_if(cmos_ba_create(&params(), &h), "CeresOptimizer engine");
    return h;
  }
};

}  // namespace ORB_SLAM2
#endif
