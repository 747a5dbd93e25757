// Drop-in for ORB_SLAM2::CeresOptimizer::{PoseOptimization, LocalBundleAdjustment, BundleAdjustment,
// GlobalBundleAdjustemnt, CheckOutlier} (include/CeresOptimizer.h:353-387) over POD views.  Static methods like
// the reference; each calling thread gets its own engine context (SURVEY.md §8b "Threading").
// OptimizeSim3 / OptimizeEssentialGraph are covered as well.
#ifndef ORB_SLAM2_CMOS_CERESOPTIMIZER_H
#define ORB_SLAM2_CMOS_CERESOPTIMIZER_H

#include <algorithm>
#include <set>
#include <utility>

#include "views.h"

namespace ORB_SLAM2 {

class CeresOptimizer {
 public:
  // Sizes of the per-thread engine; call once before the first solve if the defaults are too small.
  // max_pairs_per_obs sizes the co-observation pair lists (sum over points of L (L + 1) / 2 for track length L, divided by
  // the observation count; 0 = the library default 8).  It is only the INITIAL size: cmos_ba_set_problem grows the lists
  // when a graph with longer tracks arrives, so a dense local window or a global BA after many revisits does not throw.
  static void Configure(int max_keyframes, int max_points, int max_obs, int max_correspondences = 4096, int device = 0,
                        int max_pairs_per_obs = 0) {
    cmos_ba_params& p = params();
    p.max_cams = max_keyframes; p.max_points = max_points; p.max_obs = max_obs; p.max_pairs_per_obs = max_pairs_per_obs;
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

  // ---- graph collection (host only, no device call: testable without a GPU) --------------------------------------------
  // LocalBundleAdjustment(keyframe, stop_flag, map), CeresOptimizer.cc:349-406 and :421-502:
  //   local keyframes  = the keyframe, then GetVectorCovisibleKeyFrames() that are not bad (a bad neighbour is still MARKED
  //                      local, :358-359, so it can never become a fixed keyframe);
  //   local map points = non-null, non-bad GetMapPointMatches() of the local keyframes, first occurrence wins (:374-378);
  //   fixed keyframes  = observers of local points that are neither marked local nor already marked fixed, if not bad (:394-401);
  //   residual blocks  = per local point every observation whose keyframe is not bad and is local or fixed (:433, :460, :483);
  //   constant         = fixed keyframes, and keyframe id 0 (:476-481).
  // Where the reference iterates an unordered_map / map keyed by pointer, this uses first-insertion order.
  static void CollectLocalGraph(const MapView& m, int keyframe, CollectedGraph& out) {
    out = CollectedGraph();
    std::vector<int> kf_row(m.n_keyframes, -1), pt_row(m.n_points, -1);
    std::vector<uint8_t> marked_local(m.n_keyframes, 0), marked_fixed(m.n_keyframes, 0);
    auto add_keyframe = [&](int k, bool local) {
      kf_row[k] = (int)out.keyframe_index.size();
      out.keyframe_index.push_back(k);
      out.pose7.insert(out.pose7.end(), m.keyframe_pose7 + 7 * (size_t)k, m.keyframe_pose7 + 7 * (size_t)k + 7);
      out.flags.push_back((uint8_t)((local ? 0 : 2) | ((!local || m.keyframe_id[k] == 0) ? 1 : 0)));
    };
    add_keyframe(keyframe, true);
    marked_local[keyframe] = 1;
    for (size_t i = 0; i < m.covisible[keyframe].size(); i++) {
      const int nb = m.covisible[keyframe][i];
      const bool first = !marked_local[nb];
      marked_local[nb] = 1;
      if (!m.keyframe_bad[nb] && first) add_keyframe(nb, true);
    }
    const int n_local = (int)out.keyframe_index.size();
    for (int r = 0; r < n_local; r++) {
      const std::vector<int>& pts = m.keyframe_points[out.keyframe_index[r]];
      for (size_t i = 0; i < pts.size(); i++) {
        const int p = pts[i];
        if (p < 0 || m.point_bad[p] || pt_row[p] >= 0) continue;
        pt_row[p] = (int)out.point_index.size();
        out.point_index.push_back(p);
        out.point_pos.insert(out.point_pos.end(), m.point_pos + 3 * (size_t)p, m.point_pos + 3 * (size_t)p + 3);
      }
    }
    for (size_t r = 0; r < out.point_index.size(); r++) {
      const std::vector<std::pair<int, int> >& obs = m.observations[out.point_index[r]];
      for (size_t o = 0; o < obs.size(); o++) {
        const int k = obs[o].first;
        if (marked_local[k] || marked_fixed[k]) continue;
        marked_fixed[k] = 1;
        if (!m.keyframe_bad[k]) add_keyframe(k, false);
      }
    }
    add_observations(m, kf_row, 0, false, out);
  }

  // BundleAdjustment(map->GetAllKeyFrames(), map->GetAllMapPoints(), ...), CeresOptimizer.cc:53-57, 88-175: every keyframe
  // that is not bad (id 0 constant, :115-120), every map point that is not bad and keeps at least one observation by a
  // non-bad keyframe with id <= the largest id among them (:141, :168-173 removes points without edges).
  static void CollectGlobalGraph(const MapView& m, CollectedGraph& out) {
    out = CollectedGraph();
    std::vector<int> kf_row(m.n_keyframes, -1);
    unsigned long max_id = 0;
    for (int k = 0; k < m.n_keyframes; k++) {
      if (m.keyframe_bad[k]) continue;
      kf_row[k] = (int)out.keyframe_index.size();
      out.keyframe_index.push_back(k);
      out.pose7.insert(out.pose7.end(), m.keyframe_pose7 + 7 * (size_t)k, m.keyframe_pose7 + 7 * (size_t)k + 7);
      out.flags.push_back(m.keyframe_id[k] == 0 ? 1 : 0);
      max_id = std::max(max_id, m.keyframe_id[k]);
    }
    for (int p = 0; p < m.n_points; p++) {
      if (m.point_bad[p]) continue;
      out.point_index.push_back(p);
      out.point_pos.insert(out.point_pos.end(), m.point_pos + 3 * (size_t)p, m.point_pos + 3 * (size_t)p + 3);
    }
    add_observations(m, kf_row, max_id, true, out);
  }

  // Write-back of LocalBundleAdjustment (CeresOptimizer.cc:567-598): the (keyframe, map point) pairs to erase on both
  // sides (EraseMapPointMatch / EraseObservation), the local keyframes' poses (SetPose) and the local points' positions
  // (SetWorldPos + UpdateNormalAndDepth), all as map indices.  Fixed keyframes are not written.
  struct LocalResult {
    std::vector<std::pair<int, int> > erase;               // (map keyframe index, map point index)
    std::vector<int> keyframes, points;                    // map indices, rows of pose7 / pos below
    std::vector<double> pose7, pos;
  };
  static LocalResult CollectLocalResult(const CollectedGraph& c, const GraphView& g) {
    LocalResult r;
    for (int o = 0; o < g.n_obs; o++)
      if (!g.erase.empty() && g.erase[o])
        r.erase.push_back(std::make_pair(c.keyframe_index[c.obs_keyframe[o]], c.point_index[c.obs_point[o]]));
    for (size_t k = 0; k < c.keyframe_index.size(); k++) {
      if (c.flags[k] & 2) continue;
      r.keyframes.push_back(c.keyframe_index[k]);
      r.pose7.insert(r.pose7.end(), c.pose7.begin() + 7 * k, c.pose7.begin() + 7 * k + 7);
    }
    r.points = c.point_index;
    r.pos = c.point_pos;
    return r;
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

  // returns n_correspondences - n_bad, 0 when fewer than 10 remain (CeresOptimizer.cc:731-734); S12 in/out
  static int OptimizeSim3(Sim3MatchesView& m, Sim3POD& S12, const float th2, const bool /*bFixScale: unused by the reference*/) {
    m.is_bad.assign(std::max(m.n, 1), 0);
    int32_t inliers = 0;
    cmos_throw_if(cmos_ba_optimize_sim3(ctx(), m.n, &S12.s, S12.R, S12.t, m.K1, m.K2, m.obs1, m.inv_sigma1, m.P3D2c, m.obs2,
                                        m.inv_sigma2, m.P3D1c, th2, 100, m.is_bad.data(), nullptr, &inliers, nullptr),
                  "CeresOptimizer::OptimizeSim3");
    m.is_bad.resize(m.n);
    return inliers;
  }

  // The residual blocks of OptimizeEssentialGraph in the reference's insertion order (CeresOptimizer.cc:793-895): block 0 of
  // edge e is keyframe ej[e], block 1 keyframe ei[e]; ek[e] = 0 for a loop-connection edge (:797-809), 1 otherwise.  Pure
  // host logic (no device call), so it is testable without a GPU.
  static void CollectEssentialGraphEdges(const EssentialGraphView& g, std::vector<int32_t>& ej, std::vector<int32_t>& ei,
                                         std::vector<uint8_t>& ek) {
    const int min_weight = 100;
    const int n = g.n_keyframes;
    ej.clear(); ei.clear(); ek.clear();
    std::set<std::pair<unsigned long, unsigned long> > inserted;
    for (int i = 0; i < n; i++)
      for (size_t c = 0; c < g.loop_connections[i].size(); c++) {
        const int j = g.loop_connections[i][c];
        if ((i != g.current_keyframe || j != g.loop_keyframe) && g.loop_connection_weight[i][c] < min_weight) continue;
        ej.push_back(j); ei.push_back(i); ek.push_back(0);
        inserted.insert(std::make_pair(std::min(g.id[i], g.id[j]), std::max(g.id[i], g.id[j])));
      }
    for (int i = 0; i < n; i++) {
      const int parent = g.parent[i];
      if (parent >= 0) { ej.push_back(parent); ei.push_back(i); ek.push_back(1); }
      const std::vector<int>& le = g.loop_edges[i];
      for (size_t c = 0; c < le.size(); c++)
        if (g.id[le[c]] < g.id[i]) { ej.push_back(le[c]); ei.push_back(i); ek.push_back(1); }
      for (size_t c = 0; c < g.covisibles[i].size(); c++) {
        const int kn = g.covisibles[i][c];
        if (kn < 0 || kn == parent) continue;
        if (std::find(g.children[i].begin(), g.children[i].end(), kn) != g.children[i].end()) continue;
        if (std::find(le.begin(), le.end(), kn) != le.end()) continue;
        if (!(g.id[kn] < g.id[i])) continue;
        if (inserted.count(std::make_pair(std::min(g.id[i], g.id[kn]), std::max(g.id[i], g.id[kn])))) continue;
        ej.push_back(kn); ei.push_back(i); ek.push_back(1);
      }
    }
  }

  // Collects the residual blocks exactly as CeresOptimizer.cc:793-895 does (loop connections first, then per keyframe its
  // parent, its older loop edges and its older co-visible keyframes that are neither parent, child, loop edge nor already
  // inserted), solves on the device and fills g.Tiw / g.corrected_pos.  Where the reference iterates a std::set /
  // std::map keyed by KeyFrame* (heap-address order) the index order of the view is used.
  static void OptimizeEssentialGraph(EssentialGraphView& g, const bool& /*is_fixed_scale: unused by the reference*/ = false) {
    const int n = g.n_keyframes;
    std::vector<double> Scw(13 * (size_t)n), Snc(13 * (size_t)n, 0.0);
    std::vector<uint8_t> flags(n, 0);
    for (int k = 0; k < n; k++) {
      const Sim3POD& S = g.has_corrected[k] ? g.corrected[k] : g.pose[k];
      Scw[13 * k] = S.s; std::copy(S.R, S.R + 9, &Scw[13 * k + 1]); std::copy(S.t, S.t + 3, &Scw[13 * k + 10]);
      if (g.has_non_corrected[k]) {
        const Sim3POD& N = g.non_corrected[k];
        Snc[13 * k] = N.s; std::copy(N.R, N.R + 9, &Snc[13 * k + 1]); std::copy(N.t, N.t + 3, &Snc[13 * k + 10]);
        flags[k] |= 2;
      }
      if (k == g.loop_keyframe) flags[k] |= 1;
    }
    std::vector<int32_t> ej, ei;
    std::vector<uint8_t> ek;
    CollectEssentialGraphEdges(g, ej, ei, ek);
    g.Tiw.assign(16 * (size_t)n, 0.0);
    g.corrected_pos.assign(3 * (size_t)std::max(g.n_points, 1), 0.0);
    cmos_throw_if(cmos_ba_optimize_essential_graph(ctx(), n, Scw.data(), flags.data(), Snc.data(), (int32_t)ej.size(), ej.data(),
                                                   ei.data(), ek.data(), 100, g.n_points, g.point_pos, g.point_ref, nullptr,
                                                   g.Tiw.data(), g.corrected_pos.data(), nullptr),
                  "CeresOptimizer::OptimizeEssentialGraph");
    g.corrected_pos.resize(3 * (size_t)g.n_points);
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

  // Destroys the calling thread's engine now.  It is also destroyed when the thread exits (the slot is an RAII holder), so
  // the reference's pattern of one std::thread per global bundle adjustment (LoopClosing.cc:590) does not leak a handle —
  // at 1000 keyframes that is a 288 MB reduced-system buffer per loop closure.
  static void release() { slot().reset(); }

 private:
  // residual blocks in the reference's order: per collected point, its GetObservations() in map order.  Points that end
  // up without any block are dropped (global: RemoveParameterBlock, :168-170; local: such a point has no residual either).
  static void add_observations(const MapView& m, const std::vector<int>& kf_row, unsigned long max_id, bool use_max_id,
                               CollectedGraph& out) {
    std::vector<int> keep_index;
    std::vector<double> keep_pos;
    for (size_t r = 0; r < out.point_index.size(); r++) {
      const int p = out.point_index[r];
      const std::vector<std::pair<int, int> >& obs = m.observations[p];
      int n_edges = 0;
      for (size_t o = 0; o < obs.size(); o++) {
        const int k = obs[o].first, kp = obs[o].second;
        if (m.keyframe_bad[k] || kf_row[k] < 0) continue;
        if (use_max_id && m.keyframe_id[k] > max_id) continue;
        const KeyPoint& u = m.undistort_keypoints[k][kp];
        out.obs_keyframe.push_back(kf_row[k]);
        out.obs_point.push_back((int32_t)keep_index.size());
        out.obs_keypoint.push_back(kp);
        out.obs_uv.push_back(u.x); out.obs_uv.push_back(u.y);
        out.obs_inv_sigma2.push_back(m.inv_level_sigma2[k][u.octave]);
        n_edges++;
      }
      if (n_edges == 0) continue;
      keep_index.push_back(p);
      keep_pos.insert(keep_pos.end(), out.point_pos.begin() + 3 * r, out.point_pos.begin() + 3 * r + 3);
    }
    out.point_index.swap(keep_index);
    out.point_pos.swap(keep_pos);
  }

  static cmos_ba_params& params() {
    static cmos_ba_params p = {128, 16384, 131072, 0, 1, 4096, 0};
    return p;
  }
  struct Engine {                       // owns one cmos_ba handle; thread_local, so its destructor runs at thread exit
    cmos_ba_t h = nullptr;
    Engine() {}
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    ~Engine() { reset(); }
    void reset() { if (h) { cmos_ba_destroy(h); h = nullptr; } }
  };
  static Engine& slot() {
    static thread_local Engine e;
    return e;
  }
  static cmos_ba_t ctx() {
    Engine& e = slot();
    if (!e.h) cmos_throw_if(cmos_ba_create(&params(), &e.h), "CeresOptimizer engine");
    return e.h;
  }
};

}  // namespace ORB_SLAM2
#endif
