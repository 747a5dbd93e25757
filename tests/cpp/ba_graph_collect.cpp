// CPU-only check of CeresOptimizer::CollectLocalGraph / CollectGlobalGraph / CollectLocalResult on a hand-worked map that
// exercises the rules of CeresOptimizer.cc:349-406, 421-502 (local) and :88-175 (global).  Prints the collected graphs;
// tests/test_cpp_adapters.py compares with the expected lists.
#include <cstdio>

#include "../../include/orb_slam2/CeresOptimizer.h"

using namespace ORB_SLAM2;

static void print_graph(const char* tag, const CollectedGraph& c) {
  std::printf("%s keyframes", tag);
  for (size_t k = 0; k < c.keyframe_index.size(); k++) std::printf(" %d:%d", c.keyframe_index[k], (int)c.flags[k]);
  std::printf("\n%s points", tag);
  for (size_t p = 0; p < c.point_index.size(); p++) std::printf(" %d", c.point_index[p]);
  std::printf("\n%s obs", tag);
  for (size_t o = 0; o < c.obs_keyframe.size(); o++)
    std::printf(" (%d,%d,%g,%g,%g)", c.keyframe_index[c.obs_keyframe[o]], c.point_index[c.obs_point[o]], c.obs_uv[2 * o],
                c.obs_uv[2 * o + 1], c.obs_inv_sigma2[o]);
  std::printf("\n");
}

int main() {
  // keyframes (index: id):  0:0  1:5  2:6  3:7 (bad)  4:8  5:9  6:2
  // current keyframe = 2.  covisible(2) = {1, 3, 0}: 3 is bad (marked local, never fixed), 0 has id 0 (local but constant).
  const int nk = 7, np = 7;
  MapView m;
  m.n_keyframes = nk; m.n_points = np;
  const unsigned long ids[nk] = {0, 5, 6, 7, 8, 9, 2};
  m.keyframe_id.assign(ids, ids + nk);
  m.keyframe_bad.assign(nk, 0); m.keyframe_bad[3] = 1;
  std::vector<double> pose(7 * nk, 0.0);
  for (int k = 0; k < nk; k++) { pose[7 * k] = k; pose[7 * k + 6] = 1.0; }
  m.keyframe_pose7 = pose.data();
  m.covisible.resize(nk); m.covisible[2] = {1, 3, 0};
  // keypoints: keyframe k has 4 keypoints at (100 k + i, 10 i), octave i % 2
  std::vector<std::vector<KeyPoint> > kps(nk, std::vector<KeyPoint>(4));
  const float inv_sigma2[2] = {1.0f, 0.5f};
  for (int k = 0; k < nk; k++)
    for (int i = 0; i < 4; i++) { KeyPoint u = {}; u.x = 100.f * k + i; u.y = 10.f * i; u.octave = i % 2; kps[k][i] = u; }
  for (int k = 0; k < nk; k++) { m.undistort_keypoints.push_back(kps[k].data()); m.inv_level_sigma2.push_back(inv_sigma2); }
  std::vector<double> pos(3 * np);
  for (int p = 0; p < np; p++) { pos[3 * p] = p; pos[3 * p + 1] = 2 * p; pos[3 * p + 2] = 10 + p; }
  m.point_pos = pos.data();
  m.point_bad.assign(np, 0); m.point_bad[4] = 1;
  // map point matches (per keypoint) and the observation lists (keyframe, keypoint)
  m.keyframe_points = {{0, 1, -1, -1}, {1, 2, -1, -1}, {2, 0, 4, 3}, {3, 5, -1, -1}, {2, 3, -1, -1}, {1, 6, -1, -1}, {5, -1, -1, -1}};
  m.observations.resize(np);
  m.observations[0] = {{0, 0}, {2, 1}};
  m.observations[1] = {{0, 1}, {1, 0}, {5, 0}};            // 5 sees a local point, is not local -> fixed
  m.observations[2] = {{1, 1}, {2, 0}, {4, 0}};            // 4 -> fixed
  m.observations[3] = {{2, 3}, {3, 0}, {4, 1}};            // 3 is bad: skipped; 4 already fixed
  m.observations[4] = {{2, 2}};                            // bad point: never collected
  m.observations[5] = {{3, 1}, {6, 0}};                    // seen by the bad keyframe and keyframe 6 only: not local
  m.observations[6] = {{5, 1}};                            // seen only by a fixed keyframe: not a local point
  m.K4[0] = 700; m.K4[1] = 700; m.K4[2] = 600; m.K4[3] = 180;
  CollectedGraph c;
  CeresOptimizer::CollectLocalGraph(m, 2, c);
  print_graph("local", c);
  // write-back: pretend observation 1 and the last one were flagged by the solve
  GraphView g = c.view(m.K4);
  g.erase.assign(g.n_obs, 0); g.erase[1] = 1; g.erase[g.n_obs - 1] = 1;
  CeresOptimizer::LocalResult r = CeresOptimizer::CollectLocalResult(c, g);
  std::printf("result erase");
  for (size_t i = 0; i < r.erase.size(); i++) std::printf(" (%d,%d)", r.erase[i].first, r.erase[i].second);
  std::printf("\nresult keyframes");
  for (size_t i = 0; i < r.keyframes.size(); i++) std::printf(" %d", r.keyframes[i]);
  std::printf("\nresult points");
  for (size_t i = 0; i < r.points.size(); i++) std::printf(" %d", r.points[i]);
  std::printf("\n");
  CeresOptimizer::CollectGlobalGraph(m, c);
  print_graph("global", c);
  return 0;
}
