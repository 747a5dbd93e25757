// CPU-only check of CeresOptimizer::CollectEssentialGraphEdges: prints "j i kind" per residual block for a small map that
// exercises every rule of CeresOptimizer.cc:793-895 (tests/test_cpp_adapters.py compares with the rules restated in Python).
#include <cstdio>

#include "../../include/orb_slam2/CeresOptimizer.h"

using namespace ORB_SLAM2;

int main() {
  const int n = 8;
  EssentialGraphView g;
  g.n_keyframes = n;
  g.id.resize(n); g.pose.resize(n); g.parent.assign(n, -1); g.children.resize(n); g.loop_edges.resize(n);
  g.covisibles.resize(n); g.loop_connections.resize(n); g.loop_connection_weight.resize(n);
  g.has_corrected.assign(n, 0); g.has_non_corrected.assign(n, 0); g.corrected.resize(n); g.non_corrected.resize(n);
  // ids are not the indices: index k has id 10 * k, except that index 2 and 3 are swapped in age
  const unsigned long ids[n] = {0, 10, 30, 20, 40, 50, 60, 70};
  for (int k = 0; k < n; k++) g.id[k] = ids[k];
  // spanning tree: 0 <- 1 <- 3 <- 2 <- 4 <- 5 <- 6 <- 7 (parent by age)
  const int parent[n] = {-1, 0, 3, 1, 2, 4, 5, 6};
  for (int k = 0; k < n; k++) { g.parent[k] = parent[k]; if (parent[k] >= 0) g.children[parent[k]].push_back(k); }
  g.loop_keyframe = 1; g.current_keyframe = 7;
  // loop connections: 7 -> {1 (weight 3: kept, it is the current/loop pair), 0 (weight 150: kept), 3 (weight 20: dropped)};
  //                   6 -> {1 (weight 120: kept), 0 (weight 99: dropped)}
  g.loop_connections[7] = {1, 0, 3}; g.loop_connection_weight[7] = {3, 150, 20};
  g.loop_connections[6] = {1, 0};    g.loop_connection_weight[6] = {120, 99};
  // an old loop edge between 5 and 0 (recorded on both), and one between 2 and 3 (3 is OLDER than 2 by id)
  g.loop_edges[5] = {0}; g.loop_edges[0] = {5};
  g.loop_edges[2] = {3}; g.loop_edges[3] = {2};
  // co-visibility lists (already filtered by weight >= 100), in their weight order
  g.covisibles[4] = {2, 3, 1, 5};      // 2 = parent (skip), 3 older (keep), 1 older (keep), 5 = child (skip)
  g.covisibles[5] = {0, 4, 3};         // 0 = loop edge (skip), 4 = parent (skip), 3 older (keep)
  g.covisibles[7] = {0, 6, 5};         // 0: pair (0,70) already inserted by the loop connections (skip), 6 parent, 5 keep
  g.covisibles[3] = {2, 1};            // 2 is YOUNGER (id 30 > 20): skip; 1 = parent
  g.covisibles[6] = {1, 7};            // (10,60) inserted (skip); 7 = child
  std::vector<int32_t> ej, ei;
  std::vector<uint8_t> ek;
  CeresOptimizer::CollectEssentialGraphEdges(g, ej, ei, ek);
  for (size_t e = 0; e < ej.size(); e++) std::printf("%d %d %d\n", ej[e], ei[e], (int)ek[e]);
  return 0;
}
