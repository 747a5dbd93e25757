// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.cpp header for the rules).
//
// CPU restatement of DBoW2's TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&,
// FeatureVector&, levelsup) as the reference calls it (Frame::ComputeBoW src/Frame.cc:322-327, KeyFrame::ComputeBoW
// src/KeyFrame.cc:107-117, levelsup = 4), from the vendored sources:
//   tree descent + word / node ids        /root/reference/lib/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1260
//   FORB::distance                        /root/reference/lib/DBoW2/DBoW2/FORB.cpp:81-101
//   BowVector::addWeight / normalize      /root/reference/lib/DBoW2/DBoW2/BowVector.cpp:34-84
//   FeatureVector::addFeature             /root/reference/lib/DBoW2/DBoW2/FeatureVector.cpp:31-45
// for the weighting / scoring of ORBvoc.txt (TF_IDF, L1_NORM: loadFromTextFile, TemplatedVocabulary.h:1330-1420).
// PINNED TO THE REFERENCE ITSELF: tests/test_ref_parity.py loads the same vocabulary into the reference's own vendored
// DBoW2 (compiled unmodified into oracle/_ref/libref.so, loadFromTextFile) and requires identical BowVector values (bit for
// bit) and FeatureVector; tests/test_oracle_matcher2.py adds a brute-force descent.
//
// The vocabulary tree is passed flattened: node 0 is the root; children of node i are
// children[child_start[i] .. child_start[i+1]) in stored order; a node without children is a leaf with a word id and
// a weight (idf).  Outputs: BowVector as ascending word ids + values, FeatureVector as CSR (ascending node ids).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {
inline int hamming256(const uint8_t* a, const uint8_t* b) {
  uint32_t pa[8], pb[8];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  int d = 0;
  for (int i = 0; i < 8; i++) d += __builtin_popcount(pa[i] ^ pb[i]);
  return d;
}
}  // namespace

extern "C" {

// returns the number of words; *n_fv_nodes = number of FeatureVector nodes
int bow_oracle_transform(int n_nodes, const int32_t* child_start, const int32_t* children, const uint8_t* node_desc,
                         const double* node_weight, const int32_t* node_word, int L, const uint8_t* features, int n,
                         int levelsup, int32_t* bow_words, double* bow_values, int32_t* fv_nodes, int32_t* fv_start,
                         int32_t* fv_features, int32_t* n_fv_nodes, int32_t* feat_word, int32_t* feat_node) {
  std::map<int, double> v;
  std::map<int, std::vector<int>> fv;
  const int nid_level = L - levelsup;
  for (int i = 0; i < n; i++) {
    const uint8_t* f = features + 32 * (size_t)i;
    int nid = 0;
    int final_id = 0, current_level = 0;
    do {
      ++current_level;
      const int c0 = child_start[final_id], c1 = child_start[final_id + 1];
      final_id = children[c0];
      double best_d = hamming256(f, node_desc + 32 * (size_t)final_id);
      for (int c = c0 + 1; c < c1; c++) {
        const int id = children[c];
        const double d = hamming256(f, node_desc + 32 * (size_t)id);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid = final_id;
    } while (child_start[final_id + 1] > child_start[final_id]);
    const int word = node_word[final_id];
    const double w = node_weight[final_id];
    if (feat_word) feat_word[i] = word;
    if (feat_node) feat_node[i] = nid;
    if (w > 0) {
      auto it = v.lower_bound(word);
      if (it != v.end() && it->first == word) it->second += w; else v.insert(it, {word, w});
      fv[nid].push_back(i);
    }
  }
  double norm = 0.0;                       // L1_NORM scoring: mustNormalize -> BowVector::normalize(L1)
  for (auto& kv : v) norm += std::fabs(kv.second);
  if (norm > 0.0)
    for (auto& kv : v) kv.second /= norm;
  int k = 0;
  for (auto& kv : v) { bow_words[k] = kv.first; bow_values[k] = kv.second; k++; }
  int m = 0, pos = 0;
  for (auto& kv : fv) {
    fv_nodes[m] = kv.first; fv_start[m] = pos;
    for (int i : kv.second) fv_features[pos++] = i;
    m++;
  }
  fv_start[m] = pos;
  *n_fv_nodes = m;
  return k;
}

}  // extern "C"
