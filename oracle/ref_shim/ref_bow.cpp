// TEST INFRASTRUCTURE — C harness around the UNMODIFIED vendored DBoW2 of the reference
// (/root/reference/lib/DBoW2/DBoW2/{TemplatedVocabulary.h,FORB.cpp,BowVector.cpp,FeatureVector.cpp,ScoringObject.cpp}),
// instantiated exactly as the reference does (include/ORBVocabulary.h:30-31: TemplatedVocabulary<FORB::TDescriptor, FORB>)
// and loaded the way the reference loads ORBvoc.txt (loadFromTextFile, src/MonoORBSlam.cc).  Output shapes follow
// oracle/bow_oracle.cpp so tests can run the two side by side.
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
}

extern "C" {

void* ref_bow_load_text(const char* path) {
  ORBVocabulary* v = new ORBVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void ref_bow_destroy(void* h) { delete (ORBVocabulary*)h; }
int ref_bow_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }
int ref_bow_depth(void* h) { return ((ORBVocabulary*)h)->getDepthLevels(); }
int ref_bow_branching(void* h) { return ((ORBVocabulary*)h)->getBranchingFactor(); }

// Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:322-327, src/KeyFrame.cc:107-117): descriptors -> one 1x32 Mat per
// row, then transform(features, BowVector, FeatureVector, levelsup).  Returns the number of words.
int ref_bow_transform(void* h, const uint8_t* features, int n, int levelsup, int32_t* bow_words, double* bow_values,
                      int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features, int32_t* n_fv_nodes) {
  ORBVocabulary* voc = (ORBVocabulary*)h;
  std::vector<cv::Mat> desc(n);
  for (int i = 0; i < n; i++) {
    desc[i].create(1, 32, CV_8U);
    std::memcpy(desc[i].data, features + 32 * (size_t)i, 32);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc->transform(desc, bv, fv, levelsup);
  int k = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) {
    bow_words[k] = (int32_t)it->first; bow_values[k] = it->second;
  }
  int m = 0, pos = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++m) {
    fv_nodes[m] = (int32_t)it->first; fv_start[m] = pos;
    for (size_t j = 0; j < it->second.size(); j++) fv_features[pos++] = (int32_t)it->second[j];
  }
  fv_start[m] = pos;
  *n_fv_nodes = m;
  return k;
}

// ORBVocabulary::score (KeyFrameDatabase / LoopClosing use it on two BowVectors; L1 scoring for ORBvoc.txt).
double ref_bow_score(void* h, const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2) {
  DBoW2::BowVector a, b;
  for (int i = 0; i < n1; i++) a.insert(std::make_pair((DBoW2::WordId)w1[i], v1[i]));
  for (int i = 0; i < n2; i++) b.insert(std::make_pair((DBoW2::WordId)w2[i], v2[i]));
  return ((ORBVocabulary*)h)->score(a, b);
}

}  // extern "C"
