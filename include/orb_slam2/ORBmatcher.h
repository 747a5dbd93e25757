// Drop-in for ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:36-102) over POD views: all eleven searches plus
// DescriptorDistance run on the GPU behind the reference's method names, argument order, defaults and return
// values.  MapPoint* become indices; the map-mutating Fuse returns decisions the caller applies (views.h).
#ifndef ORB_SLAM2_CMOS_ORBMATCHER_H
#define ORB_SLAM2_CMOS_ORBMATCHER_H

#include <map>
#include <tuple>
#include <utility>

#include "views.h"

namespace ORB_SLAM2 {

class ORBmatcher {
 public:
  static const int TH_LOW = CMOS_TH_LOW;
  static const int TH_HIGH = CMOS_TH_HIGH;
  static const int HISTO_LENGTH = CMOS_HISTO_LENGTH;

  // The reference constructs an ORBmatcher as a stack temporary at every call site, several times per frame
  // (Tracking.cc:617, 766, 840, ...).  The device engines (about 30 device allocations and a stream each) are therefore
  // NOT owned by the object: they are borrowed from a per-thread cache keyed by (max_keypoints, max_points, device) and
  // live until the thread exits or ORBmatcher::release() is called — constructing a matcher costs nothing after the first.
  ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int max_keypoints = 4096, int max_points = 8192, int device = 0)
      : max_keypoints_(max_keypoints), max_points_(max_points), device_(device), mfNNratio(nnratio),
        mbCheckOrientation(checkOri) {
    Engines& e = cache().get(max_keypoints, max_points, device);
    if (!e.match) {
      cmos_match_params p = {1, max_keypoints, max_points, device};
      cmos_throw_if(cmos_match_create(&p, &e.match), "ORBmatcher");
    }
    h_ = e.match; kf_ = e.kf; eng_ = &e;
  }
  ~ORBmatcher() {}
  ORBmatcher(const ORBmatcher&) = delete;
  ORBmatcher& operator=(const ORBmatcher&) = delete;
  // destroys the calling thread's cached engines (no ORBmatcher of this thread may be alive)
  static void release() { cache().clear(); }

  // Hamming distance of two 256-bit descriptors (ORBmatcher.cc:1422-1437)
  static int DescriptorDistance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
  }

  // Track the local map: returns the number of matches, writes F.map_points (ORBmatcher.cc:42-119)
  int SearchByProjection(FrameView& F, const MapPointsView& pts, const float th = 3) {
    bind(F);
    std::vector<int32_t> assign(F.N, -1);
    int32_t n = pts.n, nm = 0;
    cmos_throw_if(cmos_match_search_by_projection_points(h_, &n, pts.in_view, pts.level, pts.view_cos, pts.proj_xy,
                                                         pts.descriptors, pts.has_obs, pts.n, th, mfNNratio,
                                                         F.claimed.data(), assign.data(), &nm, 0, nullptr),
                  "ORBmatcher::SearchByProjection(F, points)");
    for (int i = 0; i < F.N; i++)
      if (assign[i] >= 0) F.map_points[i] = assign[i];
    return nm;
  }

  // Track from the previous frame (ORBmatcher.cc:1161-1271); F.map_points[i] = index of the last-frame keypoint
  int SearchByProjection(FrameView& CurrentFrame, const LastFrameView& LastFrame, const float th) {
    bind(CurrentFrame);
    std::vector<int32_t> match(CurrentFrame.N, -1);
    int32_t nl = LastFrame.N, nm = 0;
    cmos_throw_if(cmos_match_search_by_projection_frame(h_, CurrentFrame.Tcw, LastFrame.undistort_keypoints, &nl,
                                                        LastFrame.flags, LastFrame.world_pos, LastFrame.descriptors,
                                                        LastFrame.N, th, mbCheckOrientation ? 1 : 0,
                                                        CurrentFrame.claimed.data(), match.data(), &nm, 0, nullptr),
                  "ORBmatcher::SearchByProjection(cur, last)");
    for (int i = 0; i < CurrentFrame.N; i++)
      if (match[i] >= 0) CurrentFrame.map_points[i] = match[i];
    return nm;
  }

  // ---- relocalisation: project the keyframe's map points into the frame (ORBmatcher.cc:1273-1384) ----
  // already_found[i] = sAlreadyFound.count(pKF map point i); CurrentFrame.map_points[i2] = keyframe keypoint index
  int SearchByProjection(FrameView& CurrentFrame, const KeyFrameView& KF, const uint8_t* already_found, const float th,
                         const int ORBdist) {
    kf_view(0, CurrentFrame.camera, false, CurrentFrame.undistort_keypoints, CurrentFrame.descriptors, CurrentFrame.N);
    if ((int)CurrentFrame.map_points.size() != CurrentFrame.N) CurrentFrame.map_points.assign(CurrentFrame.N, -1);
    std::vector<uint8_t> valid(KF.N), has(CurrentFrame.N);
    std::vector<float> angle(KF.N);
    for (int i = 0; i < KF.N; i++) {
      valid[i] = KF.mp_valid[i] && !(already_found && already_found[i]);
      angle[i] = KF.undistort_keypoints[i].angle;
    }
    for (int i = 0; i < CurrentFrame.N; i++) has[i] = CurrentFrame.map_points[i] >= 0;
    std::vector<int32_t> match(CurrentFrame.N, -1);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_by_projection_reloc(kf(), CurrentFrame.Tcw, KF.N, valid.data(), KF.mp_world_pos,
                                                          KF.mp_min_distance, KF.mp_max_distance, KF.mp_descriptors,
                                                          angle.data(), th, ORBdist, mbCheckOrientation ? 1 : 0, has.data(),
                                                          match.data(), &nm),
                  "ORBmatcher::SearchByProjection(F, KF)");
    for (int i = 0; i < CurrentFrame.N; i++)
      if (match[i] >= 0) CurrentFrame.map_points[i] = match[i];
    return nm;
  }

  // ---- loop closing: project points with a Sim3 (ORBmatcher.cc:258-361); vpMatched[idx] = point index or -1 ----
  int SearchByProjection(const KeyFrameView& KF, const double Scw[16], const PointsView& pts,
                         std::vector<int32_t>& vpMatched, int th) {
    kf_view(0, KF);
    if ((int)vpMatched.size() != KF.N) vpMatched.assign(KF.N, -1);
    std::vector<uint8_t> matched(KF.N);
    for (int i = 0; i < KF.N; i++) matched[i] = vpMatched[i] >= 0;
    std::vector<int32_t> assign(KF.N, -1);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_by_projection_sim3(kf(), Scw, pts.n, pts.skip, pts.world_pos, pts.normal,
                                                         pts.min_distance, pts.max_distance, pts.descriptors, th,
                                                         matched.data(), assign.data(), &nm),
                  "ORBmatcher::SearchByProjection(KF, Scw)");
    for (int i = 0; i < KF.N; i++)
      if (assign[i] >= 0) vpMatched[i] = assign[i];
    return nm;
  }

  // ---- bag of words (ORBmatcher.cc:151-256): vpMapPointMatches[idxF] = keyframe keypoint whose point F gets ----
  int SearchByBoW(const KeyFrameView& KF, const FrameView& F, std::vector<int32_t>& vpMapPointMatches) {
    kf_view(0, KF);
    kf_view(1, F.camera, false, F.undistort_keypoints, F.descriptors, F.N);
    vpMapPointMatches.assign(F.N, -1);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_by_bow(kf(), 0, KF.mp_valid, &KF.feature_vector, nullptr, &F.feature_vector, mfNNratio,
                                             mbCheckOrientation ? 1 : 0, vpMapPointMatches.data(), &nm),
                  "ORBmatcher::SearchByBoW(KF, F)");
    return nm;
  }
  // (ORBmatcher.cc:470-580): vpMatches12[idx1] = idx2 (the keypoint of KF2 whose map point matches)
  int SearchByBoW(const KeyFrameView& KF1, const KeyFrameView& KF2, std::vector<int32_t>& vpMatches12) {
    kf_view(0, KF1);
    kf_view(1, KF2);
    vpMatches12.assign(KF1.N, -1);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_by_bow(kf(), 1, KF1.mp_valid, &KF1.feature_vector, KF2.mp_valid, &KF2.feature_vector,
                                             mfNNratio, mbCheckOrientation ? 1 : 0, vpMatches12.data(), &nm),
                  "ORBmatcher::SearchByBoW(KF1, KF2)");
    return nm;
  }

  // ---- map initialisation (ORBmatcher.cc:363-468); vbPrevMatched holds x,y per F1 keypoint ----
  int SearchForInitialization(const FrameView& F1, const FrameView& F2, std::vector<float>& vbPrevMatched,
                              std::vector<int>& vnMatches12, int windowSize = 10) {
    kf_view(0, F1.camera, false, F1.undistort_keypoints, F1.descriptors, F1.N);
    kf_view(1, F2.camera, false, F2.undistort_keypoints, F2.descriptors, F2.N);
    vnMatches12.assign(F1.N, -1);
    vbPrevMatched.resize((size_t)F1.N * 2);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_for_initialization(kf(), vbPrevMatched.data(), windowSize, mfNNratio,
                                                         mbCheckOrientation ? 1 : 0, vnMatches12.data(), &nm),
                  "ORBmatcher::SearchForInitialization");
    return nm;
  }

  // ---- triangulation candidates (ORBmatcher.cc:582-722); monocular: bOnlyStereo must be false ----
  int SearchForTriangulation(const KeyFrameView& KF1, const KeyFrameView& KF2, const double F12[9],
                             std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    vMatchedPairs.clear();
    if (bOnlyStereo) return 0;            // no keypoint has a right coordinate in the monocular system (Frame.cc:127)
    kf_view(0, KF1);
    kf_view(1, KF2);
    std::vector<int32_t> m12(KF1.N, -1);
    int32_t nm = 0;
    cmos_throw_if(cmos_kfmatch_search_for_triangulation(kf(), KF1.mp_present, &KF1.feature_vector, KF2.mp_present,
                                                        &KF2.feature_vector, F12, KF1.Ow, KF2.Rcw, KF2.tcw, KF2.level_sigma2,
                                                        mbCheckOrientation ? 1 : 0, m12.data(), &nm),
                  "ORBmatcher::SearchForTriangulation");
    vMatchedPairs.reserve(nm);
    for (int i = 0; i < KF1.N; i++)
      if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
    return nm;
  }

  // ---- Sim3-guided mutual search (ORBmatcher.cc:956-1159); vpMatches12[i1] = idx2 or -1, in/out ----
  // already2[idx2] must be set by the caller for the keyframe-2 keypoints of the points already in vpMatches12
  // (pMP->GetIndexInKeyFrame(pKF2), :988-990)
  int SearchBySim3(const KeyFrameView& KF1, const KeyFrameView& KF2, std::vector<int32_t>& vpMatches12,
                   const uint8_t* already2, const float& s12, const double R12[9], const double t12[3], const float th) {
    kf_view(0, KF1);
    kf_view(1, KF2);
    if ((int)vpMatches12.size() != KF1.N) vpMatches12.assign(KF1.N, -1);
    std::vector<uint8_t> a1(KF1.N), a2(KF2.N, 0);
    for (int i = 0; i < KF1.N; i++) a1[i] = vpMatches12[i] >= 0;
    if (already2) a2.assign(already2, already2 + KF2.N);
    double pose1[12], pose2[12];
    for (int i = 0; i < 9; i++) { pose1[i] = KF1.Rcw[i]; pose2[i] = KF2.Rcw[i]; }
    for (int i = 0; i < 3; i++) { pose1[9 + i] = KF1.tcw[i]; pose2[9 + i] = KF2.tcw[i]; }
    std::vector<int32_t> m12(KF1.N, -1);
    int32_t nf = 0;
    cmos_throw_if(cmos_kfmatch_search_by_sim3(kf(), pose1, pose2, s12, R12, t12, KF1.mp_valid, a1.data(), KF1.mp_world_pos,
                                              KF1.mp_min_distance, KF1.mp_max_distance, KF1.mp_descriptors, KF2.mp_valid,
                                              a2.data(), KF2.mp_world_pos, KF2.mp_min_distance, KF2.mp_max_distance,
                                              KF2.mp_descriptors, th, m12.data(), &nf),
                  "ORBmatcher::SearchBySim3");
    for (int i = 0; i < KF1.N; i++)
      if (m12[i] >= 0) vpMatches12[i] = m12[i];
    return nf;
  }

  // ---- Fuse (ORBmatcher.cc:724-842 and 844-954): decisions only, see FuseDecisions ----
  int Fuse(const KeyFrameView& KF, const PointsView& pts, const float th, FuseDecisions& out) {
    double pose[15];
    for (int i = 0; i < 9; i++) pose[i] = KF.Rcw[i];
    for (int i = 0; i < 3; i++) { pose[9 + i] = KF.tcw[i]; pose[12 + i] = KF.Ow[i]; }
    return fuse(KF, 0, pose, pts, th, out);
  }
  int Fuse(const KeyFrameView& KF, const double Scw[16], const PointsView& pts, float th, FuseDecisions& out) {
    return fuse(KF, 1, Scw, pts, th, out);
  }

  cmos_match_t handle() { return h_; }

 protected:
  void bind(FrameView& F) {
    if ((int)F.map_points.size() != F.N) F.map_points.assign(F.N, -1);
    if ((int)F.claimed.size() != F.N) F.claimed.assign(F.N, 0);
    int32_t n = F.N;
    cmos_throw_if(cmos_match_set_frames(h_, &F.camera, F.undistort_keypoints, F.descriptors, &n, 1, F.N, 0, nullptr),
                  "ORBmatcher: Frame grid");
  }
  int fuse(const KeyFrameView& KF, int sim3, const double* pose, const PointsView& pts, float th, FuseDecisions& out) {
    kf_view(0, KF);
    out.best_idx.assign(pts.n, -1);
    out.best_dist.assign(pts.n, 256);
    int32_t nf = 0;
    cmos_throw_if(cmos_kfmatch_fuse(kf(), sim3, pose, KF.inv_level_sigma2, pts.n, pts.skip, pts.world_pos, pts.normal,
                                    pts.min_distance, pts.max_distance, pts.descriptors, th, out.best_idx.data(),
                                    out.best_dist.data(), &nf),
                  "ORBmatcher::Fuse");
    return nf;
  }
  cmos_kfmatch_t kf() {
    if (!kf_) {
      cmos_kfmatch_params p = {max_keypoints_, max_points_, 16384, device_};
      cmos_throw_if(cmos_kfmatch_create(&p, &kf_), "ORBmatcher (keyframe searches)");
      eng_->kf = kf_;
    }
    return kf_;
  }
  struct Engines { cmos_match_t match = nullptr; cmos_kfmatch_t kf = nullptr; };
  struct Cache {                                   // thread_local: destroyed at thread exit
    std::map<std::tuple<int, int, int>, Engines> m;
    Engines& get(int kp, int pts, int dev) { return m[std::make_tuple(kp, pts, dev)]; }
    void clear() {
      for (auto& kv : m) { cmos_match_destroy(kv.second.match); cmos_kfmatch_destroy(kv.second.kf); }
      m.clear();
    }
    ~Cache() { clear(); }
  };
  static Cache& cache() {
    static thread_local Cache c;
    return c;
  }
  Engines* eng_ = nullptr;
  void kf_view(int slot, const cmos_camera& cam, bool is_keyframe, const KeyPoint* kps, const uint8_t* desc, int n) {
    cmos_throw_if(cmos_kfmatch_set_view(kf(), slot, &cam, is_keyframe ? 1 : 0, kps, desc, n), "ORBmatcher: view grid");
  }
  void kf_view(int slot, const KeyFrameView& KF) {
    kf_view(slot, KF.camera, KF.is_keyframe, KF.undistort_keypoints, KF.descriptors, KF.N);
  }
  cmos_match_t h_ = nullptr;
  cmos_kfmatch_t kf_ = nullptr;
  int max_keypoints_, max_points_, device_;
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
#endif
