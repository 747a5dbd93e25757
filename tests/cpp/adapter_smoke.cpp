// Compiles the C++ drop-in adapters against libcmos_b200.so and exercises them.  Without a GPU it checks that
// construction fails loudly (no CPU fallback); with a GPU it extracts, matches and optimises a tiny scene.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/orb_slam2/CeresOptimizer.h"
#include "../../include/orb_slam2/ORBextractor.h"
#include "../../include/orb_slam2/ORBmatcher.h"
#include "../../include/orb_slam2/TrackingFrontEnd.h"

using namespace ORB_SLAM2;

int main() {
  const int W = 640, H = 480;
  std::vector<uint8_t> img((size_t)W * H);
  unsigned s = 12345;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      s = s * 1664525u + 1013904223u;
      int v = (((x / 23) + (y / 17)) & 1) ? 60 : 190;   // checkerboard: corners everywhere
      img[(size_t)y * W + x] = (uint8_t)(v + (int)((s >> 24) % 9) - 4);
    }
  try {
    ORBextractor ext(1000, 1.2f, 8, 20, 7, W, H);
    ImageView iv; iv.data = img.data(); iv.rows = H; iv.cols = W; iv.step = W;
    std::vector<KeyPoint> kps; DescriptorMat desc;
    ext(iv, ImageView(), kps, desc);
    std::printf("extract: %zu keypoints, %d descriptor rows\n", kps.size(), desc.rows);
    if (kps.empty() || desc.rows != (int)kps.size()) return 2;
    // match the frame against itself: every valid keypoint must find itself
    FrameView F;
    std::vector<float> sf = ext.GetScaleFactors();
    cmos_camera_init(&F.camera, W, H, 520.9f, 521.0f, 325.1f, 249.7f, sf.data(), (int)sf.size(), 1.2f);
    F.undistort_keypoints = kps.data(); F.descriptors = desc.data.data(); F.N = (int)kps.size();
    for (int i = 0; i < 16; i++) F.Tcw[i] = (i % 5 == 0) ? 1.0 : 0.0;
    std::vector<uint8_t> flags(kps.size(), 3);
    std::vector<double> X(kps.size() * 3);
    for (size_t i = 0; i < kps.size(); i++) {
      const double z = 10.0;
      X[3 * i] = (kps[i].x - 325.1f) / 520.9f * z; X[3 * i + 1] = (kps[i].y - 249.7f) / 521.0f * z; X[3 * i + 2] = z;
    }
    LastFrameView L; L.undistort_keypoints = kps.data(); L.N = (int)kps.size(); L.flags = flags.data();
    L.world_pos = X.data(); L.descriptors = desc.data.data();
    ORBmatcher matcher(0.9f, true);
    const int nm = matcher.SearchByProjection(F, L, 15.f);
    std::printf("SearchByProjection(cur,last): %d matches of %d\n", nm, F.N);
    if (nm < F.N / 2) return 3;
    // the keyframe searches: BoW of the frame against itself (every feature in node = octave), Fuse decisions
    {
      KeyFrameView KF;
      KF.camera = F.camera; KF.undistort_keypoints = kps.data(); KF.descriptors = desc.data.data(); KF.N = (int)kps.size();
      std::vector<uint8_t> valid(kps.size(), 1);
      KF.mp_valid = valid.data(); KF.mp_present = valid.data(); KF.mp_world_pos = X.data(); KF.mp_descriptors = desc.data.data();
      std::vector<float> mind(kps.size()), maxd(kps.size()), sig(sf.size()), isig(sf.size());
      for (size_t i = 0; i < kps.size(); i++) { maxd[i] = 10.0f * sf[kps[i].octave]; mind[i] = maxd[i] / sf.back(); }
      for (size_t l = 0; l < sf.size(); l++) { sig[l] = sf[l] * sf[l]; isig[l] = 1.0f / sig[l]; }
      KF.mp_min_distance = mind.data(); KF.mp_max_distance = maxd.data(); KF.level_sigma2 = sig.data(); KF.inv_level_sigma2 = isig.data();
      for (int i = 0; i < 9; i++) KF.Rcw[i] = (i % 4 == 0) ? 1.0 : 0.0;
      for (int i = 0; i < 3; i++) { KF.tcw[i] = 0.0; KF.Ow[i] = 0.0; }
      std::vector<int32_t> nodes, start(1, 0), feats;
      for (int l = 0; l < (int)sf.size(); l++) {
        for (size_t i = 0; i < kps.size(); i++) if (kps[i].octave == l) feats.push_back((int32_t)i);
        if ((int)feats.size() > start.back()) { nodes.push_back(l); start.push_back((int32_t)feats.size()); }
      }
      KF.feature_vector.n_nodes = (int32_t)nodes.size(); KF.feature_vector.node_ids = nodes.data();
      KF.feature_vector.start = start.data(); KF.feature_vector.features = feats.data();
      std::vector<int32_t> m12;
      const int nb = matcher.SearchByBoW(KF, KF, m12);
      int self = 0;
      for (size_t i = 0; i < m12.size(); i++) self += m12[i] == (int)i;
      std::printf("SearchByBoW(KF,KF): %d matches, %d onto themselves\n", nb, self);
      if (nb < F.N / 2 || self != nb) return 5;
      std::vector<double> normal(kps.size() * 3);
      for (size_t i = 0; i < kps.size(); i++) {
        const double n = std::sqrt(X[3 * i] * X[3 * i] + X[3 * i + 1] * X[3 * i + 1] + X[3 * i + 2] * X[3 * i + 2]);
        for (int k = 0; k < 3; k++) normal[3 * i + k] = X[3 * i + k] / n;
      }
      std::vector<uint8_t> skip(kps.size(), 0);
      PointsView pts; pts.n = (int)kps.size(); pts.skip = skip.data(); pts.world_pos = X.data(); pts.normal = normal.data();
      pts.min_distance = mind.data(); pts.max_distance = maxd.data(); pts.descriptors = desc.data.data();
      FuseDecisions fd;
      const int nf = matcher.Fuse(KF, pts, 3.0f, fd);
      int own = 0;
      for (size_t i = 0; i < fd.best_idx.size(); i++) own += fd.best_idx[i] == (int)i;
      std::printf("Fuse: %d decisions, %d onto the point's own keypoint\n", nf, own);
      if (nf < F.N / 2 || own < nf * 9 / 10) return 6;
    }
    // pose optimisation on exact correspondences perturbed by a small translation
    FramePoseView P;
    const double pose0[7] = {0.05, -0.03, 0.02, 0, 0, 0, 1};
    for (int i = 0; i < 7; i++) P.pose7[i] = pose0[i];
    std::vector<float> uv(kps.size() * 2), w(kps.size(), 1.0f);
    for (size_t i = 0; i < kps.size(); i++) { uv[2 * i] = kps[i].x; uv[2 * i + 1] = kps[i].y; }
    P.n = (int)kps.size(); P.world_pos = X.data(); P.uv = uv.data(); P.inv_sigma2 = w.data();
    P.K4[0] = 520.9f; P.K4[1] = 521.0f; P.K4[2] = 325.1f; P.K4[3] = 249.7f;
    const int inl = CeresOptimizer::PoseOptimization(&P);
    std::printf("PoseOptimization: %d inliers, t = %.2e %.2e %.2e\n", inl, P.pose7[0], P.pose7[1], P.pose7[2]);
    if (inl < P.n * 9 / 10 || std::fabs(P.pose7[0]) > 1e-3) return 4;
    {
      // essential graph: 12 keyframes on a line (identity rotations), the spanning tree measured from poses that are
      // stretched by 2 % per keyframe, the last keyframe corrected back onto the true position: the optimisation has to
      // spread the correction, keep the loop keyframe where it is and return rigid poses.
      const int n = 12;
      EssentialGraphView g;
      g.n_keyframes = n;
      g.id.resize(n); g.pose.resize(n); g.parent.assign(n, -1); g.children.resize(n); g.loop_edges.resize(n);
      g.covisibles.resize(n); g.loop_connections.resize(n); g.loop_connection_weight.resize(n);
      g.has_corrected.assign(n, 0); g.has_non_corrected.assign(n, 0); g.corrected.resize(n); g.non_corrected.resize(n);
      for (int k = 0; k < n; k++) {
        g.id[k] = (unsigned long)k;
        g.pose[k].t[0] = -1.02 * k;                 // Tcw translation = -centre
        if (k > 0) { g.parent[k] = k - 1; g.children[k - 1].push_back(k); }
        if (k > 1) g.covisibles[k].push_back(k - 2);
      }
      g.loop_keyframe = 0; g.current_keyframe = n - 1;
      g.has_corrected[n - 1] = g.has_non_corrected[n - 1] = 1;
      g.non_corrected[n - 1] = g.pose[n - 1];
      g.corrected[n - 1] = g.pose[n - 1]; g.corrected[n - 1].t[0] = -1.0 * (n - 1);
      g.loop_connections[n - 1].push_back(0); g.loop_connection_weight[n - 1].push_back(5);
      std::vector<double> px(3 * 4, 0.0); std::vector<int32_t> pref(4);
      for (int p = 0; p < 4; p++) { pref[p] = 3 * p + 1; px[3 * p] = 1.02 * pref[p]; px[3 * p + 2] = 5.0; }
      g.n_points = 4; g.point_pos = px.data(); g.point_ref = pref.data();
      CeresOptimizer::OptimizeEssentialGraph(g);
      const double c0 = -g.Tiw[3], c6 = -g.Tiw[16 * 6 + 3], c11 = -g.Tiw[16 * 11 + 3];
      std::printf("OptimizeEssentialGraph: centres %.4f %.4f %.4f, point 1 x %.4f\n", c0, c6, c11, g.corrected_pos[3]);
      // the loop keyframe stays; the 0.22 m loop error is spread over the twelve edges (the CPU oracle gives 6.0576 and
      // 11.1730 for keyframes 6 and 11, from 6.12 and 11.22); a point rides with its reference keyframe
      if (std::fabs(c0) > 1e-9 || std::fabs(c11 - 11.1730) > 1e-3 || std::fabs(c6 - 6.0576) > 1e-3) return 7;
      if (std::fabs(g.corrected_pos[3] - (-g.Tiw[16 * 4 + 3])) > 0.05 || std::fabs(g.corrected_pos[5] - 5.0) > 0.5) return 8;
    }
    {
      // OptimizeSim3 on exact correspondences of two identical cameras one metre apart: as in the reference the solver
      // accepts no step from a consistent start (quirk Q6), every correspondence is an inlier and S12 comes back unchanged.
      const int n = 60;
      std::vector<float> o1(2 * n), o2(2 * n), w(n, 1.0f);
      std::vector<double> p2c(3 * n), p1c(3 * n);
      Sim3MatchesView m;
      const float K[4] = {520.9f, 521.0f, 325.1f, 249.7f};
      for (int k = 0; k < 4; k++) { m.K1[k] = K[k]; m.K2[k] = K[k]; }
      unsigned r = 777;
      for (int i = 0; i < n; i++) {
        r = r * 1664525u + 1013904223u; const double x = ((r >> 8) % 2000) / 1000.0 - 1.0;
        r = r * 1664525u + 1013904223u; const double y = ((r >> 8) % 1000) / 1000.0 - 0.5;
        r = r * 1664525u + 1013904223u; const double z = 4.0 + ((r >> 8) % 6000) / 1000.0;
        p1c[3 * i] = x; p1c[3 * i + 1] = y; p1c[3 * i + 2] = z;              // camera-1 coordinates
        p2c[3 * i] = x - 1.0; p2c[3 * i + 1] = y; p2c[3 * i + 2] = z;        // camera 2 sits at x = +1: P1 = P2 + (1,0,0)
        o1[2 * i] = (float)(K[0] * x / z + K[2]); o1[2 * i + 1] = (float)(K[1] * y / z + K[3]);
        o2[2 * i] = (float)(K[0] * (x - 1.0) / z + K[2]); o2[2 * i + 1] = (float)(K[1] * y / z + K[3]);
      }
      m.n = n; m.obs1 = o1.data(); m.inv_sigma1 = w.data(); m.P3D2c = p2c.data(); m.obs2 = o2.data(); m.inv_sigma2 = w.data();
      m.P3D1c = p1c.data();
      Sim3POD S12; S12.t[0] = 1.0;
      const int inl = CeresOptimizer::OptimizeSim3(m, S12, 10.0f, false);
      int bad = 0;
      for (size_t i = 0; i < m.is_bad.size(); i++) bad += m.is_bad[i];
      std::printf("OptimizeSim3: %d inliers, %d outliers, s = %.6f, t = %.4f %.4f %.4f\n", inl, bad, S12.s, S12.t[0], S12.t[1], S12.t[2]);
      // (the scale column of the reference's Jacobian is analytically zero, quirk Q7: if a step is accepted at this ~1e-9
      // cost the scale moves by rounding noise, so the check is a band, not an equality)
      if (inl < n - 2 || bad > 2 || std::fabs(S12.s - 1.0) > 0.05 || std::fabs(S12.t[0] - 1.0) > 0.05) return 9;
    }
    {
      // The batched front end: two copies of the frame as one batch, last frame = the frame itself.  The three forms of the
      // last-frame inputs (arrays, packed records, association records into the device-resident map-point table) must agree
      // with each other and with the per-frame matcher call above.
      TrackingFrontEnd fe(F.camera, 1000, 1.2f, 8, 20, 7, W, H, /*lanes=*/2, /*chunk_frames=*/1);
      const int B = 2, cap = fe.capacity(), N = (int)kps.size();
      std::vector<uint8_t> imgs((size_t)B * W * H);
      std::vector<double> T((size_t)B * 16, 0.0);
      std::vector<KeyPoint> lk((size_t)B * cap);
      std::vector<int32_t> lc(B, N), start(1, 0);
      std::vector<uint8_t> lf((size_t)B * cap, 0), ld((size_t)B * cap * 32, 0);
      std::vector<double> lx((size_t)B * cap * 3, 0.0);
      std::vector<cmos_last_point> pts;
      std::vector<cmos_track_assoc> as;
      for (int f = 0; f < B; f++) {
        std::copy(img.begin(), img.end(), imgs.begin() + (size_t)f * W * H);
        for (int i = 0; i < 4; i++) T[(size_t)f * 16 + 5 * i] = 1.0;
        std::copy(kps.begin(), kps.end(), lk.begin() + (size_t)f * cap);
        std::copy(flags.begin(), flags.end(), lf.begin() + (size_t)f * cap);
        std::copy(X.begin(), X.end(), lx.begin() + (size_t)f * cap * 3);
        std::copy(desc.data.begin(), desc.data.begin() + (size_t)N * 32, ld.begin() + (size_t)f * cap * 32);
        for (int i = 0; i < N; i++) {
          cmos_last_point r;
          std::copy(desc.data.begin() + (size_t)i * 32, desc.data.begin() + (size_t)(i + 1) * 32, r.descriptor);
          for (int k = 0; k < 3; k++) r.xw[k] = X[3 * (size_t)i + k];
          r.angle = kps[i].angle; r.index = (uint16_t)i; r.octave = (int8_t)kps[i].octave; r.flags = flags[i];
          pts.push_back(r);
          const cmos_track_assoc a = {(int32_t)i, kps[i].angle, (uint16_t)i, (int8_t)kps[i].octave, flags[i]};   // slot = keypoint index
          as.push_back(a);
        }
        start.push_back((int32_t)pts.size());
      }
      struct Out {
        std::vector<KeyPoint> k; std::vector<uint8_t> d; std::vector<int32_t> c, m, n;
        Out(int B, int cap) : k((size_t)B * cap), d((size_t)B * cap * 32), c(B), m((size_t)B * cap, -1), n(B) {}
      };
      Out o0(B, cap), o1(B, cap), o2(B, cap);
      auto batch = [&](Out& o) {
        TrackingFrontEnd::Batch b;
        b.images = imgs.data(); b.frame_stride = (int64_t)W * H; b.pitch = W; b.width = W; b.height = H; b.n_frames = B;
        b.Tcw = T.data(); b.keypoints = o.k.data(); b.descriptors = o.d.data(); b.counts = o.c.data(); b.capacity = cap;
        b.match = o.m.data(); b.nmatches = o.n.data();
        return b;
      };
      const int64_t t0 = fe.Submit(batch(o0), lk.data(), lc.data(), lf.data(), lx.data(), ld.data(), cap, 15.f);
      const int64_t t1 = fe.SubmitPoints(batch(o1), pts.data(), start.data(), 15.f);      // two batches in flight
      fe.Wait(t0); fe.Wait(t1);
      fe.ReserveMapPoints(N);
      fe.UpdateMapPointRange(0, N, X.data(), desc.data.data());
      fe.Wait(fe.SubmitMap(batch(o2), as.data(), start.data(), 15.f));
      std::printf("TrackingFrontEnd: %d + %d matches (arrays), %d + %d (records), %d + %d (map-point table)\n", o0.n[0], o0.n[1],
                  o1.n[0], o1.n[1], o2.n[0], o2.n[1]);
      for (int f = 0; f < B; f++)
        if (o0.c[f] != N || o0.n[f] != nm || o1.n[f] != nm || o2.n[f] != nm) return 11;
      if (o0.m != o1.m || o0.m != o2.m) return 12;
    }
    CeresOptimizer::release();
    std::printf("ADAPTERS_OK\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("ADAPTER_ERROR %s\n", e.what());
    return 10;
  }
}
