// Compiles the C++ drop-in adapters against libcmos_b200.so and exercises them.  Without a GPU it checks that
// construction fails loudly (no CPU fallback); with a GPU it extracts, matches and optimises a tiny scene.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/orb_slam2/CeresOptimizer.h"
#include "../../include/orb_slam2/ORBextractor.h"
#include "../../include/orb_slam2/ORBmatcher.h"

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
    CeresOptimizer::release();
    std::printf("ADAPTERS_OK\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("ADAPTER_ERROR %s\n", e.what());
    return 10;
  }
}
