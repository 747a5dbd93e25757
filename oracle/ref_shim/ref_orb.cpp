// TEST INFRASTRUCTURE — C harness around the UNMODIFIED reference class ORB_SLAM2::ORBextractor
// (/root/reference/include/ORBextractor.h, src/ORBextractor.cc), compiled against the OpenCV stand-in in this
// directory.  Same entry-point shapes as oracle/orb_oracle.cpp so tests can run the two side by side.
#include <cstring>
#include <vector>

#include "ORBextractor.h"

namespace refarena { void begin(); void end(); }

namespace {
struct Handle {
  ORB_SLAM2::ORBextractor ex;
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  Handle(int nf, float sc, int nl, int ini, int mn) : ex(nf, sc, nl, ini, mn) {}
};
}  // namespace

extern "C" {

void* ref_orb_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
  return new Handle(nfeatures, scale, nlevels, ini_th, min_th);
}
void ref_orb_destroy(void* h) { delete (Handle*)h; }

void ref_orb_tables(void* h, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2) {
  Handle* e = (Handle*)h;
  std::vector<float> a = e->ex.GetScaleFactors(), b = e->ex.GetInverseScaleFactors(),
                     c = e->ex.GetScaleSigmaSquares(), d = e->ex.GetInverseScaleSigmaSquares();
  for (size_t i = 0; i < a.size(); i++) { sf[i] = a[i]; inv_sf[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

// ORBextractor::operator() (ORBextractor.cc:1043) on a caller-owned CV_8UC1 image.  Returns the keypoint count.
int ref_orb_extract(void* h, const uint8_t* img, int w, int ht, int pitch) {
  Handle* e = (Handle*)h;
  e->kps = std::vector<cv::KeyPoint>();   // drop arena-backed storage of the previous call before the arena rewinds
  e->desc.release();
  cv::Mat image(ht, w, CV_8UC1, (void*)img, (size_t)pitch);
  refarena::begin();
  {
    std::vector<cv::KeyPoint> kps;
    e->ex(image, cv::Mat(), kps, e->desc);
    refarena::end();
    e->kps.assign(kps.begin(), kps.end());   // copy to ordinary heap storage; kps' arena block is a no-op delete
  }
  return (int)e->kps.size();
}

int ref_orb_result(void* h, void* kps28, uint8_t* desc32) {
  Handle* e = (Handle*)h;
  int n = (int)e->kps.size();
  if (kps28 && n) std::memcpy(kps28, e->kps.data(), (size_t)n * sizeof(cv::KeyPoint));
  if (desc32)
    for (int i = 0; i < n; i++) std::memcpy(desc32 + 32 * (size_t)i, e->desc.ptr(i), 32);
  return n;
}

void ref_orb_level_size(void* h, int level, int* w, int* ht) {
  const cv::Mat& m = ((Handle*)h)->ex.mvImagePyramid[level];
  *w = m.cols; *ht = m.rows;
}
// Pyramid level INCLUDING its 19-pixel border (out: (rows+38) x (cols+38), tight): the border is what IC_Angle and
// the cell FAST read, so it is part of what the reference computes (ORBextractor.cc:1115-1128).
void ref_orb_level_image_bordered(void* h, int level, uint8_t* out) {
  const cv::Mat& m = ((Handle*)h)->ex.mvImagePyramid[level];
  const int B = 19, W = m.cols + 2 * B;
  for (int y = -B; y < m.rows + B; y++) std::memcpy(out + (size_t)(y + B) * W, m.data + (long)y * (long)m.step - B, W);
}

}  // extern "C"
