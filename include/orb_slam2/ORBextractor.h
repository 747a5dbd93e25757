// Drop-in for the reference's ORB_SLAM2::ORBextractor (include/ORBextractor.h:46-111): same constructor,
// operator() and getters; the work happens in libcmos_b200.so (cmos_orb_*).  Header-only adapter.
#ifndef ORB_SLAM2_CMOS_ORBEXTRACTOR_H
#define ORB_SLAM2_CMOS_ORBEXTRACTOR_H

#include "views.h"

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  // max_width / max_height size the device buffers once (the reference allocates per call); device = CUDA ordinal
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int max_width = 1241,
               int max_height = 376, int device = 0)
      : nlevels_(nlevels), scaleFactor_(scaleFactor) {
    cmos_orb_params p = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height, 1, device};
    cmos_throw_if(cmos_orb_create(&p, &h_), "ORBextractor");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    cmos_orb_get_scale_factors(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                               mvInvLevelSigma2.data());
    cmos_orb_keypoint_capacity(h_, &capacity_);
    kp_buf_.resize(capacity_);
    desc_buf_.resize((size_t)capacity_ * 32);
  }
  ~ORBextractor() { cmos_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Compute the ORB features and descriptors on an image; mask is ignored like the reference (:58).
  void operator()(const ImageView& image, const ImageView& /*mask*/, std::vector<KeyPoint>& keypoints,
                  DescriptorMat& descriptors) {
    keypoints.clear();
    if (image.empty()) { descriptors.create(0); return; }            // ORBextractor.cc:1046-1047
    int32_t count = 0;
    cmos_throw_if(cmos_orb_extract(h_, image.data, 0, (int32_t)image.step, image.cols, image.rows, 1, kp_buf_.data(),
                                   desc_buf_.data(), &count, capacity_), "ORBextractor::operator()");
    keypoints.assign(kp_buf_.begin(), kp_buf_.begin() + count);
    descriptors.create(count);
    std::copy(desc_buf_.begin(), desc_buf_.begin() + (size_t)count * 32, descriptors.data.begin());
  }

  int GetLevels() { return nlevels_; }
  float GetScaleFactor() { return scaleFactor_; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  cmos_orb_t handle() { return h_; }   // batched / device-resident use: cmos_orb_extract_device

 protected:
  cmos_orb_t h_ = nullptr;
  int nlevels_;
  float scaleFactor_;
  int32_t capacity_ = 0;
  std::vector<KeyPoint> kp_buf_;
  std::vector<uint8_t> desc_buf_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM2
#endif
