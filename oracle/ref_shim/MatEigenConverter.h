// TEST INFRASTRUCTURE — shadows /root/reference/include/MatEigenConverter.h in the oracle/_ref build.  The reference's header
// drags in LoopClosing.h (Tracking, LocalMapping, Sophus, Ceres ...), none of which the compiled files need: Frame.cc and
// KeyFrame.cc only call toDescriptorVector (src/MatEigenConverter.cc:87-95: one cv::Mat row view per descriptor), which
// oracle/ref_shim/ref_matcher.cpp defines with the same body.
#ifndef MAT_EIGEN_CONVERTER_H_
#define MAT_EIGEN_CONVERTER_H_
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include <vector>
class MatEigenConverter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};
#endif
