// TEST INFRASTRUCTURE — not product code.  Minimal stand-in for the OpenCV C++ API, used ONLY to compile the
// UNMODIFIED reference sources (/root/reference/src/ORBextractor.cc, /root/reference/lib/DBoW2/**) into
// oracle/_ref/libref.so (recipe: oracle/ref_shim/Makefile).  OpenCV's C++ headers are not in this image; cv2 is.
//
// What is real and what is stand-in:
//   * everything ORBextractor.cc / DBoW2 do themselves (cell loop, quadtree, IC angle, rBRIEF, orchestration,
//     vocabulary descent, weighting, scoring) is the reference's own compiled code;
//   * cv::Mat / KeyPoint / Point / Size / Rect / Input/OutputArray are re-implemented here with OpenCV's
//     semantics for the calls the reference makes (ROI views share the parent buffer, create() keeps a buffer of
//     matching size and type, `m = Mat::zeros(..)` assigns IN PLACE like a MatExpr, copyMakeBorder honours
//     BORDER_ISOLATED and the src-inside-dst aliasing of ORBextractor.cc:1115-1123);
//   * the five image primitives (FAST, resize, copyMakeBorder, GaussianBlur, fastAtan2) delegate to
//     oracle/orb_oracle.cpp's restatements, which tests/test_oracle_orb.py pins bit-exact to cv2 4.13.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <sstream>   // the real core.hpp pulls these in; DBoW2's TemplatedVocabulary.h relies on it
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

// oracle/orb_oracle.cpp (cv2-pinned restatements of the OpenCV primitives)
extern "C" {
void orb_oracle_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch);
void orb_oracle_blur7(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
int orb_oracle_fast(const uint8_t* img, int w, int h, int pitch, int threshold, int* out, int cap);
void orb_oracle_fast_atan2(const float* y, const float* x, float* out, int n);
// oracle/matcher_oracle.cpp: cv::undistortPoints restated, pinned bit-exact to cv2 4.13 (tests/test_oracle_matcher.py)
void frame_oracle_undistort_points(const float* K4, const float* dist, int n_dist, const float* xy_in, int n, float* xy_out);
}

typedef unsigned char uchar;

// cvRound: round half to even (SSE2 cvtsd2si / lrint in every OpenCV since 2.x); cvFloor / cvCeil as documented.
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

typedef ::uchar uchar;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U>
  Point_& operator*=(U s) { x = (T)(x * s); y = (T)(y * s); return *this; }   // saturate_cast<float> is the identity
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {   // 28 bytes, the layout of cv::KeyPoint
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

inline size_t shim_elem_size(int type) { return type == CV_32F ? 4 : 1; }

struct MatZeros { int rows, cols, type; };   // the one MatExpr the reference builds

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0)   // header over caller-owned memory
      : rows(r), cols(c), step(step_ ? step_ : c * shim_elem_size(type)), data((uchar*)ext), type_(type),
        whole_rows_(r), whole_cols_(c) {}
  Mat(const MatZeros& z) { *this = z; }
  Mat(const Mat& o) { share(o); }
  Mat& operator=(const Mat& o) {
    if (this != &o) { Mat keep(o); release(); share(keep); }
    return *this;
  }
  // Mat::operator=(const MatExpr&) evaluates INTO the existing matrix: create() is a no-op for a header of the
  // same size and type (also a row-range view, ORBextractor.cc:1037 relies on it), then the elements are set.
  Mat& operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step, 0, cols * shim_elem_size(type_));
    return *this;
  }
  ~Mat() { release(); }

  static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }

  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && type_ == type) return;
    release();
    rows = r; cols = c; type_ = type; channels_ = 1;
    step = c * shim_elem_size(type);
    whole_rows_ = r; whole_cols_ = c; ofs_x_ = ofs_y_ = 0;
    size_t bytes = step * (size_t)r;
    if (bytes) {
      buf_ = (uchar*)std::malloc(bytes);   // malloc, not operator new: pixel buffers stay out of the node arena
      ref_ = (long*)std::malloc(sizeof(long));
      *ref_ = 1;
      data = buf_;
    }
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() {
    if (ref_ && --*ref_ == 0) { std::free(buf_); std::free(ref_); }
    buf_ = nullptr; ref_ = nullptr; data = nullptr; rows = cols = 0; step = 0;
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t elemSize() const { return shim_elem_size(type_); }
  size_t step1() const { return step / shim_elem_size(type_); }
  bool isSubmatrix() const { return rows != whole_rows_ || cols != whole_cols_; }
  void locateROI(Size& whole, Point& ofs) const { whole = Size(whole_cols_, whole_rows_); ofs = Point(ofs_x_, ofs_y_); }

  Mat operator()(const Rect& r) const {
    assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
    Mat m(*this);
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    m.rows = r.height; m.cols = r.width;
    m.ofs_x_ = ofs_x_ + r.x; m.ofs_y_ = ofs_y_ + r.y;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }

  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, cols * elemSize());
  }

  // element i of a single-row or single-column matrix (Frame.cc:330 reads dist_coef_.at<float>(0))
  template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  // reshape(cn): only the channel count changes, N x 2 one-channel <-> N x 1 two-channel over the same bytes (Frame.cc:343-345)
  Mat reshape(int cn) const {
    Mat m(*this);
    const int total = cols * channels_;
    assert(total % cn == 0 && step == (size_t)cols * channels_ * elemSize());
    m.channels_ = cn; m.cols = total / cn; m.whole_cols_ = m.cols;
    return m;
  }
  int channels() const { return channels_; }
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    for (int y = 0; y < r; y++) for (int x = 0; x < c; x++) { if (type == CV_32F) m.at<float>(y, x) = 1.f; else m.at<uchar>(y, x) = 1; }
    return m;
  }
  void convertTo(Mat& dst, int type) const {      // CV_8U / CV_32F -> CV_32F, in place allowed (Frame.cc:485)
    Mat src = (dst.data == data) ? clone() : *this;
    if (type != CV_32F) throw std::runtime_error("cv shim: convertTo supports CV_32F targets only");
    Mat out(src.rows, src.cols, CV_32F);
    for (int y = 0; y < src.rows; y++)
      for (int x = 0; x < src.cols; x++) out.at<float>(y, x) = src.type() == CV_32F ? src.at<float>(y, x) : (float)src.at<uchar>(y, x);
    dst = out;
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

 private:
  void share(const Mat& o) {
    rows = o.rows; cols = o.cols; step = o.step; data = o.data; type_ = o.type_; channels_ = o.channels_;
    buf_ = o.buf_; ref_ = o.ref_;
    whole_rows_ = o.whole_rows_; whole_cols_ = o.whole_cols_; ofs_x_ = o.ofs_x_; ofs_y_ = o.ofs_y_;
    if (ref_) ++*ref_;
  }
  int type_ = CV_8U;
  int channels_ = 1;
  uchar* buf_ = nullptr;
  long* ref_ = nullptr;
  int whole_rows_ = 0, whole_cols_ = 0, ofs_x_ = 0, ofs_y_ = 0;
};

// InputArray / OutputArray: thin proxies over Mat, the only array kind the reference passes.
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }
 protected:
  Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) { m_ = &m; }
  void release() const { if (m_) m_->release(); }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void create(Size sz, int type) const { m_->create(sz, type); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline int borderInterpolate101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

// cv::copyMakeBorder for CV_8UC1 / BORDER_REFLECT_101 [+ BORDER_ISOLATED] (ORBextractor.cc:1122,1127), including
// the two behaviours the reference depends on: without BORDER_ISOLATED a sub-matrix source borrows real pixels of
// its parent; dst.create() keeps a buffer of matching size, so a source that is a view INTO dst stays valid.
inline void copyMakeBorder(InputArray src_, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  Mat src = src_.getMat();
  assert(src.type() == CV_8UC1);
  if (src.isSubmatrix() && (borderType & BORDER_ISOLATED) == 0) {
    Size whole; Point ofs;
    src.locateROI(whole, ofs);
    int dtop = std::min(ofs.y, top), dbottom = std::min(whole.height - src.rows - ofs.y, bottom);
    int dleft = std::min(ofs.x, left), dright = std::min(whole.width - src.cols - ofs.x, right);
    Mat grown(src.rows + dtop + dbottom, src.cols + dleft + dright, CV_8UC1, src.data - (size_t)dtop * src.step - dleft,
              src.step);
    src = grown;
    top -= dtop; left -= dleft; bottom -= dbottom; right -= dright;
  }
  dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
  if (top == 0 && left == 0 && bottom == 0 && right == 0) {
    if (src.data != dst.data || src.step != dst.step) src.copyTo(dst);
    return;
  }
  borderType &= ~BORDER_ISOLATED;
  if (borderType != BORDER_REFLECT_101) throw std::runtime_error("cv shim: only BORDER_REFLECT_101 is implemented");
  std::vector<int> tab(left + right);
  for (int i = 0; i < left; i++) tab[i] = borderInterpolate101(i - left, src.cols);
  for (int i = 0; i < right; i++) tab[left + i] = borderInterpolate101(src.cols + i, src.cols);
  for (int y = 0; y < src.rows; y++) {       // interior rows first (memmove: src may be the same bytes)
    uchar* d = dst.data + (size_t)(y + top) * dst.step;
    const uchar* s = src.data + (size_t)y * src.step;
    if (d + left != s) std::memmove(d + left, s, src.cols);
    for (int i = 0; i < left; i++) d[i] = s[tab[i]];
    for (int i = 0; i < right; i++) d[left + src.cols + i] = s[tab[left + i]];
  }
  const int dcols = dst.cols;
  for (int i = 0; i < top; i++) {
    int j = borderInterpolate101(i - top, src.rows);
    std::memcpy(dst.data + (size_t)i * dst.step, dst.data + (size_t)(j + top) * dst.step, dcols);
  }
  for (int i = 0; i < bottom; i++) {
    int j = borderInterpolate101(src.rows + i, src.rows);
    std::memcpy(dst.data + (size_t)(top + src.rows + i) * dst.step, dst.data + (size_t)(j + top) * dst.step, dcols);
  }
}

// cv::resize(..., INTER_LINEAR) for CV_8UC1 (ORBextractor.cc:1120).
inline void resize(InputArray src_, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
  Mat src = src_.getMat();
  if (interpolation != INTER_LINEAR || src.type() != CV_8UC1 || fx != 0 || fy != 0)
    throw std::runtime_error("cv shim: resize supports CV_8UC1 / INTER_LINEAR / explicit dsize only");
  dst.create(dsize, src.type());
  orb_oracle_resize(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (ORBextractor.cc:1086), in place allowed.
inline void GaussianBlur(InputArray src_, Mat& dst, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT) {
  Mat src = src_.getMat();
  if (ksize.width != 7 || ksize.height != 7 || sx != 2 || sy != 2 || borderType != BORDER_REFLECT_101 ||
      src.type() != CV_8UC1 || src.isSubmatrix())
    throw std::runtime_error("cv shim: GaussianBlur supports 7x7, sigma 2, BORDER_REFLECT_101, whole CV_8UC1 images only");
  Mat in = src.data == dst.data ? src.clone() : src;
  dst.create(src.rows, src.cols, src.type());
  orb_oracle_blur7(in.data, in.cols, in.rows, (int)in.step, dst.data, (int)dst.step);
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression=true): TYPE_9_16; keypoints in row-major scan order as
// KeyPoint(x, y, 7.f, -1, score) (ORBextractor.cc:809,814).
inline void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax = true) {
  Mat img = image_.getMat();
  if (!nonmax || img.type() != CV_8UC1) throw std::runtime_error("cv shim: FAST supports CV_8UC1 with non-max suppression only");
  keypoints.clear();
  if (img.empty()) return;
  int cap = std::max(1, (img.rows * img.cols + 3) / 4);
  std::vector<int> out((size_t)cap * 3);
  int n = orb_oracle_fast(img.data, img.cols, img.rows, (int)img.step, threshold, out.data(), cap);
  assert(n <= cap);
  keypoints.reserve(n);
  for (int i = 0; i < n; i++)
    keypoints.push_back(KeyPoint((float)out[3 * i], (float)out[3 * i + 1], 7.f, -1, (float)out[3 * i + 2]));
}

// The little float-matrix arithmetic Frame::ComputeStereoMatches needs (dead code in this monocular fork, but Frame.cc must
// compile unmodified): Mat - Mat, scalar * Mat, L1 norm of a difference.
enum { NORM_L1 = 2 };
inline Mat operator*(double s, const Mat& a) {
  Mat o(a.rows, a.cols, CV_32F);
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) o.at<float>(y, x) = (float)(s * a.at<float>(y, x));
  return o;
}
inline Mat operator-(const Mat& a, const Mat& b) {
  Mat o(a.rows, a.cols, CV_32F);
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) o.at<float>(y, x) = a.at<float>(y, x) - b.at<float>(y, x);
  return o;
}
inline double norm(const Mat& a, const Mat& b, int type) {
  if (type != NORM_L1) throw std::runtime_error("cv shim: only NORM_L1");
  double s = 0;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) s += std::fabs((double)a.at<float>(y, x) - b.at<float>(y, x));
  return s;
}

// cv::undistortPoints(src, dst, K, dist, R = empty, P = K) on an N x 1 two-channel CV_32F matrix (Frame.cc:344,371).
inline void undistortPoints(InputArray src_, Mat& dst, InputArray K_, InputArray dist_, InputArray R_, InputArray P_) {
  Mat src = src_.getMat(), K = K_.getMat(), dist = dist_.getMat(), P = P_.getMat();
  if (!R_.empty() || src.type() != CV_32F || src.channels() != 2 || K.type() != CV_32F || dist.type() != CV_32F)
    throw std::runtime_error("cv shim: undistortPoints supports N x 1 CV_32FC2 points, float K / dist, no rectification");
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    if (K.at<float>(i, j) != P.at<float>(i, j)) throw std::runtime_error("cv shim: undistortPoints needs P == K");
  const float K4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
  const int nd = dist.rows * dist.cols;
  std::vector<float> d(nd);
  for (int i = 0; i < nd; i++) d[i] = dist.at<float>(i);
  const int n = src.rows * src.cols;
  std::vector<float> in((size_t)2 * n), out((size_t)2 * n);
  std::memcpy(in.data(), src.data, in.size() * sizeof(float));
  frame_oracle_undistort_points(K4, d.data(), nd, in.data(), n, out.data());
  if (dst.data != src.data) { dst.create(src.rows, src.cols * 2, CV_32F); dst = dst.reshape(2); }
  std::memcpy(dst.data, out.data(), out.size() * sizeof(float));
}

inline float fastAtan2(float y, float x) {
  float a;
  orb_oracle_fast_atan2(&y, &x, &a, 1);
  return a;
}

// Only ComputeKeyPointsOld (dead code, ORBextractor.cc:855-1031) calls this; kept so the file compiles unmodified.
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n < 0 || (int)kps.size() <= n) return;
    if (n == 0) { kps.clear(); return; }
    std::nth_element(kps.begin(), kps.begin() + n, kps.end(),
                     [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    float ambiguous = kps[n - 1].response;
    auto end = std::partition(kps.begin() + n, kps.end(), [=](const KeyPoint& k) { return k.response >= ambiguous; });
    kps.resize(end - kps.begin());
  }
};

// cv::FileStorage: DBoW2's YAML save/load (TemplatedVocabulary.h:1453-1625) must compile; the reference itself
// loads its vocabulary with loadFromTextFile (plain ifstream), so these are never executed.
class FileNode {
 public:
  FileNode operator[](const std::string&) const { fail(); return FileNode(); }
  FileNode operator[](const char*) const { fail(); return FileNode(); }
  FileNode operator[](int) const { fail(); return FileNode(); }
  size_t size() const { fail(); return 0; }
  operator int() const { fail(); return 0; }
  operator double() const { fail(); return 0; }
  operator std::string() const { fail(); return std::string(); }
 private:
  static void fail() { throw std::runtime_error("cv shim: cv::FileStorage is not available"); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const char*, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode()[0]; }
  FileNode operator[](const char*) const { return FileNode()[0]; }
};
template <typename T>
inline FileStorage& operator<<(FileStorage& fs, const T&) {
  throw std::runtime_error("cv shim: cv::FileStorage is not available");
  return fs;
}

}  // namespace cv
