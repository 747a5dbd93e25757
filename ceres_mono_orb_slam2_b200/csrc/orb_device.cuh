// Device-side constants, geometry structs and bit-exact float helpers of the ORB front-end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmos {

constexpr int kBorder = 19;       // EDGE_THRESHOLD, ORBextractor.cc:74
constexpr int kMinBorder = 16;    // EDGE_THRESHOLD-3, ORBextractor.cc:773
constexpr int kXOff = 32;         // interior column 0 sits 32 bytes into a plane row (aligned rows)
constexpr int kMaxLevels = 16;    // == CMOS_MAX_LEVELS
constexpr int kFastThreads = 256;
constexpr int kTileW = 80, kTileH = 72;   // FAST cell tile in shared memory (cell <= 60+6, +3 misalignment)
constexpr int kCellListCap = 1024;
constexpr int kSmallTileW = 48, kSmallTileH = 48, kSmallListCap = 512;   // cells up to 44 x 48: (44-6)*(48-6)/4 = 399 survivors at most
constexpr int kOctThreads = 512;
constexpr int kBlurTW = 128, kBlurTH = 32, kBlurInPitch = kBlurTW + 16;   // input tile row: 4 + 128 + 4 used, 144 = a multiple of 16 (TMA box)
constexpr int kDescThreads = 256;

struct LevelGeom {
  int w, h;            // interior size, ComputePyramid (ORBextractor.cc:1112)
  int pitch, rows;     // bordered plane: rows = h + 38
  int plane_off;       // byte offset of the plane inside one frame's pyramid block
  float scale;         // mvScaleFactor[level]
  int patch;           // (int)(31 * scale)
  int quota;           // mnFeaturesPerLevel[level]
  int kp_off;          // first slot of this level in the per-frame staging array
  int n_cols, n_rows, w_cell, h_cell;   // cell grid, ORBextractor.cc:784-787
  int cand_off, cand_cap;               // slice of the per-frame candidate array
  int n_ini;           // quadtree roots, ORBextractor.cc:543
  float hx;            // root width
  int ow, oh;          // maxBorder - minBorder
};

struct OrbGeom {
  int nlevels, ini_th, min_th, kp_cap;
  long long frame_bytes;   // pyramid bytes per frame
  long long cand_frame;    // candidate slots per frame
  int umax[16];
  int vlim[16];            // vlim[|u|] = largest |v| with umax[|v|] >= |u|: column u of the centroid disc spans rows -vlim..vlim
  LevelGeom lv[kMaxLevels];
};

// ---- float helpers that must round like the CPU's unfused float32 / glibc ------------------------

// cv::fastAtan2: degree-7 odd polynomial, every op rounded separately.
__device__ __forceinline__ float dev_fast_atan2(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
  const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

__device__ constexpr float kFactorPi = (float)(3.14159265358979323846 / 180.f);   // ORBextractor.cc:107

// cosf/sinf exactly as glibc >= 2.28 computes them (sincosf by Szabolcs Nagy, ARM optimized-routines):
// double-precision range reduction by pi/2 and a degree-7/8 minimax polynomial, rounded once to float.
// The reference's computeOrbDescriptor calls std::cos/std::sin on floats (ORBextractor.cc:113); the CPU
// oracle calls glibc; tests/test_oracle_orb.py checks this restatement against libm on the host side and
// tests/test_orb_gpu.py checks the device against it.  Valid for |x| < 120 (angles are in [0, 2*pi)).
struct SinCosTab { double sign[4]; double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; };

__device__ __forceinline__ float sincosf_poly(double x, double x2, double c0, double c1, double c2, double c3,
                                              double c4, double s1, double s2, double s3, int n) {
  if ((n & 1) == 0) {
    double x3 = __dmul_rn(x, x2);
    double t1 = __dadd_rn(s2, __dmul_rn(x2, s3));
    double x7 = __dmul_rn(x3, x2);
    double s = __dadd_rn(x, __dmul_rn(x3, s1));
    return (float)__dadd_rn(s, __dmul_rn(x7, t1));
  } else {
    double x4 = __dmul_rn(x2, x2);
    double t2 = __dadd_rn(c3, __dmul_rn(x2, c4));
    double t1 = __dadd_rn(c0, __dmul_rn(x2, c1));
    double x6 = __dmul_rn(x4, x2);
    double c = __dadd_rn(t1, __dmul_rn(x4, c2));
    return (float)__dadd_rn(c, __dmul_rn(x6, t2));
  }
}

// which: 0 = sin, 1 = cos
__device__ __forceinline__ float glibc_sincosf(float y, int which) {
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
               c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  double x = (double)y;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
  if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff)) {
    double x2 = __dmul_rn(x, x);
    if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ff)) return which ? 1.0f : y;
    return sincosf_poly(x, x2, c0, c1, c2, c3, c4, s1, s2, s3, which);
  }
  double r = __dmul_rn(x, hpi_inv);
  int n = ((int)r + 0x800000) >> 24;
  x = __dsub_rn(x, __dmul_rn((double)n, hpi));
  const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
  const bool flip = (n & 2) != 0;                                      // second table = negated cosine coeffs
  const double f = flip ? -1.0 : 1.0;
  return sincosf_poly(__dmul_rn(x, sgn), __dmul_rn(x, x), f * c0, f * c1, f * c2, f * c3, f * c4, s1, s2, s3,
                      which ? (n ^ 1) : n);
}

__device__ __forceinline__ float glibc_cosf(float y) { return glibc_sincosf(y, 1); }
__device__ __forceinline__ float glibc_sinf(float y) { return glibc_sincosf(y, 0); }

}  // namespace cmos
