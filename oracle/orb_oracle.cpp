// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (ceres_mono_orb_slam2_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker / timed CPU baseline.
//
// CPU restatement of the reference's ORB front-end, path (1) of BASELINE.json:north_star:
//   ORBextractor::ORBextractor      /root/reference/src/ORBextractor.cc:410-470
//   ORBextractor::operator()        /root/reference/src/ORBextractor.cc:1043-1105
//   ComputePyramid                  /root/reference/src/ORBextractor.cc:1107-1132
//   ComputeKeyPointsOctTree         /root/reference/src/ORBextractor.cc:765-853
//   DistributeOctTree / DivideNode  /root/reference/src/ORBextractor.cc:481-763
//   IC_Angle / computeOrientation   /root/reference/src/ORBextractor.cc:77-104,472-479
//   computeOrbDescriptor            /root/reference/src/ORBextractor.cc:108-147
// The OpenCV primitives the reference calls (cv::resize INTER_LINEAR, cv::copyMakeBorder REFLECT_101,
// cv::FAST TYPE_9_16 + NMS, cv::GaussianBlur 7x7 s=2, cv::fastAtan2, cvRound) are third-party and
// NOT vendored in /root/reference (CMakeLists.txt:21, unpinned).  They are restated here from OpenCV's
// published fixed-point algorithms and PINNED against cv2 4.13.0 (the only OpenCV in this image) by
// tests/test_oracle_orb.py, bit for bit.  The reference itself ships no tests or golden vectors.
// The reference's OWN code of this path is PINNED TO THE REFERENCE ITSELF: tests/test_ref_parity.py compiles the unmodified
// src/ORBextractor.cc against a minimal OpenCV stand-in (oracle/ref_shim/, whose five image primitives are the restatements
// below) and requires keypoints, descriptors and bordered pyramid levels identical to this file's, on the BASELINE
// configs[0] / configs[1] frames and on edge images.
//
// Float policy (SURVEY.md §7 hard part 2): compiled with -ffp-contract=off, no -march=native;
// cosf/sinf are glibc's; cvRound is round-half-to-even.
// Octree tie-break policy: the reference sorts pair<int size, ExtractorNode* ptr> (ORBextractor.cc:684) so
// equal-size nodes are ordered by heap address, which is not reproducible.  Here the node's creation
// sequence number stands in for the pointer (ascending seq == ascending "address").

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "../include/cmos_orb_pattern.h"

namespace {

constexpr int kPatchSize = 31;      // ORBextractor.cc:72
constexpr int kHalfPatch = 15;      // ORBextractor.cc:73
constexpr int kEdge = 19;           // ORBextractor.cc:74

inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound: round half to even
inline int cv_round_d(double v) { return (int)lrint(v); }

struct KeyPoint {      // == cv::KeyPoint memory layout (28 bytes)
  float x, y, size, angle, response;
  int32_t octave, class_id;
};

struct Image {         // owns a bordered buffer; (0,0) is the interior origin
  int w = 0, h = 0, pitch = 0;
  std::vector<uint8_t> buf;
  uint8_t* interior() { return buf.data() + kEdge * pitch + kEdge; }
  const uint8_t* interior() const { return buf.data() + kEdge * pitch + kEdge; }
};

// ---------------------------------------------------------------------------------------------
// cv::resize(..., INTER_LINEAR) for CV_8UC1 — OpenCV's fixed-point path (INTER_RESIZE_COEF_BITS=11).
// Called at ORBextractor.cc:1120.
void resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh,
                      int dpitch) {
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> alpha(2 * dw), beta(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    alpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
    alpha[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    beta[2 * dy] = (short)cv_round_f((1.f - fy) * 2048.f);
    beta[2 * dy + 1] = (short)cv_round_f(fy * 2048.f);
  }
  std::vector<int> row0(dw), row1(dw);
  auto hpass = [&](int sy, std::vector<int>& out) {
    sy = std::min(std::max(sy, 0), sh - 1);
    const uint8_t* s = src + (size_t)sy * spitch;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int s1 = (sx + 1 < sw) ? s[sx + 1] : s[sx];   // second tap has weight 0 when clamped
      out[dx] = s[sx] * alpha[2 * dx] + s1 * alpha[2 * dx + 1];
    }
  };
  for (int dy = 0; dy < dh; dy++) {
    hpass(yofs[dy], row0);
    hpass(yofs[dy] + 1, row1);
    int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
    uint8_t* d = dst + (size_t)dy * dpitch;
    for (int dx = 0; dx < dw; dx++) {
      int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
      d[dx] = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * (n - 1) - i;
  }
  return i;
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101) in place around the interior (ORBextractor.cc:1122,1127).
void fill_border101(Image& im) {
  uint8_t* in = im.interior();
  for (int y = -kEdge; y < im.h + kEdge; y++) {
    int sy = reflect101(y, im.h);
    for (int x = -kEdge; x < im.w + kEdge; x++) {
      if (y >= 0 && y < im.h && x >= 0 && x < im.w) continue;
      in[y * im.pitch + x] = in[sy * im.pitch + reflect101(x, im.w)];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, true) TYPE_9_16 on a sub-image (ORBextractor.cc:809,814).
// m(p) = max over the 16 arcs of 9 contiguous ring pixels of min(ring-c) and of min(c-ring);
// corner iff m > t, score = m-1; 3x3 strict non-max suppression with score 0 outside the 3-px margin.
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int fast_arc_max(const uint8_t* p, int pitch) {
  int c = p[0], d[25];
  for (int k = 0; k < 16; k++) d[k] = (int)p[kRingDy[k] * pitch + kRingDx[k]] - c;
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -255;
  for (int s = 0; s < 16; s++) {
    int lo = 255, hi = -255;
    for (int k = 0; k < 9; k++) { lo = std::min(lo, d[s + k]); hi = std::max(hi, d[s + k]); }
    best = std::max(best, std::max(lo, -hi));
  }
  return best;
}

struct Cand { int x, y, score; };

void fast_detect(const uint8_t* img, int w, int h, int pitch, int threshold, std::vector<Cand>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  std::vector<int> score((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      // Early reject, exact: an arc of 9 of the 16 ring pixels contains at least one pixel of every opposite pair
      // (k, k + 8), so a pair with both |ring - c| <= t rules the corner out.  OpenCV's FAST does the same kind of test
      // before scoring; without it this CPU path (also the bench's CPU baseline) would score every pixel.
      const uint8_t* p = img + y * pitch + x;
      const int c = p[0];
      bool reject = false;
      for (int k = 0; k < 8 && !reject; k += 2) {          // pairs (0,8), (2,10), (4,12), (6,14) — the rest is left to the score
        const int d0 = std::abs((int)p[kRingDy[k] * pitch + kRingDx[k]] - c);
        const int d1 = std::abs((int)p[kRingDy[k + 8] * pitch + kRingDx[k + 8]] - c);
        reject = d0 <= threshold && d1 <= threshold;
      }
      if (reject) continue;
      int m = fast_arc_max(p, pitch);
      if (m > threshold) score[y * w + x] = m - 1;
    }
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int s = score[y * w + x];
      if (!s) continue;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if ((dx || dy) && score[(y + dy) * w + x + dx] >= s) { keep = false; break; }
      if (keep) out.push_back({x, y, s});
    }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (ORBextractor.cc:1086):
// OpenCV 4.x bit-exact fixed-point path, Q8 kernel [18 34 48 56 48 34 18], final (sum + 2^15) >> 16.
void gaussian_blur7(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> tmp((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int acc = 0;
      for (int i = -3; i <= 3; i++) acc += k[i + 3] * src[y * spitch + reflect101(x + i, w)];
      tmp[(size_t)y * w + x] = (uint16_t)acc;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int i = -3; i <= 3; i++) acc += (uint32_t)k[i + 3] * tmp[(size_t)reflect101(y + i, h) * w + x];
      dst[y * dpitch + x] = (uint8_t)((acc + 32768u) >> 16);
    }
}

// cv::fastAtan2 (degrees), float32, unfused (ORBextractor.cc:103).
float fast_atan2(float y, float x) {
  const float s = (float)(180.0 / M_PI);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
  const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
  const float eps = (float)2.2204460492503131e-16;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------------------------------------
struct Node {        // ExtractorNode (ORBextractor.h:32-44); rectangle as UL.x, UR.x, UL.y, BR.y
  int x0, x1, y0, y1;
  std::vector<int> keys;     // indices into the candidate array, insertion order preserved
  bool no_more = false;
  long seq = 0;              // creation order, stands in for the heap address
  std::list<Node>::iterator self;
};

struct Extractor {
  int nfeatures, nlevels, ini_th, min_th;
  double scale_factor;   // ORBextractor.h:95 — member is double, ctor arg is float
  std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
  std::vector<int> quota, umax;
  std::vector<Image> pyr, blurred;
  std::vector<std::vector<KeyPoint>> cand, dist;   // per level: FAST candidates, distributed kps
  std::vector<KeyPoint> out_kps;
  std::vector<uint8_t> out_desc;
  long seq_counter = 0;

  Extractor(int nf, float scale, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), ini_th(ini), min_th(mn), scale_factor(scale) {
    sf.resize(nl); sigma2.resize(nl); inv_sf.resize(nl); inv_sigma2.resize(nl);
    sf[0] = 1.f; sigma2[0] = 1.f;
    for (int i = 1; i < nl; i++) {
      sf[i] = (float)(sf[i - 1] * scale_factor);
      sigma2[i] = sf[i] * sf[i];
    }
    for (int i = 0; i < nl; i++) { inv_sf[i] = 1.0f / sf[i]; inv_sigma2[i] = 1.0f / sigma2[i]; }
    quota.resize(nl);
    float factor = (float)(1.0f / scale_factor);
    float per = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      quota[l] = cv_round_f(per);
      sum += quota[l];
      per *= factor;
    }
    quota[nl - 1] = std::max(nf - sum, 0);
    umax.assign(kHalfPatch + 1, 0);
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    pyr.resize(nl); blurred.resize(nl); cand.resize(nl); dist.resize(nl);
  }

  void compute_pyramid(const uint8_t* img, int w, int h, int pitch) {
    for (int l = 0; l < nlevels; l++) {
      float s = inv_sf[l];
      Image& im = pyr[l];
      im.w = cv_round_f((float)w * s);
      im.h = cv_round_f((float)h * s);
      im.pitch = im.w + 2 * kEdge;
      im.buf.assign((size_t)im.pitch * (im.h + 2 * kEdge), 0);
      if (l == 0) {
        for (int y = 0; y < h; y++) std::memcpy(im.interior() + y * im.pitch, img + (size_t)y * pitch, w);
      } else {
        resize_linear_u8(pyr[l - 1].interior(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].pitch, im.interior(),
                         im.w, im.h, im.pitch);
      }
      fill_border101(im);
    }
  }

  void divide(const Node& p, Node out[4], const std::vector<KeyPoint>& kps) {
    const int half_x = (int)std::ceil(static_cast<float>(p.x1 - p.x0) / 2);
    const int half_y = (int)std::ceil(static_cast<float>(p.y1 - p.y0) / 2);
    const int mx = p.x0 + half_x, my = p.y0 + half_y;
    out[0] = Node{p.x0, mx, p.y0, my};
    out[1] = Node{mx, p.x1, p.y0, my};
    out[2] = Node{p.x0, mx, my, p.y1};
    out[3] = Node{mx, p.x1, my, p.y1};
    for (int idx : p.keys) {
      const KeyPoint& kp = kps[idx];
      int q = (kp.x < mx ? 0 : 1) + (kp.y < my ? 0 : 2);
      out[q].keys.push_back(idx);
    }
    for (int q = 0; q < 4; q++) out[q].no_more = out[q].keys.size() == 1;
  }

  std::vector<KeyPoint> distribute(const std::vector<KeyPoint>& kps, int min_x, int max_x, int min_y,
                                   int max_y, int N) {
    std::vector<KeyPoint> result;
    const int n_ini = (int)std::round(static_cast<float>(max_x - min_x) / (max_y - min_y));
    if (n_ini < 1) return result;   // reference divides by zero here; guarded (documented deviation)
    const float hx = static_cast<float>(max_x - min_x) / n_ini;
    std::list<Node> nodes;
    std::vector<Node*> roots(n_ini);
    for (int i = 0; i < n_ini; i++) {
      Node n{(int)(hx * static_cast<float>(i)), (int)(hx * static_cast<float>(i + 1)), 0, max_y - min_y};
      n.seq = seq_counter++;
      nodes.push_back(n);
      roots[i] = &nodes.back();
    }
    for (size_t i = 0; i < kps.size(); i++) roots[(size_t)(kps[i].x / hx)]->keys.push_back((int)i);
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->keys.size() == 1) { it->no_more = true; ++it; }
      else if (it->keys.empty()) it = nodes.erase(it);
      else ++it;
    }
    typedef std::pair<std::pair<int, long>, Node*> SizedNode;   // ((size, seq), node)
    std::vector<SizedNode> expandable;
    auto push_children = [&](Node ch[4], int* n_expand) {
      for (int q = 0; q < 4; q++) {
        if (ch[q].keys.empty()) continue;
        ch[q].seq = seq_counter++;
        nodes.push_front(ch[q]);
        nodes.front().self = nodes.begin();
        if (ch[q].keys.size() > 1) {
          if (n_expand) ++*n_expand;
          expandable.push_back({{(int)ch[q].keys.size(), ch[q].seq}, &nodes.front()});
        }
      }
    };
    bool finish = false;
    while (!finish) {
      int prev = (int)nodes.size(), n_expand = 0;
      expandable.clear();
      for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->no_more) { ++it; continue; }
        Node ch[4];
        divide(*it, ch, kps);
        push_children(ch, &n_expand);
        it = nodes.erase(it);
      }
      if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
        finish = true;
      } else if ((int)nodes.size() + n_expand * 3 > N) {
        while (!finish) {
          prev = (int)nodes.size();
          std::vector<SizedNode> todo = expandable;
          expandable.clear();
          std::sort(todo.begin(), todo.end(),
                    [](const SizedNode& a, const SizedNode& b) { return a.first < b.first; });
          for (int j = (int)todo.size() - 1; j >= 0; j--) {
            Node ch[4];
            divide(*todo[j].second, ch, kps);
            push_children(ch, nullptr);
            nodes.erase(todo[j].second->self);
            if ((int)nodes.size() >= N) break;
          }
          if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
        }
      }
    }
    result.reserve(nodes.size());
    for (const Node& n : nodes) {
      int best = n.keys[0];
      for (size_t k = 1; k < n.keys.size(); k++)
        if (kps[n.keys[k]].response > kps[best].response) best = n.keys[k];
      result.push_back(kps[best]);
    }
    return result;
  }

  void compute_keypoints() {
    const float W = 30;
    for (int l = 0; l < nlevels; l++) {
      const Image& im = pyr[l];
      const int min_bx = kEdge - 3, min_by = min_bx;
      const int max_bx = im.w - kEdge + 3, max_by = im.h - kEdge + 3;
      std::vector<KeyPoint>& todo = cand[l];
      todo.clear();
      dist[l].clear();
      const float width = (float)(max_bx - min_bx), height = (float)(max_by - min_by);
      const int n_cols = (int)(width / W), n_rows = (int)(height / W);
      if (n_cols < 1 || n_rows < 1) continue;   // reference divides by zero; guarded
      const int w_cell = (int)std::ceil(width / n_cols), h_cell = (int)std::ceil(height / n_rows);
      std::vector<Cand> cell;
      for (int i = 0; i < n_rows; i++) {
        const float ini_y = (float)(min_by + i * h_cell);
        float max_y = ini_y + h_cell + 6;
        if (ini_y >= max_by - 3) continue;
        if (max_y > max_by) max_y = (float)max_by;
        for (int j = 0; j < n_cols; j++) {
          const float ini_x = (float)(min_bx + j * w_cell);
          float max_x = ini_x + w_cell + 6;
          if (ini_x >= max_bx - 6) continue;
          if (max_x > max_bx) max_x = (float)max_bx;
          const uint8_t* sub = im.interior() + (int)ini_y * im.pitch + (int)ini_x;
          const int cw = (int)max_x - (int)ini_x, ch = (int)max_y - (int)ini_y;
          fast_detect(sub, cw, ch, im.pitch, ini_th, cell);
          if (cell.empty()) fast_detect(sub, cw, ch, im.pitch, min_th, cell);
          for (const Cand& c : cell)
            todo.push_back(KeyPoint{(float)c.x + j * w_cell, (float)c.y + i * h_cell, 7.f, -1.f,
                                    (float)c.score, 0, -1});
        }
      }
      dist[l] = distribute(todo, min_bx, max_bx, min_by, max_by, quota[l]);
      const int patch = (int)(kPatchSize * sf[l]);
      for (KeyPoint& kp : dist[l]) {
        kp.x += min_bx;
        kp.y += min_by;
        kp.octave = l;
        kp.size = (float)patch;
      }
    }
    for (int l = 0; l < nlevels; l++)
      for (KeyPoint& kp : dist[l]) kp.angle = ic_angle(pyr[l], kp.x, kp.y);
  }

  float ic_angle(const Image& im, float px, float py) const {
    int m01 = 0, m10 = 0;
    const uint8_t* c = im.interior() + cv_round_f(py) * im.pitch + cv_round_f(px);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
      int vsum = 0, d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int a = c[u + v * im.pitch], b = c[u - v * im.pitch];
        vsum += a - b;
        m10 += u * (a + b);
      }
      m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
  }

  static void describe(const KeyPoint& kp, const uint8_t* img, int pitch, uint8_t* desc) {
    const float factor_pi = (float)(M_PI / 180.f);
    float angle = kp.angle * factor_pi;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = img + cv_round_f(kp.y) * pitch + cv_round_f(kp.x);
    const int8_t* p = cmos_orb_pattern_xy;
    auto val = [&](int i) -> int {
      float x = (float)p[2 * i], y = (float)p[2 * i + 1];
      return c[cv_round_f(x * b + y * a) * pitch + cv_round_f(x * a - y * b)];
    };
    for (int i = 0; i < 32; i++, p += 32) {
      int v = 0;
      for (int k = 0; k < 8; k++) v |= (val(2 * k) < val(2 * k + 1)) << k;
      desc[i] = (uint8_t)v;
    }
  }

  int run(const uint8_t* img, int w, int h, int pitch) {
    out_kps.clear();
    out_desc.clear();
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(img, w, h, pitch);
    compute_keypoints();
    for (int l = 0; l < nlevels; l++) {
      std::vector<KeyPoint>& kps = dist[l];
      if (kps.empty()) continue;
      const Image& im = pyr[l];
      Image& bl = blurred[l];
      bl.w = im.w; bl.h = im.h; bl.pitch = im.w;
      bl.buf.assign((size_t)im.w * im.h, 0);
      gaussian_blur7(im.interior(), im.w, im.h, im.pitch, bl.buf.data(), bl.pitch);
      size_t off = out_desc.size();
      out_desc.resize(off + kps.size() * 32);
      for (size_t i = 0; i < kps.size(); i++) describe(kps[i], bl.buf.data(), bl.pitch, &out_desc[off + i * 32]);
      if (l != 0) {
        float s = sf[l];
        for (KeyPoint& kp : kps) { kp.x *= s; kp.y *= s; }
      }
      out_kps.insert(out_kps.end(), kps.begin(), kps.end());
    }
    return (int)out_kps.size();
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points for ctypes (tests/, bench.py cpu_baseline).
extern "C" {

void* orb_oracle_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
  return new Extractor(nfeatures, scale, nlevels, ini_th, min_th);
}
void orb_oracle_destroy(void* h) { delete (Extractor*)h; }

void orb_oracle_tables(void* h, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2, int* quota,
                       int* umax) {
  Extractor* e = (Extractor*)h;
  for (int i = 0; i < e->nlevels; i++) {
    sf[i] = e->sf[i]; inv_sf[i] = e->inv_sf[i]; sigma2[i] = e->sigma2[i];
    inv_sigma2[i] = e->inv_sigma2[i]; quota[i] = e->quota[i];
  }
  for (int i = 0; i < 16; i++) umax[i] = e->umax[i];
}

int orb_oracle_extract(void* h, const uint8_t* img, int w, int ht, int pitch) {
  return ((Extractor*)h)->run(img, w, ht, pitch);
}
// Copies the result of the last extract; returns the count.
int orb_oracle_result(void* h, void* kps28, uint8_t* desc32) {
  Extractor* e = (Extractor*)h;
  if (kps28) std::memcpy(kps28, e->out_kps.data(), e->out_kps.size() * sizeof(KeyPoint));
  if (desc32) std::memcpy(desc32, e->out_desc.data(), e->out_desc.size());
  return (int)e->out_kps.size();
}
void orb_oracle_level_size(void* h, int level, int* w, int* ht) {
  Extractor* e = (Extractor*)h;
  *w = e->pyr[level].w; *ht = e->pyr[level].h;
}
// Bordered level image: (w+38) x (h+38), tightly packed.
void orb_oracle_level_image(void* h, int level, uint8_t* out) {
  Extractor* e = (Extractor*)h;
  std::memcpy(out, e->pyr[level].buf.data(), e->pyr[level].buf.size());
}
void orb_oracle_level_blurred(void* h, int level, uint8_t* out) {
  Extractor* e = (Extractor*)h;
  std::memcpy(out, e->blurred[level].buf.data(), e->blurred[level].buf.size());
}
// FAST candidates of a level in the reference's vToDistributeKeys order (coords relative to minBorder).
int orb_oracle_level_candidates(void* h, int level, void* kps28) {
  Extractor* e = (Extractor*)h;
  if (kps28) std::memcpy(kps28, e->cand[level].data(), e->cand[level].size() * sizeof(KeyPoint));
  return (int)e->cand[level].size();
}
// Distributed keypoints of a level after orientation (and after pt *= scale for level > 0).
int orb_oracle_level_keypoints(void* h, int level, void* kps28) {
  Extractor* e = (Extractor*)h;
  if (kps28) std::memcpy(kps28, e->dist[level].data(), e->dist[level].size() * sizeof(KeyPoint));
  return (int)e->dist[level].size();
}

// Primitive entry points, for pinning against cv2.
void orb_oracle_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch) {
  resize_linear_u8(src, sw, sh, spitch, dst, dw, dh, dpitch);
}
void orb_oracle_blur7(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) {
  gaussian_blur7(src, w, h, spitch, dst, dpitch);
}
// out: triples (x, y, score); returns count (<= cap written).
int orb_oracle_fast(const uint8_t* img, int w, int h, int pitch, int threshold, int* out, int cap) {
  std::vector<Cand> c;
  fast_detect(img, w, h, pitch, threshold, c);
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
    out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].score;
  }
  return (int)c.size();
}
void orb_oracle_fast_atan2(const float* y, const float* x, float* out, int n) {
  for (int i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]);
}
void orb_oracle_sincos(const float* deg, float* c, float* s, int n) {
  const float factor_pi = (float)(M_PI / 180.f);
  for (int i = 0; i < n; i++) { float a = deg[i] * factor_pi; c[i] = cosf(a); s[i] = sinf(a); }
}
// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1422-1437): bit-hack popcount over 8 x int32.
int orb_oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int32_t pa[8], pb[8];
  std::memcpy(pa, a, 32); std::memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    unsigned int v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

}  // extern "C"
