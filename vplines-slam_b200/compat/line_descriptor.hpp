// line_descriptor.hpp -- header-only C++ facade with the OpenCV-3.4 line_descriptor surface
// (opencv_contrib 3.4 modules/line_descriptor/include/opencv2/line_descriptor/descriptor.hpp)
// over the C ABI of libvplines_b200.so (include/vpl_capi.h).
//
//   cv::line_descriptor::LSDDetector::detect(const Mat&, vector<KeyLine>&, int scale, int numOctaves, const Mat& mask)
//   cv::line_descriptor::BinaryDescriptor::compute(const Mat&, vector<KeyLine>&, Mat& descriptors, bool returnFloatDescr)
//   cv::line_descriptor::BinaryDescriptorMatcher::match / knnMatch(const Mat& q, const Mat& t, ...)
//
// With OpenCV headers available (-DVPL_WITH_OPENCV) the classes take real cv::Mat / cv::DMatch;
// without them a minimal cv::Mat view (data, rows, cols, step, type) is provided so that the
// facade -- and code written against it -- compiles in this image, which has no OpenCV C++.
// Same error behaviour as opencv_contrib: std::runtime_error("Error, depth image!= 0") on a
// non-8-bit image, a printed message and early return on empty inputs.  No CPU fallback: a
// failed vpl_create / vpl_* call throws std::runtime_error with vpl_last_error().
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vpl_capi.h"

#ifdef VPL_WITH_OPENCV
#include <opencv2/core.hpp>
#else
namespace cv {
enum { CV_8U_ = 0, CV_32F_ = 5 };
#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_32FC1 5
#endif
struct Point2f {
  float x = 0, y = 0;
  Point2f() {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 3.402823466e+38f;
  DMatch() {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};
// Minimal owning / non-owning 2-D matrix: enough of cv::Mat for this surface.
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;  // bytes per row
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0)
      : rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((unsigned char*)ext), type_(type) {}
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step);
    data = buf_->data();
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return 1; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
  unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
  const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }

 private:
  static size_t esz(int type) { return (type & 7) == 5 ? 4 : 1; }
  int type_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};
}  // namespace cv
#endif  // VPL_WITH_OPENCV

namespace cv {
namespace line_descriptor {

// cv::line_descriptor::KeyLine -- field for field (and bit for bit: it is VplKeyLine).
struct KeyLine {
  float angle = 0;
  int class_id = -1;
  int octave = 0;
  Point2f pt;
  float response = 0;
  float size = 0;
  float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
  float sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
  float lineLength = 0;
  int numOfPixels = 0;
  Point2f getStartPoint() const { return Point2f(startPointX, startPointY); }
  Point2f getEndPoint() const { return Point2f(endPointX, endPointY); }
  Point2f getStartPointInOctave() const { return Point2f(sPointInOctaveX, sPointInOctaveY); }
  Point2f getEndPointInOctave() const { return Point2f(ePointInOctaveX, ePointInOctaveY); }
};
static_assert(sizeof(KeyLine) == sizeof(VplKeyLine), "KeyLine must alias VplKeyLine");

namespace detail {
// One lazily created context per thread, grown on demand (the reference's detector and
// matcher objects likewise own scratch reused across frames and are not re-entrant).
struct Ctx {
  VplContext* h = nullptr;
  VplConfig cfg{};
  ~Ctx() { if (h) vpl_destroy(h); }
  VplContext* get(int w, int h_, int octaves, int lines) {
    if (h && w <= cfg.max_width && h_ <= cfg.max_height && octaves <= cfg.max_octaves && lines <= cfg.max_lines) return h;
    VplConfig c;
    vpl_default_config(&c);
    c.max_width = std::max(w, h ? cfg.max_width : 0);
    c.max_height = std::max(h_, h ? cfg.max_height : 0);
    c.max_octaves = std::max(octaves, h ? cfg.max_octaves : 1);
    c.max_lines = std::max(lines, h ? cfg.max_lines : 4096);
    c.max_batch = 4;
    c.num_slots = 2;
    if (h) { vpl_destroy(h); h = nullptr; }
    if (vpl_create(&c, &h) != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(nullptr));
    cfg = c;
    return h;
  }
};
inline Ctx& ctx() { static thread_local Ctx c; return c; }
inline void check(VplContext* h, int r) {
  if (r != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(h));
}
inline void check_u8(const Mat& image) {
  if (image.depth() != 0) throw std::runtime_error("Error, depth image!= 0");
  if (image.channels() != 1) throw std::runtime_error("vplines_b200: only CV_8UC1 images (convert with cvtColor first)");
}
}  // namespace detail

class LSDDetector {
 public:
  static std::shared_ptr<LSDDetector> createLSDDetector() { return std::make_shared<LSDDetector>(); }
  void detect(const Mat& image, std::vector<KeyLine>& keylines, int scale, int numOctaves, const Mat& mask = Mat()) {
    detail::check_u8(image);
    if (!mask.empty() && (mask.rows != image.rows || mask.cols != image.cols || mask.depth() != 0))
      throw std::runtime_error("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1");
    const int cap = 1 << 14;
    VplContext* h = detail::ctx().get(image.cols, image.rows, numOctaves, cap);
    std::vector<VplKeyLine> buf((size_t)detail::ctx().cfg.max_lines);
    int32_t count = 0;
    const uint8_t* p[1] = {image.data};
    detail::check(h, vpl_lsd_detect_batch(h, p, 1, image.cols, image.rows, image.step, scale, numOctaves, buf.data(), &count,
                                          (int)buf.size()));
    keylines.clear();
    keylines.reserve((size_t)count);
    for (int i = 0; i < count; ++i) {
      KeyLine k;
      std::memcpy(static_cast<void*>(&k), &buf[(size_t)i], sizeof(k));
      if (!mask.empty() && mask.at<unsigned char>((int)k.startPointY, (int)k.startPointX) == 0 &&
          mask.at<unsigned char>((int)k.endPointY, (int)k.endPointX) == 0)
        continue;
      keylines.push_back(k);
    }
  }
};

class BinaryDescriptor {
 public:
  static std::shared_ptr<BinaryDescriptor> createBinaryDescriptor() { return std::make_shared<BinaryDescriptor>(); }
  int descriptorSize() const { return 32; }
  void compute(const Mat& image, std::vector<KeyLine>& keylines, Mat& descriptors, bool returnFloatDescr = false) const {
    detail::check_u8(image);
    if (keylines.empty()) { std::printf("Error: keypoint list is empty\n"); return; }
    int maxOct = 0;
    for (const KeyLine& k : keylines) maxOct = std::max(maxOct, k.octave);
    const int n = (int)keylines.size();
    VplContext* h = detail::ctx().get(image.cols, image.rows, maxOct + 1, n);
    const uint8_t* p[1] = {image.data};
    int32_t count = n;
    if (returnFloatDescr) {
      descriptors.create(n, 72, CV_32FC1);
      detail::check(h, vpl_lbd_compute_float_batch(h, p, 1, image.cols, image.rows, image.step,
                                                   reinterpret_cast<const VplKeyLine*>(keylines.data()), &count, n,
                                                   reinterpret_cast<float*>(descriptors.data)));
      return;
    }
    descriptors.create(n, 32, CV_8UC1);
    detail::check(h, vpl_lbd_compute_batch(h, p, 1, image.cols, image.rows, image.step,
                                           reinterpret_cast<const VplKeyLine*>(keylines.data()), &count, n, descriptors.data));
  }
};

class BinaryDescriptorMatcher {
 public:
  static std::shared_ptr<BinaryDescriptorMatcher> createBinaryDescriptorMatcher() { return std::make_shared<BinaryDescriptorMatcher>(); }
  void knnMatch(const Mat& queryDescriptors, const Mat& trainDescriptors, std::vector<std::vector<DMatch>>& matches, int k,
                const Mat& mask = Mat(), bool compactResult = false) const {
    matches.clear();
    if (queryDescriptors.rows == 0 || trainDescriptors.rows == 0) { std::printf("Error: descriptors matrices cannot be void\n"); return; }
    if (queryDescriptors.cols != 32 || trainDescriptors.cols != 32 || queryDescriptors.depth() != 0 || trainDescriptors.depth() != 0)
      throw std::runtime_error("vplines_b200: descriptors must be CV_8UC1 with 32 columns");
    const int nq = queryDescriptors.rows, nt = trainDescriptors.rows;
    VplContext* h = detail::ctx().get(8, 8, 1, std::max(nq, nt));
    std::vector<uint8_t> q((size_t)nq * 32), t((size_t)nt * 32);
    for (int i = 0; i < nq; ++i) std::memcpy(&q[(size_t)i * 32], queryDescriptors.ptr(i), 32);
    for (int i = 0; i < nt; ++i) std::memcpy(&t[(size_t)i * 32], trainDescriptors.ptr(i), 32);
    std::vector<VplDMatch> out((size_t)nq * k);
    int32_t cq = nq, ct = nt;
    detail::check(h, vpl_match_batch(h, q.data(), &cq, nq, t.data(), &ct, nt, 1, k, out.data()));
    for (int i = 0; i < nq; ++i) {
      if (!mask.empty() && mask.at<unsigned char>(i) == 0) {
        if (!compactResult) matches.emplace_back();
        continue;
      }
      std::vector<DMatch> row;
      for (int r = 0; r < k; ++r) {
        const VplDMatch& m = out[(size_t)i * k + r];
        if (m.trainIdx >= 0) row.emplace_back(m.queryIdx, m.trainIdx, 0, m.distance);
      }
      matches.push_back(row);
    }
  }
  void match(const Mat& queryDescriptors, const Mat& trainDescriptors, std::vector<DMatch>& matches, const Mat& mask = Mat()) const {
    std::vector<std::vector<DMatch>> knn;
    knnMatch(queryDescriptors, trainDescriptors, knn, 1, mask, true);
    matches.clear();
    for (auto& r : knn)
      if (!r.empty()) matches.push_back(r[0]);
  }
};

}  // namespace line_descriptor
}  // namespace cv
