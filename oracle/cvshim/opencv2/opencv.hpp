/*
 * oracle/cvshim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/vpl_oracle.h).
 *
 * A stand-in for the handful of OpenCV-3.4 C++ names that the reference's line primitives
 * library uses, so that the reference's OWN sources
 *     /root/reference/line_matching/src/edline_detector.{h,cpp}, line.h
 * and  line_matching.{h,cpp}, lk_tracker_invoker_2d.cpp, klt.h (+ the first 40 lines of klt.cpp:
 * getImageNormParams and the KLT constructor)
 * compile unmodified, from where they lie, into oracle/_ref/ (recipe: oracle/Makefile target
 * `ref`).  OpenCV C++ itself is not in this image, so the *library* calls are
 * answered here; everything the reference wrote itself (edge drawing walk, least-squares fit,
 * extension, Helmholtz validation, nfa) runs as the reference wrote it.
 *
 * Semantics of each stand-in (what OpenCV does for exactly the argument types the reference
 * passes; anything else aborts):
 *   Sobel(u8 -> CV_16SC1, 3x3), GaussianBlur(u8, 5x5, sigma 1): oracle/orc_prims.c, which is
 *       pinned bit-for-bit against cv2 4.13 (tests/test_oracle_prims.py).
 *   absdiff / add on CV_16S: saturating element-wise;  compare(CMP_LT): 255 / 0 (u8).
 *   threshold(CV_16S, THRESH_TOZERO): src > floor(thresh) ? src : 0          [probed on cv2 4.13]
 *   Mat / scalar  (MatExpr -> convertTo(alpha = 1/s)): saturate_cast<short>(cvRound(v * (float)alpha)),
 *       cvRound = round-half-to-even                                          [probed on cv2 4.13]
 *   Mat_<float> * Mat_<float> (cv::gemm, CV_32F): GEMMSingleMul<float,double> -- double
 *       accumulators, result cast to float (modules/core/src/matmul.cpp, from memory; for the
 *       integer pixel coordinates the reference multiplies, the double sums are exact, so the
 *       accumulation order does not matter).
 *   meanStdDev(CV_16S): integer sum, double sum of squares, mean = s * (1/N),
 *       sd = sqrt(max(q * (1/N) - mean^2, 0))                 [pinned bit-exact on cv2 4.13, 2000 windows]
 *   cvFloor / cvRound(float): the SSE forms (INT_MIN on NaN / overflow, nearest-even).
 *   KLT::calc2D's own OpenCV calls (buildOpticalFlowPyramid, copyMakeBorder) are NOT emulated: the
 *       glue (oracle/ref_linematch_glue.cpp) defines KLT::calc2D on top of the oracle's pyramid and
 *       Scharr code (pinned against cv2.buildOpticalFlowPyramid / cv2.Scharr) and runs the
 *       reference's LKTrackerInvoker2D on it level by level, as klt.cpp:598-627 does.
 *   drawing / GUI calls (only reachable with debug_show > 0) are no-ops.
 *   parallel_for_: runs the body once over the whole range on the calling thread (the order the
 *       reference produces with one thread; with more threads its output order is a race,
 *       edline_detector.cpp:1081-1083).
 */
#ifndef VPL_CVSHIM_OPENCV_HPP
#define VPL_CVSHIM_OPENCV_HPP
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

extern "C" {
void orc_gaussian_blur5(const uint8_t* src, int w, int h, uint8_t* dst);
void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy);
}

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 0
#define CV_8SC1 1
#define CV_16UC1 2
#define CV_16SC1 3
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_AA 16
#define CV_PI 3.1415926535897932384626433832795 /* opencv2/core/cvdef.h */
#define CV_CPU_SSE2 3
#define CV_Assert(expr)                                                        \
  do {                                                                         \
    if (!(expr)) {                                                             \
      std::fprintf(stderr, "cvshim: CV_Assert(%s) failed\n", #expr);           \
      std::abort();                                                            \
    }                                                                          \
  } while (0)
#define CVSHIM_UNSUPPORTED(what)                                               \
  do {                                                                         \
    std::fprintf(stderr, "cvshim: unsupported use: %s\n", what);               \
    std::abort();                                                              \
  } while (0)

namespace cv {

enum { THRESH_TOZERO = 3 };
enum { CMP_LT = 3 };

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  int area() const { return width * height; }
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};
struct Size2f {
  float width, height;
  Size2f() : width(0), height(0) {}
  Size2f(float w, float h) : width(w), height(h) {}
};
template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <class U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
  double ddot(const Point_& o) const { return (double)x * o.x + (double)y * o.y; }
};
template <class T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
static inline Point_<float> operator*(const Point_<float>& a, float b) { return Point_<float>(a.x * b, a.y * b); }
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point;
typedef Point_<int> Point2i;
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(double x_, double y_, double w, double h) : x((int)x_), y((int)y_), width((int)w), height((int)h) {}
};
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double& operator[](int i) { return val[i]; }
  const double& operator[](int i) const { return val[i]; }
};
struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
  int type, maxCount;
  double epsilon;
  TermCriteria() : type(0), maxCount(0), epsilon(0) {}
  TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {}
};
enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
enum { COLOR_GRAY2BGR = 8, FONT_HERSHEY_TRIPLEX = 4 };
static inline bool checkHardwareSupport(int) { return false; }
template <class T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(); }
  Vec(T a, T b) { val[0] = a; val[1] = b; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;
struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
};

static inline size_t cvshimElemSize1(int type) {
  static const size_t sz[8] = {1, 1, 2, 2, 4, 4, 8, 0};
  return sz[type & 7];
}
static inline size_t cvshimElemSize(int type) { return cvshimElemSize1(type) * (size_t)((type >> 3) + 1); }
template <class T>
struct DataType;
template <>
struct DataType<float> {
  enum { type = CV_32FC1 };
};
template <>
struct DataType<int> {
  enum { type = CV_32SC1 };
};
template <>
struct DataType<short> {
  enum { type = CV_16SC1, depth = CV_16S };
};

class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step; /* bytes per row */
  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), type_(0) { create(r, c, type); }
  /* header over caller memory (no copy), like cv::Mat(rows, cols, type, void*, step) */
  Mat(int r, int c, int type, void* ext, size_t step_ = 0)
      : rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c * cvshimElemSize(type)), type_(type) {}
  Mat(Size sz, int type, void* ext)
      : rows(sz.height), cols(sz.width), data((uchar*)ext), step((size_t)sz.width * cvshimElemSize(type)), type_(type) {}
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return; /* same as cv::Mat::create */
    buf_.reset(new std::vector<uchar>((size_t)r * c * cvshimElemSize(type), 0));
    rows = r;
    cols = c;
    type_ = type;
    step = (size_t)c * cvshimElemSize(type);
    data = buf_->data();
  }
  static Mat zeros(double r, double c, int type) { return Mat((int)r, (int)c, type); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return cvshimElemSize(type_); }
  size_t elemSize1() const { return cvshimElemSize1(type_); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  template <class T>
  T* ptr(int r = 0) {
    return (T*)(data + (size_t)r * step);
  }
  template <class T>
  const T* ptr(int r = 0) const {
    return (const T*)(data + (size_t)r * step);
  }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <class T>
  T& at(int r, int c) {
    return ((T*)(data + (size_t)r * step))[c];
  }
  template <class T>
  const T& at(int r, int c) const {
    return ((const T*)(data + (size_t)r * step))[c];
  }
  Mat operator()(const Rect&) const { return *this; } /* only reached from debug drawing */
  Mat& setTo(int v) {
    if (v != 0 && elemSize() != 1) CVSHIM_UNSUPPORTED("Mat::setTo(nonzero) on a multi-byte type");
    for (int r = 0; r < rows; r++) std::memset(ptr(r), v, (size_t)cols * elemSize());
    return *this;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    return m;
  }

 protected:
  int type_;
  std::shared_ptr<std::vector<uchar> > buf_;
};

/* cvRound on the default rounding mode: nearest, ties to even. */
static inline int cvshimRound(float v) { return (int)lrintf(v); }
static inline short cvshimSat16(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

/* MatExpr `m / s` on CV_16S: m.convertTo(dst, CV_16S, 1./s, 0) */
static inline Mat operator/(const Mat& a, double s) {
  if (a.type() != CV_16SC1) CVSHIM_UNSUPPORTED("Mat / scalar on a type other than CV_16S");
  Mat d(a.rows, a.cols, CV_16SC1);
  const float alpha = (float)(1. / s);
  const short* pa = a.ptr<short>();
  short* pd = d.ptr<short>();
  for (size_t i = 0, n = (size_t)a.rows * a.cols; i < n; i++) pd[i] = cvshimSat16(cvshimRound(pa[i] * alpha));
  return d;
}

template <class T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  template <class U>
  Mat_(const Mat_<U>& o) : Mat(o.rows, o.cols, DataType<T>::type) { /* Mat_<T>(const Mat&): convertTo */
    const U* s = o.template ptr<U>();
    T* d = this->template ptr<T>();
    for (size_t i = 0, n = (size_t)o.rows * o.cols; i < n; i++) d[i] = (T)s[i];
  }
  T* operator[](int r) { return this->template ptr<T>(r); }
  const T* operator[](int r) const { return this->template ptr<T>(r); }
  Mat_<T> t() const {
    Mat_<T> d(cols, rows);
    const T* s = this->template ptr<T>();
    T* q = d.template ptr<T>();
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) q[(size_t)j * rows + i] = s[(size_t)i * cols + j];
    return d;
  }
};

/* cv::gemm for CV_32F: GEMMSingleMul<float,double> */
static inline Mat_<float> operator*(const Mat_<float>& a, const Mat_<float>& b) {
  CV_Assert(a.cols == b.rows);
  Mat_<float> d(a.rows, b.cols);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += (double)a[i][k] * (double)b[k][j];
      d[i][j] = (float)s;
    }
  return d;
}
static inline Mat_<float> operator+(const Mat_<float>& a, const Mat_<float>& b) {
  CV_Assert(a.rows == b.rows && a.cols == b.cols);
  Mat_<float> d(a.rows, a.cols);
  const float *pa = a.ptr<float>(), *pb = b.ptr<float>();
  float* pd = d.ptr<float>();
  for (size_t i = 0, n = (size_t)a.rows * a.cols; i < n; i++) pd[i] = pa[i] + pb[i];
  return d;
}

static inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigma) {
  if (src.type() != CV_8UC1 || ksize.width != 5 || ksize.height != 5 || sigma != 1.0)
    CVSHIM_UNSUPPORTED("GaussianBlur other than u8, 5x5, sigma 1");
  Mat d(src.rows, src.cols, CV_8UC1);
  orc_gaussian_blur5(src.data, src.cols, src.rows, d.data);
  dst = d;
}
static inline void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize) {
  if (src.type() != CV_8UC1 || ddepth != CV_16SC1 || ksize != 3 || dx + dy != 1)
    CVSHIM_UNSUPPORTED("Sobel other than u8 -> CV_16S, 3x3, first derivative");
  dst.create(src.rows, src.cols, CV_16SC1);
  std::vector<int16_t> other((size_t)src.rows * src.cols);
  if (dx == 1)
    orc_sobel3(src.data, src.cols, src.rows, dst.ptr<int16_t>(), other.data());
  else
    orc_sobel3(src.data, src.cols, src.rows, other.data(), dst.ptr<int16_t>());
}
static inline void absdiff(const Mat& a, const Mat& b, Mat& dst) {
  if (a.type() != CV_16SC1 || b.type() != CV_16SC1) CVSHIM_UNSUPPORTED("absdiff on a type other than CV_16S");
  dst.create(a.rows, a.cols, CV_16SC1);
  const short *pa = a.ptr<short>(), *pb = b.ptr<short>();
  short* pd = dst.ptr<short>();
  for (size_t i = 0, n = (size_t)a.rows * a.cols; i < n; i++) pd[i] = cvshimSat16(std::abs((int)pa[i] - (int)pb[i]));
}
static inline void add(const Mat& a, const Mat& b, Mat& dst) {
  if (a.type() != CV_16SC1 || b.type() != CV_16SC1) CVSHIM_UNSUPPORTED("add on a type other than CV_16S");
  dst.create(a.rows, a.cols, CV_16SC1);
  const short *pa = a.ptr<short>(), *pb = b.ptr<short>();
  short* pd = dst.ptr<short>();
  for (size_t i = 0, n = (size_t)a.rows * a.cols; i < n; i++) pd[i] = cvshimSat16((int)pa[i] + (int)pb[i]);
}
static inline double threshold(const Mat& src, Mat& dst, double thresh, double /*maxval*/, int type) {
  if (src.type() != CV_16SC1 || type != THRESH_TOZERO) CVSHIM_UNSUPPORTED("threshold other than CV_16S THRESH_TOZERO");
  const int ithresh = (int)std::floor(thresh);
  Mat d(src.rows, src.cols, CV_16SC1);
  const short* ps = src.ptr<short>();
  short* pd = d.ptr<short>();
  for (size_t i = 0, n = (size_t)src.rows * src.cols; i < n; i++) pd[i] = ps[i] > ithresh ? ps[i] : (short)0;
  if (dst.data && dst.rows == src.rows && dst.cols == src.cols && dst.type() == CV_16SC1)
    std::memcpy(dst.data, d.data, (size_t)src.rows * src.cols * 2);
  else
    dst = d;
  return thresh;
}
static inline void compare(const Mat& a, const Mat& b, Mat& dst, int op) {
  if (a.type() != CV_16SC1 || b.type() != CV_16SC1 || op != CMP_LT) CVSHIM_UNSUPPORTED("compare other than CV_16S CMP_LT");
  dst.create(a.rows, a.cols, CV_8UC1);
  const short *pa = a.ptr<short>(), *pb = b.ptr<short>();
  uchar* pd = dst.ptr();
  for (size_t i = 0, n = (size_t)a.rows * a.cols; i < n; i++) pd[i] = pa[i] < pb[i] ? 255 : 0;
}

class ParallelLoopBody {
 public:
  virtual ~ParallelLoopBody() {}
  virtual void operator()(const Range& range) const = 0;
};
static inline void parallel_for_(const Range& range, const ParallelLoopBody& body, double /*nstripes*/ = -1.) {
  if (range.end > range.start) body(range);
}
class Mutex {
 public:
  void lock() {}
  void unlock() {}
};

/* only reached from the reference's unused debug helper writeMat() */
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const std::string&, int) {}
  template <class T>
  FileStorage& operator<<(const T&) {
    return *this;
  }
  void release() {}
};

/* ---- what klt.h / lk_tracker_invoker_2d.cpp / line_matching.cpp need ------------------------ */
static inline int cvFloor(float v) { /* SSE cvFloor: INT_MIN on NaN / overflow */
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  int i = (int)v;
  return i - (v < (float)i);
}
static inline int cvFloor(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  int i = (int)v;
  return i - (v < (double)i);
}
static inline int cvRound(float v) { /* _mm_cvtss_si32 */
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)lrintf(v);
}
static inline int cvRound(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int)lrint(v);
}
static inline int cvRound(int v) { return v; }

template <class T>
class AutoBuffer {
 public:
  explicit AutoBuffer(size_t n) : v_(n) {}
  operator T*() { return v_.data(); }

 private:
  std::vector<T> v_;
};

static inline void meanStdDev(const Mat& src, Mat& mean, Mat& sd) {
  if (src.type() != CV_16SC1) CVSHIM_UNSUPPORTED("meanStdDev on a type other than CV_16SC1");
  long long s = 0;
  double q = 0;
  for (int r = 0; r < src.rows; r++) {
    const short* p = src.ptr<short>(r);
    for (int c = 0; c < src.cols; c++) {
      s += p[c];
      q += (double)p[c] * p[c];
    }
  }
  double scale = 1.0 / ((double)src.rows * src.cols);
  double m = (double)s * scale;
  double var = q * scale - m * m;
  mean.create(1, 1, CV_64FC1);
  sd.create(1, 1, CV_64FC1);
  mean.at<double>(0, 0) = m;
  sd.at<double>(0, 0) = std::sqrt(var > 0 ? var : 0);
}

/* Input/Output array proxies: just enough to carry the argument types Matching() passes */
class _InputArray {
 public:
  enum { NONE = 0, MAT = 1, STD_VECTOR = 3, STD_VECTOR_MAT = 5 };
  _InputArray() : m(nullptr), vp(nullptr), vu(nullptr), vf(nullptr) {}
  _InputArray(const Mat& a) : m(&a), vp(nullptr), vu(nullptr), vf(nullptr) {}
  _InputArray(const std::vector<Point2f>& a) : m(nullptr), vp((std::vector<Point2f>*)&a), vu(nullptr), vf(nullptr) {}
  _InputArray(const std::vector<uchar>& a) : m(nullptr), vp(nullptr), vu((std::vector<uchar>*)&a), vf(nullptr) {}
  _InputArray(const std::vector<float>& a) : m(nullptr), vp(nullptr), vu(nullptr), vf((std::vector<float>*)&a) {}
  int kind() const { return m ? MAT : (vp || vu || vf) ? STD_VECTOR : NONE; }
  bool needed() const { return kind() != NONE; }
  const Mat* m;
  std::vector<Point2f>* vp;
  std::vector<uchar>* vu;
  std::vector<float>* vf;
};
typedef _InputArray _OutputArray;
typedef _InputArray _InputOutputArray;
typedef const _InputArray& InputArray;
typedef const _InputArray& OutputArray;
typedef const _InputArray& InputOutputArray;
static inline const _InputArray& noArray() {
  static _InputArray none;
  return none;
}

class SparsePyrLKOpticalFlow {
 public:
  virtual ~SparsePyrLKOpticalFlow() {}
};

template <class T>
class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
  void release() { this->reset(); }
};
template <class T, class... A>
static inline Ptr<T> makePtr(A&&... a) {
  return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...));
}

/* drawing / GUI: reachable only with debug_show > 0 -- no-ops */
struct RotatedRect {
  RotatedRect(const Point2f&, const Size2f&, float) {}
  void points(Point2f* p) const { for (int i = 0; i < 4; i++) p[i] = Point2f(); }
};
class LineIterator {
 public:
  template <class A, class B>
  LineIterator(const Mat&, A, B, int = 8) : count(0) {}
  int count;
  uchar* operator*() { return dummy_; }
  LineIterator& operator++() { return *this; }
  LineIterator operator++(int) { return *this; }

 private:
  uchar dummy_[4];
};
template <class... A> static inline void line(A&&...) {}
template <class... A> static inline void circle(A&&...) {}
template <class... A> static inline void arrowedLine(A&&...) {}
template <class... A> static inline void putText(A&&...) {}
template <class... A> static inline void cvtColor(A&&...) {}
template <class... A> static inline void addWeighted(A&&...) {}
template <class... A> static inline void imshow(A&&...) {}
template <class... A> static inline bool imwrite(A&&...) { return true; }
static inline int waitKey(int = 0) { return -1; }

}  // namespace cv
#endif
