/* TEST INFRASTRUCTURE ONLY.  linefeature_tracker.h includes <opencv2/features2d.hpp>; what readImage needs beyond
 * opencv.hpp is cv::remap(INTER_LINEAR, float maps) and cv::CLAHE: both answered by the oracle's restatements
 * (oracle/orc_preproc.c), which tests/test_oracle_preproc.py pins bit for bit against cv2 4.13 on the EuRoC
 * undistortion map and on grids that do not divide the image. */
#ifndef VPL_CVSHIM_FEATURES2D
#define VPL_CVSHIM_FEATURES2D
#include "opencv.hpp"
extern "C" {
void orc_remap_linear(const uint8_t* src, int w, int h, const float* mapx, const float* mapy, int dw, int dh, uint8_t* dst);
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles, uint8_t* dst);
}
#ifndef CV_INTER_LINEAR
#define CV_INTER_LINEAR 1
#endif
namespace cv {
static inline void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation) {
  if (interpolation != CV_INTER_LINEAR || src.type() != CV_8UC1 || !src.isContinuous())
    CVSHIM_UNSUPPORTED("remap other than INTER_LINEAR on a continuous CV_8UC1 image");
  Mat out(map1.rows, map1.cols, CV_8UC1);
  orc_remap_linear(src.data, src.cols, src.rows, map1.ptr<float>(), map2.ptr<float>(), map1.cols, map1.rows, out.data);
  dst = out;
}
class CLAHE {
 public:
  CLAHE(double clip, Size grid) : clip_(clip), tiles_(grid.width) {
    if (grid.width != grid.height) CVSHIM_UNSUPPORTED("CLAHE with a non-square tile grid");
  }
  void apply(const Mat& src, Mat& dst) {
    Mat out(src.rows, src.cols, CV_8UC1);
    Mat in = src.isContinuous() ? src : src.clone();
    orc_clahe(in.data, in.cols, in.rows, clip_, tiles_, out.data);
    dst = out;
  }
 private:
  double clip_;
  int tiles_;
};
static inline Ptr<CLAHE> createCLAHE(double clipLimit, Size tileGridSize) { return makePtr<CLAHE>(clipLimit, tileGridSize); }
}  // namespace cv
#endif
