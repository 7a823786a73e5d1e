/* TEST INFRASTRUCTURE ONLY.  Stand-in for camodocal::Camera.  vanishing_point_detection.h holds a CameraPtr it never
 * touches; LineFeatureTracker::readIntrinsicParameter (line_feature_tracker.cpp:26-34) asks the camera for the
 * undistortion maps and K: this stand-in hands back the maps and intrinsics the caller of the reference code put
 * into g_cvshim_intrinsics (oracle/ref_tracker_glue.cpp), so that the tracker runs on given maps without camodocal
 * (which needs Ceres / Eigen, absent from this image). */
#ifndef VPL_CVSHIM_CAMODOCAL_CAMERA
#define VPL_CVSHIM_CAMODOCAL_CAMERA
#include <cstring>
#include <memory>
#include <string>

#include "opencv2/opencv.hpp"
namespace camodocal {
struct CvshimIntrinsics {
  const float* mapx;
  const float* mapy;
  int w, h;
  float fx, fy, cx, cy;
};
extern CvshimIntrinsics g_cvshim_intrinsics;
class Camera {
 public:
  /* Camera::initUndistortRectifyMap(map1, map2): CV_32FC1 maps, returns the 3x3 CV_32F camera matrix */
  cv::Mat initUndistortRectifyMap(cv::Mat& map1, cv::Mat& map2) const {
    const CvshimIntrinsics& I = g_cvshim_intrinsics;
    map1.create(I.h, I.w, CV_32FC1);
    map2.create(I.h, I.w, CV_32FC1);
    std::memcpy(map1.data, I.mapx, (size_t)I.w * I.h * sizeof(float));
    std::memcpy(map2.data, I.mapy, (size_t)I.w * I.h * sizeof(float));
    cv::Mat K(3, 3, CV_32FC1);
    K.at<float>(0, 0) = I.fx; K.at<float>(0, 2) = I.cx;
    K.at<float>(1, 1) = I.fy; K.at<float>(1, 2) = I.cy;
    K.at<float>(2, 2) = 1.f;
    return K;
  }
};
typedef std::shared_ptr<Camera> CameraPtr;
}  // namespace camodocal
#endif
