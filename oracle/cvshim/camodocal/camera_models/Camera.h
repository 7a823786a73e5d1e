/* TEST INFRASTRUCTURE ONLY.  vanishing_point_detection.h holds a camodocal::CameraPtr member it never
 * touches; this stand-in declares the type. */
#ifndef VPL_CVSHIM_CAMODOCAL_CAMERA
#define VPL_CVSHIM_CAMODOCAL_CAMERA
#include <memory>
namespace camodocal {
class Camera {};
typedef std::shared_ptr<Camera> CameraPtr;
}  // namespace camodocal
#endif
