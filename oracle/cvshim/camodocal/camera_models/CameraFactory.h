/* TEST INFRASTRUCTURE ONLY -- see Camera.h next to this file. */
#ifndef VPL_CVSHIM_CAMODOCAL_FACTORY
#define VPL_CVSHIM_CAMODOCAL_FACTORY
#include "Camera.h"
namespace camodocal {
class CameraFactory {
 public:
  static CameraFactory* instance() {
    static CameraFactory f;
    return &f;
  }
  CameraPtr generateCameraFromYamlFile(const std::string&) { return std::make_shared<Camera>(); }
};
}  // namespace camodocal
#endif
