/* oracle/cvshim stand-in for OpenCV's private lkpyramid.hpp: only deriv_type is used
 * (line_matching/thirdparty/opencv-3.4/.../lkpyramid.hpp:8). TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <opencv2/opencv.hpp>
namespace cv {
namespace detail {
typedef short deriv_type;
}
}
