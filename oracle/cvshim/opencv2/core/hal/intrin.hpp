/* oracle/cvshim: everything lives in <opencv2/opencv.hpp> (see there). TEST INFRASTRUCTURE ONLY. */
#include <opencv2/opencv.hpp>
