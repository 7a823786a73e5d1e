/* oracle/cvshim/glog/logging.h -- TEST INFRASTRUCTURE ONLY.  vanishing_point_detection.cpp includes
 * glog but logs nothing; empty stand-in. */
#ifndef VPL_CVSHIM_GLOG
#define VPL_CVSHIM_GLOG
#endif
