/* TEST INFRASTRUCTURE ONLY.  feature_tracker/include/parameters.h includes <ros/ros.h> for ROS_INFO / ROS_WARN and a
 * NodeHandle in one prototype; the line tracker itself makes no ROS call.  Logging goes nowhere. */
#ifndef VPL_CVSHIM_ROS
#define VPL_CVSHIM_ROS
#include <cassert> /* the real <ros/ros.h> brings it in: line_feature_tracker.cpp:279 uses assert() */
#include <string>
#include <vector>
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
namespace ros {
class NodeHandle {};
}  // namespace ros
#endif
