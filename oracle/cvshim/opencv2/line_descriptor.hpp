/* TEST INFRASTRUCTURE ONLY.  vanishing_point_detection.h:17-18 includes opencv_contrib's
 * line_descriptor but uses nothing of it; empty stand-in. */
