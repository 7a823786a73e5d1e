// vplines_seam.hpp -- the two private seams of LineFeatureTracker::readImage, re-expressed on
// the line_descriptor facade, so the B200 path can stand where the reference calls
//     void edline_detect(cv::Mat& image, std::vector<Line>& lines, const bool& smoothed)
//     void match_line_match(const cv::Mat& cur_img, const cv::Mat& prev_img, std::vector<Line>& cur_lsd,
//                           std::vector<Line>& prev_lsd, std::vector<int>& line_prev_to_line_cur)
// (/root/reference/feature_tracker/include/linefeature_tracker.h:74-79; called at
// feature_tracker/src/line_feature_tracker.cpp:87 and :115).  `Line` mirrors the POD of
// /root/reference/line_matching/src/line.h:8-17: the tracker reads line_endpoint, length,
// center and line_equation (unit normal, used by SidenessCheck, line_matching.cpp:412-446).
// The gating constants are the ones of the upstream PL-VINS tracker this fork descends from
// (Hamming distance < 30); the reference itself defines none.
#pragma once
#include <cmath>
#include <vector>

#include "line_descriptor.hpp"

namespace vplines {

struct Line {  // line_matching/src/line.h:8-17 (the fields the tracker uses)
  float line_endpoint[4] = {0, 0, 0, 0};
  double line_equation[3] = {0, 0, 0};
  float center[2] = {0, 0};
  float length = 0;
};

inline Line keyline_to_line(const cv::line_descriptor::KeyLine& k) {
  Line l;
  l.line_endpoint[0] = k.startPointX; l.line_endpoint[1] = k.startPointY;
  l.line_endpoint[2] = k.endPointX; l.line_endpoint[3] = k.endPointY;
  l.center[0] = k.pt.x; l.center[1] = k.pt.y;
  l.length = k.lineLength * (k.octave > 0 ? (float)(1 << k.octave) : 1.0f);
  // unit normal (a, b, c) with a*x + b*y + c = 0 through both end points
  double dx = (double)k.endPointX - k.startPointX, dy = (double)k.endPointY - k.startPointY;
  double n = std::sqrt(dx * dx + dy * dy);
  if (n > 0) {
    l.line_equation[0] = dy / n;
    l.line_equation[1] = -dx / n;
    l.line_equation[2] = -(l.line_equation[0] * k.startPointX + l.line_equation[1] * k.startPointY);
  }
  return l;
}

// Holds the KeyLines / descriptors of the frames so that match_line_match needs no recomputation.
class B200LineFrontEnd {
 public:
  int scale = 2, num_octaves = 1;
  float min_line_length = 0.0f;  // config "min_line_length" (euroc_config.yaml:84)
  int max_hamming = 30;

  // stands in for LineFeatureTracker::edline_detect
  void detect(const cv::Mat& image, std::vector<Line>& lines) {
    prev_kl_.swap(cur_kl_);
    std::swap(prev_desc_, cur_desc_);
    std::vector<cv::line_descriptor::KeyLine> kl, keep;
    det_.detect(image, kl, scale, num_octaves);
    for (auto& k : kl)
      if (k.octave == 0 && k.lineLength >= min_line_length) keep.push_back(k);
    cur_kl_ = keep;
    cur_desc_ = cv::Mat();
    if (!cur_kl_.empty()) bd_.compute(image, cur_kl_, cur_desc_);
    lines.clear();
    lines.reserve(cur_kl_.size());
    for (auto& k : cur_kl_) lines.push_back(keyline_to_line(k));
  }

  // stands in for LineFeatureTracker::match_line_match: prev -> cur index, -1 = unmatched
  void match(std::vector<int>& line_prev_to_line_cur) {
    line_prev_to_line_cur.assign(prev_kl_.size(), -1);
    if (prev_kl_.empty() || cur_kl_.empty()) return;
    std::vector<cv::DMatch> m;
    bm_.match(cur_desc_, prev_desc_, m);  // query = current frame, train = previous frame
    std::vector<float> best(prev_kl_.size(), 1e30f);
    for (auto& d : m) {
      if (d.distance >= (float)max_hamming) continue;
      if (d.distance < best[(size_t)d.trainIdx]) {
        best[(size_t)d.trainIdx] = d.distance;
        line_prev_to_line_cur[(size_t)d.trainIdx] = d.queryIdx;
      }
    }
  }

 private:
  cv::line_descriptor::LSDDetector det_;
  cv::line_descriptor::BinaryDescriptor bd_;
  cv::line_descriptor::BinaryDescriptorMatcher bm_;
  std::vector<cv::line_descriptor::KeyLine> cur_kl_, prev_kl_;
  cv::Mat cur_desc_, prev_desc_;
};

}  // namespace vplines
