// The reference-named facade (compat/line_matching_b200.hpp: EDLineDetector::EDline,
// LineMatching::Matching) driven the way LineFeatureTracker::readImage drives the reference's
// classes (feature_tracker/src/line_feature_tracker.cpp:87, :115, :291-321); prints a digest that
// the Python side compares with the CPU oracle.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../vplines-slam_b200/compat/line_matching_b200.hpp"
#include "../../vplines-slam_b200/compat/vplines_batch.hpp"

using vplines::ref::EDLineDetector;
using vplines::ref::EDLineParam;
using vplines::ref::Line;
using vplines::ref::LineMatching;
using vplines::ref::Vector3d;
using vplines::ref::vanishing_point_detection;

static uint32_t fixed_seed() { return 1700000123u; }

static void make_image(cv::Mat& m, int w, int h, int shift) {
  m.create(h, w, CV_8UC1);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int v = 110 + ((x / 7 + y / 11) % 3) * 4;
      if (x + shift > 60 && x + shift < 200 && y > 40 && y < 150) v = 40;
      if (y > 170 && y < 178 && x > 30) v = 220;
      if ((x + shift) - y > 150 && (x + shift) - y < 160) v = 230;
      m.at<unsigned char>(y, x) = (unsigned char)v;
    }
}

static unsigned long digest_lines(const std::vector<Line>& l) {
  unsigned long d = 1469598103934665603ul;
  for (const Line& x : l) {
    unsigned char buf[52];
    std::memcpy(buf, x.line_endpoint.data(), 16);
    std::memcpy(buf + 16, x.line_equation.data(), 24);
    std::memcpy(buf + 40, x.center.data(), 8);
    std::memcpy(buf + 48, &x.length, 4);
    for (unsigned char b : buf) d = (d ^ b) * 1099511628211ul;
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--compile-only") return 0;
  try {
    cv::Mat a, b;
    make_image(a, 320, 240, 0);
    make_image(b, 320, 240, 2);
    // the tracker node's detector (line_feature_tracker_node.cpp:203) with min_line_length 20
    EDLineParam param = {5, 1.0f, 30, 5, 2, 20, 1.8};
    EDLineDetector line_detctor(param);
    LineMatching line_matching;
    std::vector<Line> prev_lsd, cur_lsd;
    line_detctor.EDline(a, prev_lsd, true);   // edline_detect(image, lines, smoothed = true), :87
    line_detctor.EDline(b, cur_lsd, true);
    std::vector<int> line_prev_to_line_cur;
    bool ok = line_matching.Matching(a, b, prev_lsd, cur_lsd, line_prev_to_line_cur, *(cv::Mat*)nullptr, *(cv::Mat*)nullptr,
                                     *(cv::Mat*)nullptr, true, true, 0, 0, 0);  // match_line_match, :299-310
    unsigned long dm = 1469598103934665603ul;
    int matched = 0;
    for (int v : line_prev_to_line_cur) { dm = (dm ^ (unsigned long)(v + 7)) * 1099511628211ul; matched += v >= 0; }
    std::printf("REFSEAM ok=%d lines_a=%zu lines_b=%zu digest_a=%lu digest_b=%lu matched=%d match_digest=%lu\n", (int)ok,
                prev_lsd.size(), cur_lsd.size(), digest_lines(prev_lsd), digest_lines(cur_lsd), matched, dm);
    // unsmoothed input (the demo's call, test_edline_detector.cpp:27) and an empty side
    std::vector<Line> blurred, none;
    line_detctor.EDline(a, blurred, false);
    std::vector<int> r;
    bool empty_ok = line_matching.Matching(a, b, none, cur_lsd, r, *(cv::Mat*)nullptr, *(cv::Mat*)nullptr, *(cv::Mat*)nullptr, true, true, 0);
    std::printf("REFSEAM unsmoothed_lines=%zu unsmoothed_digest=%lu empty_returns=%d\n", blurred.size(), digest_lines(blurred),
                (int)empty_ok);
    // the vanishing-point stage as readImage drives it (line_feature_tracker.cpp:233-262): two frames on one object
    vanishing_point_detection vpdetect;
    vpdetect.init(230.0f, 160.0f, 120.0f, 0.5);
    vpdetect.seed_source = fixed_seed;
    for (int frame = 0; frame < 2; ++frame) {
      std::vector<Line>& L = frame ? cur_lsd : prev_lsd;
      std::vector<Vector3d> _vps;
      std::vector<int> local_vp_ids;
      vpdetect.run_vanishing_point_detection(frame ? b : a, L, L, _vps, local_vp_ids);
      unsigned long dv = 1469598103934665603ul;
      for (const Vector3d& v : _vps) {
        unsigned char buf[24];
        std::memcpy(buf, v.v, 24);
        for (unsigned char x : buf) dv = (dv ^ x) * 1099511628211ul;
      }
      unsigned long di = 1469598103934665603ul;
      for (int v : local_vp_ids) di = (di ^ (unsigned long)v) * 1099511628211ul;
      std::printf("REFSEAM vp frame=%d n=%zu vps_digest=%lu ids_digest=%lu status=%d\n", frame, local_vp_ids.size(), dv, di,
                  vpdetect.last_status());
    }
    // C++ batch driver of the fused readImage pipeline: 5 frames in batches of 3 (one-frame overlap), whole run vs two
    // shards (one-frame halo): same lines, matches, vanishing points and labels
    {
      std::vector<cv::Mat> seq(5);
      std::vector<const uint8_t*> ptrs;
      for (int i = 0; i < 5; ++i) { make_image(seq[(size_t)i], 320, 240, 2 * i); ptrs.push_back(seq[(size_t)i].data); }
      const uint32_t seeds[5] = {900, 901, 902, 903, 904};
      VplEDLineParam ed = {5, 1.0f, 30, 5, 2, 20, 1.8};
      VplLineMatchParam lm;
      vpl_linematch_default_param(&lm);
      auto run = [&](int world, unsigned long& dg, long& nlines, long& nmatch, long& nlab) {
        dg = 1469598103934665603ul; nlines = nmatch = nlab = 0;
        for (int r = 0; r < world; ++r) {
          int64_t s0, e0; int halo;
          vplines::shard_range(5, r, world, s0, e0, halo);
          vplines::BatchReadImage be(0, 320, 240, 512, 3, 2, ed, lm, 230.0f, 160.0f, 120.0f);
          be.run(ptrs.data(), 320, seeds, s0, e0, halo, true, [&](int64_t f, const vplines::LineFrameResult& fr) {
            nlines += (long)fr.lines.size();
            auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) dg = (dg ^ b[i]) * 1099511628211ul; };
            mix(&f, sizeof(f));
            for (const VplLine& l : fr.lines) mix(&l, 52);
            for (int v : fr.prev_to_cur) { mix(&v, 4); nmatch += v >= 0; }
            mix(fr.vps, 72);
            for (int v : fr.vp_idx) { mix(&v, 4); nlab += v != 3; }
          });
        }
      };
      unsigned long d1, d2; long a1, b1, c1, a2, b2, c2;
      run(1, d1, a1, b1, c1);
      run(2, d2, a2, b2, c2);
      std::printf("REFSEAM readimage lines=%ld matched=%ld labelled=%ld digest=%lu sharded_equal=%d\n", a1, b1, c1, d1,
                  (int)(d1 == d2 && a1 == a2 && b1 == b2 && c1 == c2));
    }
  } catch (const std::exception& e) {
    std::printf("REFSEAM error: %s\n", e.what());
    return 2;
  }
  return 0;
}
