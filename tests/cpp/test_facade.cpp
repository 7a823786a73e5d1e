// C++ facade check: drives LSDDetector / BinaryDescriptor / BinaryDescriptorMatcher and the
// Line / vector<int> seam through libvplines_b200.so on a PGM-less synthetic image pair and
// prints a digest that the Python side compares with the ctypes path.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../vplines-slam_b200/compat/vplines_batch.hpp"
#include "../../vplines-slam_b200/compat/vplines_seam.hpp"

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

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--compile-only") return 0;
  try {
    cv::Mat a, b;
    make_image(a, 320, 240, 0);
    make_image(b, 320, 240, 3);
    vplines::B200LineFrontEnd fe;
    std::vector<vplines::Line> la, lb;
    std::vector<int> prev_to_cur;
    fe.detect(a, la);
    fe.detect(b, lb);
    fe.match(prev_to_cur);
    int matched = 0;
    for (int v : prev_to_cur) matched += v >= 0;
    std::printf("FACADE lines_a=%zu lines_b=%zu matched=%d\n", la.size(), lb.size(), matched);
    // raw surface
    auto det = cv::line_descriptor::LSDDetector::createLSDDetector();
    std::vector<cv::line_descriptor::KeyLine> kl;
    det->detect(a, kl, 2, 2);
    cv::Mat desc;
    cv::line_descriptor::BinaryDescriptor::createBinaryDescriptor()->compute(a, kl, desc);
    std::vector<std::vector<cv::DMatch>> knn;
    cv::line_descriptor::BinaryDescriptorMatcher::createBinaryDescriptorMatcher()->knnMatch(desc, desc, knn, 2);
    int self = 0;
    for (size_t i = 0; i < knn.size(); ++i) self += (!knn[i].empty() && knn[i][0].distance == 0.0f);
    unsigned long digest = 1469598103934665603ul;
    for (int i = 0; i < desc.rows; ++i)
      for (int c = 0; c < 32; ++c) digest = (digest ^ desc.at<unsigned char>(i, c)) * 1099511628211ul;
    std::printf("FACADE keylines=%zu self_matches=%d desc_digest=%lu\n", kl.size(), self, digest);
    // batch driver: 5 frames in batches of 2 over 2 slots, plus the same sequence as two shards
    {
      std::vector<cv::Mat> seq(5);
      std::vector<const uint8_t*> ptrs;
      for (int i = 0; i < 5; ++i) { make_image(seq[(size_t)i], 320, 240, 2 * i); ptrs.push_back(seq[(size_t)i].data); }
      unsigned long dg[2] = {1469598103934665603ul, 1469598103934665603ul};
      long lines[2] = {0, 0}, matched[2] = {0, 0};
      auto mix = [&](int which) {
        return [&, which](int64_t, const vplines::FrameResult& r) {
          lines[which] += (long)r.keylines.size();
          for (uint8_t b : r.descriptors) dg[which] = (dg[which] ^ b) * 1099511628211ul;
          for (const VplDMatch& m : r.matches) { matched[which] += m.trainIdx >= 0; dg[which] = (dg[which] ^ (unsigned long)(m.trainIdx + 7)) * 1099511628211ul; }
        };
      };
      vplines::BatchFrontEnd be(0, 320, 240, 1, 2048, 2, 2);
      be.run(ptrs.data(), 320, 0, 5, 0, 2, 1, mix(0));
      for (int rank = 0; rank < 2; ++rank) {
        int64_t s, e; int halo;
        vplines::shard_range(5, rank, 2, s, e, halo);
        vplines::BatchFrontEnd shard(0, 320, 240, 1, 2048, 2, 2);
        shard.run(ptrs.data(), 320, s, e, halo, 2, 1, mix(1));
      }
      std::printf("FACADE batch lines=%ld matched=%ld digest=%lu sharded_equal=%d\n", lines[0], matched[0], dg[0],
                  (int)(dg[0] == dg[1] && lines[0] == lines[1] && matched[0] == matched[1]));
      // the grouped fast path (contiguous frames, upload ahead, group submits): batches of 2 over 2 slots = 3 groups,
      // whole and as two shards
      {
        std::vector<uint8_t> flat((size_t)5 * 320 * 240);
        for (int i = 0; i < 5; ++i)
          for (int y = 0; y < 240; ++y) std::copy(seq[(size_t)i].ptr(y), seq[(size_t)i].ptr(y) + 320, flat.begin() + ((size_t)i * 240 + y) * 320);
        unsigned long dgg[2] = {1469598103934665603ul, 1469598103934665603ul};
        long lg[2] = {0, 0}, mgm[2] = {0, 0};
        auto mixg = [&](int which) {
          return [&, which](int64_t, const vplines::FrameResult& r) {
            lg[which] += (long)r.keylines.size();
            for (uint8_t b : r.descriptors) dgg[which] = (dgg[which] ^ b) * 1099511628211ul;
            for (const VplDMatch& m : r.matches) { mgm[which] += m.trainIdx >= 0; dgg[which] = (dgg[which] ^ (unsigned long)(m.trainIdx + 7)) * 1099511628211ul; }
          };
        };
        vplines::BatchFrontEnd bg(0, 320, 240, 1, 2048, 2, 2);
        bg.run_grouped(flat.data(), 0, 5, 0, 2, 1, mixg(0));
        for (int rank = 0; rank < 2; ++rank) {
          int64_t s, e; int halo;
          vplines::shard_range(5, rank, 2, s, e, halo);
          vplines::BatchFrontEnd shard(0, 320, 240, 1, 2048, 2, 2);
          shard.run_grouped(flat.data(), s, e, halo, 2, 1, mixg(1));
        }
        std::printf("FACADE grouped equal=%d sharded_equal=%d\n", (int)(dgg[0] == dg[0] && lg[0] == lines[0] && mgm[0] == matched[0]),
                    (int)(dgg[1] == dg[0] && lg[1] == lines[0] && mgm[1] == matched[0]));
      }
      // the in-process multi-GPU driver: one host thread + one context per device (as many devices as the box has,
      // at least two contexts), ordered host gather
      {
        unsigned long dm = 1469598103934665603ul;
        long lm = 0, mm = 0;
        std::vector<int> devs;
        const int nd = vpl_device_count();
        for (int d = 0; d < (nd > 1 ? nd : 2); ++d) devs.push_back(nd > 1 ? d : 0);
        if (devs.size() > 5) devs.resize(5);
        vplines::MultiGpuFrontEnd mg(devs, 320, 240, 1, 2048, 2, 2);
        int64_t expect = 0;
        bool ordered = true;
        mg.run(ptrs.data(), 320, 5, 2, 1, [&](int64_t f, const vplines::FrameResult& r) {
          ordered = ordered && f == expect++;
          lm += (long)r.keylines.size();
          for (uint8_t b : r.descriptors) dm = (dm ^ b) * 1099511628211ul;
          for (const VplDMatch& m : r.matches) { mm += m.trainIdx >= 0; dm = (dm ^ (unsigned long)(m.trainIdx + 7)) * 1099511628211ul; }
        });
        std::printf("FACADE multigpu world=%d equal=%d ordered=%d\n", mg.world(), (int)(dm == dg[0] && lm == lines[0] && mm == matched[0]),
                    (int)(ordered && expect == 5));
      }
    }
    cv::Mat f32(4, 4, CV_32FC1);
    try {
      det->detect(f32, kl, 2, 1);
      std::printf("FACADE depth_check=missing\n");
    } catch (const std::runtime_error& e) {
      std::printf("FACADE depth_check=%s\n", e.what());
    }
  } catch (const std::exception& e) {
    std::printf("FACADE error: %s\n", e.what());
    return 2;
  }
  return 0;
}
