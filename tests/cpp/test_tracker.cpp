// The reference-named LineFeatureTracker facade (compat/linefeature_tracker_b200.hpp) run as the tracker node runs the
// reference's class: one object, readImage() per frame.  Reads a sequence from argv[1], writes every frame's
// curframe_ to argv[2]; tests/test_gpu_tracker.py compares it with what the reference's own readImage produced.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../vplines-slam_b200/compat/linefeature_tracker_b200.hpp"

using namespace vplines::ref;

static uint32_t g_seed = 0;
static uint32_t seed_now() { return g_seed; }

template <class T>
static bool rd(FILE* f, T* p, size_t n) { return fread(p, sizeof(T), n, f) == n; }
template <class T>
static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--compile-only")) return 0;
  if (argc < 3) { printf("usage: test_tracker in.bin out.bin\n"); return 1; }
  if (vpl_device_count() <= 0) { printf("no CUDA device: the facade has no CPU path\n"); return 2; }
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 1;
  int32_t n, w, h, equalize, max_h, max_v;
  float k[4], minlen, fiterr;
  if (!rd(in, &n, 1) || !rd(in, &w, 1) || !rd(in, &h, 1) || !rd(in, k, 4) || !rd(in, &equalize, 1) || !rd(in, &max_h, 1) ||
      !rd(in, &max_v, 1) || !rd(in, &minlen, 1) || !rd(in, &fiterr, 1))
    return 1;
  std::vector<float> mapx((size_t)w * h), mapy((size_t)w * h);
  std::vector<uint32_t> seeds((size_t)n);
  std::vector<uint8_t> frames((size_t)n * w * h);
  if (!rd(in, mapx.data(), mapx.size()) || !rd(in, mapy.data(), mapy.size()) || !rd(in, seeds.data(), seeds.size()) ||
      !rd(in, frames.data(), frames.size()))
    return 1;
  fclose(in);
  FILE* out = fopen(argv[2], "wb");
  try {
    LineFeatureTracker t;
    // as main() of line_feature_tracker_node.cpp (:196-207)
    EDLineParam param = {5, 1.0f, 30, 5, 2, (int)minlen, (double)fiterr};
    t.line_detctor = EDLineDetector(param);
    t.line_matching = LineMatching();
    t.EQUALIZE = equalize; t.max_h_lines = max_h; t.max_v_lines = max_v;
    t.setIntrinsics(mapx.data(), mapy.data(), w, h, k[0], k[1], k[2], k[3]);
    t.vpdetect.seed_source = seed_now;
    for (int i = 0; i < n; ++i) {
      g_seed = seeds[(size_t)i];
      cv::Mat img(h, w, CV_8UC1, frames.data() + (size_t)i * w * h, (size_t)w);
      t.readImage(img);
      const FrameLines& F = *t.curframe_;
      int32_t hdr[5] = {(int32_t)F.vecLine.size(), (int32_t)F.vps.size(), (int32_t)F.t_cnt.size(), t.lines_exit ? 1 : 0,
                        t.vpdetect.last_status()};
      wr(out, hdr, 5);
      for (const Line& l : F.vecLine) {
        float e[4] = {l.line_endpoint[0], l.line_endpoint[1], l.line_endpoint[2], l.line_endpoint[3]};
        double q[3] = {l.line_equation[0], l.line_equation[1], l.line_equation[2]};
        float c[4] = {l.center[0], l.center[1], l.length, 0.f};
        wr(out, e, 4); wr(out, q, 3); wr(out, c, 4);
      }
      wr(out, F.lineID.data(), F.lineID.size());
      for (const Vector4d& v : F.vps) wr(out, v.v, 4);
      wr(out, F.t_cnt.data(), F.t_cnt.size());
    }
  } catch (const std::exception& e) {
    printf("exception: %s\n", e.what());
    fclose(out);
    return 3;
  }
  fclose(out);
  printf("ok frames=%d\n", n);
  return 0;
}
