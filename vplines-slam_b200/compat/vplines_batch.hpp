// vplines_batch.hpp -- C++ batch driver over the C ABI: streams a frame sequence through
// vpl_frontend_submit / vpl_frontend_collect in pipelined batches (run) or, for frames that lie contiguous in host
// memory, through vpl_frontend_upload / vpl_frontend_submit_group (run_grouped), chaining consecutive batches
// so that every frame is matched against its predecessor, and shards a sequence over GPUs by
// contiguous range with a one-frame halo.  It does for a whole sequence
// what LineFeatureTracker::readImage does per frame
// (/root/reference/feature_tracker/src/line_feature_tracker.cpp:84-126: detect, describe, match
// to the previous frame).  No collective: the host only gathers results.
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vpl_capi.h"

namespace vplines {

struct FrameResult {
  std::vector<VplKeyLine> keylines;
  std::vector<uint8_t> descriptors;  // keylines.size() x 32
  std::vector<VplDMatch> matches;    // keylines.size() x k, vs the previous frame (trainIdx -1: none)
};

// Frames [start, end) owned by `rank`; halo = 1 if frame start-1 must also be processed here so
// that the pair (start-1, start) is matched on exactly one GPU.
inline void shard_range(int64_t n_frames, int rank, int world, int64_t& start, int64_t& end, int& halo) {
  start = n_frames * rank / world;
  end = n_frames * (rank + 1) / world;
  halo = (start > 0 && end > start) ? 1 : 0;
}

class BatchFrontEnd {
 public:
  BatchFrontEnd(int device, int width, int height, int octaves, int max_lines, int max_batch, int num_slots = 2)
      : w_(width), h_(height), octaves_(octaves), cap_(max_lines), batch_(max_batch), slots_(num_slots) {
    VplConfig c;
    vpl_default_config(&c);
    c.device = device; c.max_width = width; c.max_height = height; c.max_octaves = octaves;
    c.max_lines = max_lines; c.max_batch = max_batch; c.num_slots = num_slots;
    if (vpl_create(&c, &ctx_) != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(nullptr));
    kl_.resize((size_t)max_batch * max_lines);
    desc_.resize((size_t)max_batch * max_lines * 32);
    counts_.resize((size_t)max_batch);
  }
  ~BatchFrontEnd() { if (ctx_) vpl_destroy(ctx_); }
  BatchFrontEnd(const BatchFrontEnd&) = delete;
  BatchFrontEnd& operator=(const BatchFrontEnd&) = delete;
  VplContext* context() { return ctx_; }

  // frames: n pointers to w x h CV_8UC1 images with row pitch `stride`.  Processes frames
  // [start - halo, end); calls sink(frame_index, result) for frames [start, end) in order.
  void run(const uint8_t* const* frames, size_t stride, int64_t start, int64_t end, int halo, int scale, int k,
           const std::function<void(int64_t, const FrameResult&)>& sink) {
    matches_.resize((size_t)batch_ * cap_ * std::max(k, 1));
    struct Pending { int slot; int64_t first; int n; };
    std::vector<Pending> pending;
    const int64_t lo = start - halo;
    int slot = 0;
    auto collect = [&](const Pending& p) {
      check(vpl_frontend_collect(ctx_, p.slot, kl_.data(), counts_.data(), cap_, desc_.data(), k > 0 ? matches_.data() : nullptr));
      for (int i = 0; i < p.n; ++i) {
        if (p.first + i < start) continue;  // halo frame: only there to be matched against
        FrameResult r;
        const int c = counts_[(size_t)i];
        r.keylines.assign(kl_.begin() + (size_t)i * cap_, kl_.begin() + (size_t)i * cap_ + c);
        r.descriptors.assign(desc_.begin() + (size_t)i * cap_ * 32, desc_.begin() + ((size_t)i * cap_ + c) * 32);
        if (k > 0) r.matches.assign(matches_.begin() + (size_t)i * cap_ * k, matches_.begin() + ((size_t)i * cap_ + c) * k);
        sink(p.first + i, r);
      }
    };
    for (int64_t f = lo; f < end;) {
      const int n = (int)std::min<int64_t>(batch_, end - f);
      if ((int)pending.size() == slots_) { collect(pending.front()); pending.erase(pending.begin()); }
      check(vpl_frontend_submit(ctx_, slot, frames + f, n, w_, h_, stride, scale, octaves_, k, f > lo ? 1 : 0));
      pending.push_back({slot, f, n});
      slot = (slot + 1) % slots_;
      f += n;
    }
    for (const Pending& p : pending) collect(p);
  }

  // The fast path for frames that lie CONTIGUOUS in host memory (frame f at frames + f * width * height, e.g. a decoded
  // bag segment): the range is pinned once, every batch is copied to the device ahead of its turn on the context's copy
  // stream (vpl_frontend_upload) and the slots' batches are submitted as one group (vpl_frontend_submit_group: their
  // region-engine launches start together and fill the SMs' warp slots); results come back dense.  Same results, same
  // order as run().  On one B200 with 4736-frame batches on two slots: 58 k frames/s against 47 k for run().
  void run_grouped(const uint8_t* frames, int64_t start, int64_t end, int halo, int scale, int k,
                   const std::function<void(int64_t, const FrameResult&)>& sink) {
    const size_t fb = (size_t)w_ * h_;
    const int64_t lo = start - halo;
    if (end <= lo) return;
    struct Batch { int64_t first; int n; };
    std::vector<Batch> bs;
    for (int64_t f = lo; f < end; f += batch_) bs.push_back({f, (int)std::min<int64_t>(batch_, end - f)});
    const size_t rows = (size_t)batch_ * cap_, kk = (size_t)std::max(k, 1);
    std::vector<std::vector<VplKeyLine>> kl((size_t)slots_);
    std::vector<std::vector<uint8_t>> desc((size_t)slots_);
    std::vector<std::vector<VplDMatch>> mt((size_t)slots_);
    std::vector<std::vector<int32_t>> counts((size_t)slots_);
    for (int s = 0; s < slots_; ++s) {
      kl[(size_t)s].resize(rows); desc[(size_t)s].resize(rows * 32); mt[(size_t)s].resize(rows * kk); counts[(size_t)s].resize((size_t)batch_);
    }
    check(vpl_host_register(ctx_, frames + (size_t)lo * fb, (size_t)(end - lo) * fb));
    std::vector<const uint8_t*> ptrs((size_t)batch_);
    auto upload = [&](size_t bi) {
      for (int i = 0; i < bs[bi].n; ++i) ptrs[(size_t)i] = frames + (size_t)(bs[bi].first + i) * fb;
      check(vpl_frontend_upload(ctx_, (int)(bi % (size_t)slots_), ptrs.data(), bs[bi].n, w_, h_, (size_t)w_));
    };
    auto collect = [&](size_t bi) {
      const size_t s = bi % (size_t)slots_;
      int64_t total = 0;
      check(vpl_frontend_collect_dense(ctx_, (int)s, counts[s].data(), kl[s].data(), desc[s].data(), k > 0 ? mt[s].data() : nullptr,
                                       (int64_t)rows, &total));
      size_t off = 0;
      for (int i = 0; i < bs[bi].n; ++i) {
        const size_t c = (size_t)counts[s][(size_t)i];
        if (bs[bi].first + i >= start) {  // (a halo frame is only there to be matched against)
          FrameResult r;
          r.keylines.assign(kl[s].begin() + off, kl[s].begin() + off + c);
          r.descriptors.assign(desc[s].begin() + off * 32, desc[s].begin() + (off + c) * 32);
          if (k > 0) r.matches.assign(mt[s].begin() + off * (size_t)k, mt[s].begin() + (off + c) * (size_t)k);
          sink(bs[bi].first + i, r);
        }
        off += c;
      }
    };
    try {
      // group g is submitted BEFORE group g-1 is collected (it queues behind it on the slots' streams and writes the
      // slots' other result generation): the download of g-1 and this thread's work on it overlap the kernels of g
      const size_t nb = bs.size(), S = (size_t)slots_;
      auto submit = [&](size_t g) {
        const int m = (int)std::min(S, nb - g);
        std::vector<int> slots((size_t)m), ns((size_t)m), chain((size_t)m);
        for (int i = 0; i < m; ++i) { slots[(size_t)i] = i; ns[(size_t)i] = bs[g + (size_t)i].n; chain[(size_t)i] = (g + (size_t)i) > 0; }
        check(vpl_frontend_submit_group(ctx_, m, slots.data(), ns.data(), w_, h_, scale, octaves_, k, chain.data()));
      };
      for (size_t bi = 0; bi < std::min(S, nb); ++bi) upload(bi);
      submit(0);
      for (size_t g = S; g < nb; g += S) {
        for (size_t bi = g; bi < std::min(g + S, nb); ++bi) upload(bi);  // (group g-2 is collected: the buffers are free)
        submit(g);
        for (size_t bi = g - S; bi < g; ++bi) collect(bi);
      }
      for (size_t bi = (nb - 1) / S * S; bi < nb; ++bi) collect(bi);  // the last group
    } catch (...) {
      vpl_sync(ctx_);
      vpl_host_unregister(ctx_, frames + (size_t)lo * fb);
      throw;
    }
    check(vpl_host_unregister(ctx_, frames + (size_t)lo * fb));
  }

 private:
  void check(int r) { if (r != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(ctx_)); }
  VplContext* ctx_ = nullptr;
  int w_, h_, octaves_, cap_, batch_, slots_;
  std::vector<VplKeyLine> kl_;
  std::vector<uint8_t> desc_;
  std::vector<VplDMatch> matches_;
  std::vector<int32_t> counts_;
};

// ---- one process, several GPUs -----------------------------------------------------------
// One host thread + one context + its streams per GPU: the sequence is cut into contiguous frame ranges, GPU g also
// processes frame start_g - 1 (halo) so that every consecutive pair is matched on exactly one GPU, and the host
// gathers: results are delivered to `sink` in frame order after all shards have finished.  No collective, no peer
// copy.  `devices` may name a device more than once (two contexts on one GPU: how the single-GPU tests exercise it).
class MultiGpuFrontEnd {
 public:
  MultiGpuFrontEnd(const std::vector<int>& devices, int width, int height, int octaves, int max_lines, int max_batch,
                   int num_slots = 2) {
    if (devices.empty()) throw std::runtime_error("vplines_b200: no device given");
    for (int d : devices) shards_.emplace_back(new BatchFrontEnd(d, width, height, octaves, max_lines, max_batch, num_slots));
  }
  int world() const { return (int)shards_.size(); }

  // frames [0, n_frames); sink(frame_index, result) is called for every frame, in order, from the calling thread
  void run(const uint8_t* const* frames, size_t stride, int64_t n_frames, int scale, int k,
           const std::function<void(int64_t, const FrameResult&)>& sink) {
    const int G = world();
    std::vector<std::vector<FrameResult>> out((size_t)G);
    std::vector<std::exception_ptr> err((size_t)G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g)
      th.emplace_back([&, g] {
        try {
          int64_t s, e;
          int halo;
          shard_range(n_frames, g, G, s, e, halo);
          out[(size_t)g].reserve((size_t)(e - s));
          shards_[(size_t)g]->run(frames, stride, s, e, halo, scale, k,
                                  [&](int64_t, const FrameResult& r) { out[(size_t)g].push_back(r); });
        } catch (...) {
          err[(size_t)g] = std::current_exception();
        }
      });
    for (std::thread& t : th) t.join();
    for (const std::exception_ptr& e : err)
      if (e) std::rethrow_exception(e);
    int64_t f = 0;
    for (int g = 0; g < G; ++g)
      for (const FrameResult& r : out[(size_t)g]) sink(f++, r);
  }

 private:
  std::vector<std::unique_ptr<BatchFrontEnd>> shards_;
};

// ---- readImage's whole line pipeline over a frame sequence (vpl_readimage_*) ---------------------------------
// What LineFeatureTracker::readImage does per frame (feature_tracker/src/line_feature_tracker.cpp:56-288), for a
// sequence: remap + CLAHE (when set_preprocess was called), EDline, Matching(previous frame, frame), vanishing
// points on the frame's own lines.  Batches, and shards of different ranks, overlap by one frame so that every
// consecutive pair is matched exactly once; the vanishing points of the overlap frame are computed twice and
// delivered once.  seeds[f]: what time(NULL) returned for frame f (the reference seeds rand() with it).
struct LineFrameResult {
  std::vector<VplLine> lines;
  std::vector<int32_t> prev_to_cur;  // one entry per line of the PREVIOUS frame (-1: unmatched); empty for frame 0
  double vps[9];                     // three unit vectors
  std::vector<int32_t> vp_idx;       // one label per line: 0..2, 3 = none
  int32_t vp_status;                 // see vpl_vp_detect_batch
};

class BatchReadImage {
 public:
  BatchReadImage(int device, int width, int height, int max_lines, int max_batch, int num_slots, const VplEDLineParam& ed,
                 const VplLineMatchParam& lm, float f, float cx, float cy)
      : w_(width), h_(height), cap_(max_lines), batch_(max_batch), slots_(num_slots) {
    if (max_batch < 2) throw std::runtime_error("vplines_b200: a batch must hold the overlap frame and one new frame");
    VplConfig c;
    vpl_default_config(&c);
    c.device = device; c.max_width = width; c.max_height = height; c.max_octaves = 1;
    c.max_lines = max_lines; c.max_batch = max_batch; c.num_slots = num_slots; c.lsd_path = 0;
    if (vpl_create(&c, &ctx_) != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(nullptr));
    check(vpl_edlines_configure(ctx_, &ed));
    check(vpl_linematch_configure(ctx_, &lm));
    check(vpl_vp_configure(ctx_, f, cx, cy));
    lines_.resize((size_t)max_batch * max_lines);
    p2c_.resize((size_t)max_batch * max_lines);
    idx_.resize((size_t)max_batch * max_lines);
    counts_.resize((size_t)max_batch);
    status_.resize((size_t)max_batch);
    vps_.resize((size_t)max_batch * 9);
  }
  ~BatchReadImage() { if (ctx_) vpl_destroy(ctx_); }
  BatchReadImage(const BatchReadImage&) = delete;
  BatchReadImage& operator=(const BatchReadImage&) = delete;
  VplContext* context() { return ctx_; }
  // undistortion maps (w x h floats each, or both null) and CLAHE clip limit (<= 0: off), readImage :62-68
  void set_preprocess(const float* mapx, const float* mapy, double clahe_clip, int clahe_tiles = 8) {
    check(vpl_set_preprocess(ctx_, mapx, mapy, w_, h_, clahe_clip, clahe_tiles));
  }

  // Processes frames [start - halo, end); calls sink(frame_index, result) for frames [start, end) in order.
  void run(const uint8_t* const* frames, size_t stride, const uint32_t* seeds, int64_t start, int64_t end, int halo,
           bool smoothed, const std::function<void(int64_t, const LineFrameResult&)>& sink) {
    struct Pending { int slot; int64_t first; int n; bool overlap; };
    std::vector<Pending> pending;
    const int64_t lo = start - halo;
    auto collect = [&](const Pending& p) {
      check(vpl_readimage_collect(ctx_, p.slot, lines_.data(), counts_.data(), cap_, p2c_.data(), vps_.data(), idx_.data(),
                                  status_.data()));
      for (int i = p.overlap ? 1 : 0; i < p.n; ++i) {
        LineFrameResult r;
        const int c = counts_[(size_t)i];
        r.lines.assign(lines_.begin() + (size_t)i * cap_, lines_.begin() + (size_t)i * cap_ + c);
        if (i > 0) r.prev_to_cur.assign(p2c_.begin() + (size_t)i * cap_, p2c_.begin() + (size_t)i * cap_ + counts_[(size_t)i - 1]);
        std::copy(vps_.begin() + (size_t)i * 9, vps_.begin() + (size_t)i * 9 + 9, r.vps);
        r.vp_idx.assign(idx_.begin() + (size_t)i * cap_, idx_.begin() + (size_t)i * cap_ + c);
        r.vp_status = status_[(size_t)i];
        sink(p.first + i, r);
      }
    };
    int slot = 0;
    for (int64_t f = start; f < end;) {
      const bool overlap = f > lo;  // a frame to match against precedes f: the batch starts on it
      const int64_t first = overlap ? f - 1 : f;
      const int n = (int)std::min<int64_t>(batch_, end - first);
      if ((int)pending.size() == slots_) { collect(pending.front()); pending.erase(pending.begin()); }
      // frame_count0 = index of the batch's first frame: only frame 0 of the whole sequence is a first call
      check(vpl_readimage_submit(ctx_, slot, frames + first, n, w_, h_, stride, smoothed ? 1 : 0, seeds + first, (int)first));
      pending.push_back({slot, first, n, overlap});
      slot = (slot + 1) % slots_;
      f = first + n;
    }
    for (const Pending& p : pending) collect(p);
  }

 private:
  void check(int r) { if (r != VPL_OK) throw std::runtime_error(std::string("vplines_b200: ") + vpl_last_error(ctx_)); }
  VplContext* ctx_ = nullptr;
  int w_, h_, cap_, batch_, slots_;
  std::vector<VplLine> lines_;
  std::vector<int32_t> p2c_, idx_, counts_, status_;
  std::vector<double> vps_;
};

}  // namespace vplines
