// capi.cu -- context, memory plan and the C ABI of libvplines_b200.so (see
// include/vpl_capi.h for what each entry point replaces in the reference).
//
// One context = one GPU, `num_slots` independent batch slots.  A slot owns every
// device buffer a batch needs (layout in vpl_common.cuh), a pinned staging area
// and a stream, so consecutive batches overlap: while slot s runs its region
// engine, slot s+1 uploads and runs its streaming kernels.  Nothing is allocated
// after vpl_create.  There is no CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "vpl_common.cuh"

using namespace vpl;

namespace {

thread_local std::string g_create_error;


struct OctBuf {
  uint8_t* pyr = nullptr;
  short2* grad = nullptr;
  uint8_t* scl = nullptr;
  float* ang = nullptr;
  Pix* pix = nullptr;           // 16-byte engine record per scaled pixel
  // speculative region engine (allocated at its first use, for kSpecMaxBatch frames)
  uint32_t* spec_tag = nullptr;
  uint32_t* spec_arena = nullptr;
  EngDesc* eng_desc = nullptr;
  RectCand* eng_rects = nullptr;
  int* ord = nullptr;
  int* n_ord = nullptr;
  uint32_t* reg = nullptr;      // region list of the default engine (4 B per scaled pixel); u16 bins of the ordering kernel before
  RectCand* cand = nullptr;
  int* n_cand = nullptr;
  unsigned int* maxq = nullptr;
};

// What the batch in flight on a slot is (exactly one; every collect requires its own kind).
enum BatchKind { BK_NONE = 0, BK_FRONTEND, BK_EDLINES, BK_LINEFRONT, BK_VP, BK_READIMAGE };

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;        // everything of the batch finished
  cudaEvent_t last_ready = nullptr;  // last-frame descriptors copied to last_desc
  cudaEvent_t match_done = nullptr;  // this slot's match kernels finished
  cudaEvent_t front_done = nullptr;  // group submits: this slot's kernels before the region engine finished
  cudaEvent_t ev[VPL_NUM_STAGES][2];
  bool ev_used[VPL_NUM_STAGES];
  uint8_t* d_img = nullptr;
  // vpl_frontend_upload: the slot's NEXT batch, copied on a stream of its own while the current batch runs
  uint8_t* d_img_next = nullptr;
  cudaEvent_t uploaded = nullptr;
  int staged_n = 0, staged_w = 0, staged_h = 0;
  uint8_t* d_pre = nullptr;   // remap output (pre-processing scratch)
  uint8_t* d_raw = nullptr;   // raw frames as uploaded, kept when pre-processing is on (resident re-runs of the fused path)
  uint8_t* d_lut = nullptr;   // CLAHE tile LUTs, B x 256 tiles max x 256
  OctBuf oct[kMaxOctaves];
  VplKeyLine* d_kl = nullptr;
  int* d_counts = nullptr;
  uint8_t* d_desc = nullptr;
  VplDMatch* d_match = nullptr;
  uint8_t* d_last_desc = nullptr;  // descriptors of the last frame of the batch
  int* d_last_count = nullptr;
  int* d_offsets = nullptr;          // exclusive prefix of counts (n+1), dense output layout
  VplKeyLine* d_kl_dense = nullptr;  // outputs compacted frame after frame for the download
  uint8_t* d_desc_dense = nullptr;
  VplDMatch* d_match_dense = nullptr;
  int* h_offsets = nullptr;
  // Results of a front-end batch as they wait for their collect.  Two generations per slot: a batch staged with
  // vpl_frontend_upload may be submitted while the slot's previous batch is still uncollected (its kernels queue behind
  // the previous batch's on the slot's stream and write the other generation), so that the previous batch's download
  // and the host's turn-around overlap kernels.  res[0] is the slot's own set of buffers, res[1] is allocated on the
  // first such submit.
  struct FrontRes {
    cudaEvent_t done = nullptr;
    int* d_offsets = nullptr; VplKeyLine* d_kl_dense = nullptr; uint8_t* d_desc_dense = nullptr; VplDMatch* d_match_dense = nullptr;
    int* h_counts = nullptr; int* h_offsets = nullptr; int* h_flags = nullptr;
    int n = 0, k = 0;
  } res[2];
  unsigned gen_sub = 0, gen_col = 0;  // front-end batches submitted / collected on this slot (gen_sub - gen_col <= 2)
  int pending_gen[2] = {0, 0};        // which generation batch number i (mod 2) wrote
  bool dense = false;                // this batch's outputs are downloaded in dense form at collect
  int64_t last_d2h_bytes = 0;
  int* d_flags = nullptr;  // [0] candidate overflow, [1] keyline overflow, [2] EDLines line overflow
  // EDLines (allocated by vpl_edlines_configure)
  EdBuffers ed = {};
  VplLine* d_lines = nullptr;
  VplLine* h_lines = nullptr;
  int* h_ed_status = nullptr;
  int ed_smoothed = 1;
  // line matching (allocated by vpl_linematch_configure)
  uint8_t* lm_pyr = nullptr;
  short2* lm_deriv = nullptr;
  LmBuffers lm = {};
  int* h_r2c = nullptr;
  int* d_lm_counts = nullptr;  // line counts of the standalone match call (2 per pair)
  int lm_pairs = 0, lm_pstride = 1;
  VplSegment* d_seg = nullptr;
  int* d_seg_count = nullptr;
  // vanishing points (allocated by vpl_vp_configure)
  VpBuffers vp = {};
  VplLine* d_vp_lines = nullptr;   // B x cap: the hypothesis / vote set
  VplLine* d_vp_all = nullptr;     // B x cap: the classified set
  int* d_vp_n = nullptr;           // 2 x B: n_lines, n_all
  unsigned* d_vp_seeds = nullptr;  // B
  double* d_vps = nullptr;         // B x 9
  int* d_vp_idx = nullptr;         // B x cap
  double* d_line_vps = nullptr;    // B x cap x 4
  int* d_vp_ids = nullptr;         // B x cap: the tracker's line ids (PointCloud packing)
  float* d_cloud = nullptr;        // B x cap x 10
  int* h_vp_ids = nullptr;
  float* h_cloud = nullptr;
  VplLine* h_vp_lines = nullptr;
  VplLine* h_vp_all = nullptr;
  int* h_vp_n = nullptr;
  unsigned* h_vp_seeds = nullptr;
  double* h_vps = nullptr;
  int* h_vp_idx = nullptr;
  double* h_line_vps = nullptr;
  int* h_vp_status = nullptr;      // 2 x B: status, flags
  bool vp_same = true;
  int vp_n = 0, vp_fc0 = 0;
  // the line set the slot's last vanishing-point run classified (vpl_vp_pack_cloud packs exactly that)
  const VplLine* vp_src_lines = nullptr;  // device, B x max_lines
  const int* vp_src_n = nullptr;          // device, B
  const int* vp_src_hn = nullptr;         // host copy of the counts (valid after the collect)
  // pinned host staging
  uint8_t* h_img = nullptr;
  VplKeyLine* h_kl = nullptr;
  int* h_counts = nullptr;
  uint8_t* h_desc = nullptr;
  VplDMatch* h_match = nullptr;
  int* h_flags = nullptr;
  // state of the batch in flight
  int n = 0, w = 0, h = 0, num_octaves = 0, scale = 2, k = 0;
  int m_pairs = 0, m_cap_q = 0, m_cap_t = 0;  // geometry of the last standalone match (vpl_match_run_resident)
  bool in_flight = false;
  int kind = BK_NONE;       // kind of the batch in flight
  int resident = BK_NONE;   // kind of the batch whose frames / lines the slot still holds (run_resident re-runs it)
  int last_consumer = -1;  // slot whose match reads our last_desc
};

}  // namespace

struct VplContext {
  VplConfig cfg;
  int cand_cap = 0;
  int max_k = 8;
  std::vector<Slot> slots;
  std::string err;
  LsdConst lc;
  double stage_ms[VPL_NUM_STAGES];
  int64_t stage_launches[VPL_NUM_STAGES];
  int64_t launches = 0;
  double* d_lgam = nullptr;  // log_gamma table for the NFA kernel
  int lgam_n = 0;
  std::vector<std::pair<const uint8_t*, size_t>> pinned;  // vpl_host_register ranges
  cudaEvent_t mark = nullptr;        // vpl_debug_mark: origin of vpl_debug_timeline
  cudaStream_t down_stream = nullptr;  // dense rows of a collected front-end batch (the slot's stream may be busy with the next one)
  cudaStream_t up_stream = nullptr;  // vpl_frontend_upload: ONE copy stream for all slots, so uploads reach the device in call order
  // optional pre-processing (readImage: remap + CLAHE)
  float* d_mapx = nullptr;
  float* d_mapy = nullptr;
  uint16_t* d_wtab = nullptr;
  int pre_w = 0, pre_h = 0, pre_tiles = 8;
  bool pre_remap = false;
  double pre_clip = 0.0;
  bool ed_ready = false;  // vpl_edlines_configure has run
  VplEDLineParam edp;
  bool lm_ready = false;  // vpl_linematch_configure has run
  VplLineMatchParam lmp;
  bool vp_ready = false;  // vpl_vp_configure has run
  VpParams vpp;
  double* d_vp_lambda = nullptr;
  int engine_ring_cap = 0;  // list entries per lane of the speculative region engine, 0 = 2*ws*hs/32 (vpl_debug_set_engine_ring_cap)
  int engine_kind = 0;      // 0 = warp-cooperative engine (lsd_engine.cu), 1 = speculative engine (lsd_engine_spec.cu)
  float* d_fdesc = nullptr;      // float descriptors of vpl_lbd_compute_float_batch (allocated at its first use)
  float2* d_cssn_lut = nullptr;  // (cosf, sinf) of the level-line angle by gradient differences, kLutN x kLutN
  int prev_slot = -1;  // slot of the previously submitted batch (for chaining)
  bool have_prev = false;
};

namespace {

int fail(VplContext* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CK(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, VPL_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
cudaError_t dmalloc(T** p, size_t count) {
  return cudaMalloc((void**)p, count * sizeof(T) + 256);
}
template <typename T>
cudaError_t hmalloc(T** p, size_t count) {
  return cudaMallocHost((void**)p, count * sizeof(T) + 256);
}

void octave_geom(int w, int h, int o, int& wo, int& ho, int& ws, int& hs) {
  wo = w; ho = h;
  for (int i = 0; i < o; ++i) { wo /= 2; ho /= 2; }
  ws = (int)lrint(wo * 0.8);
  hs = (int)lrint(ho * 0.8);
}

// log-Gamma exactly as LSD computes it (lsd.cpp log_gamma_lanczos / log_gamma_windschitl; the
// reference carries the same two formulas at line_matching/src/edline_detector.h:210-240).
double host_log_gamma(double x) {
  if (x > 15.0)
    return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                              1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    b += q[n] * pow(x, (double)n);
  }
  return a + log(b);
}

// Built with FMA contraction?  (a*b+c differs between fused and unfused here.)
__global__ void fma_probe_kernel(float a, float b, float c, float* out) { *out = a * b + c; }

struct StageTimer {
  Slot& s;
  VplContext* c;
  int stage;
  bool on;
  StageTimer(VplContext* c_, Slot& s_, int stage_) : s(s_), c(c_), stage(stage_), on(c_->cfg.profile != 0) {
    if (on) { cudaEventRecord(s.ev[stage][0], s.stream); }
  }
  void launches(int n) { c->launches += n; c->stage_launches[stage] += n; }
  ~StageTimer() {
    if (on) { cudaEventRecord(s.ev[stage][1], s.stream); s.ev_used[stage] = true; }
  }
};

void harvest_times(VplContext* c, Slot& s) {
  if (!c->cfg.profile) return;
  for (int i = 0; i < VPL_NUM_STAGES; ++i) {
    if (!s.ev_used[i]) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.ev[i][0], s.ev[i][1]) == cudaSuccess) c->stage_ms[i] += ms;
    s.ev_used[i] = false;
  }
}

int check_dims(VplContext* c, int n, int w, int h, int num_octaves, int scale, bool lsd = false) {
  if (!c) return VPL_E_INVALID;
  if (lsd && !c->cfg.lsd_path)
    return fail(c, VPL_E_INVALID, "this context was created with lsd_path = 0 (EDLines / KLT front end only)");
  if (n < 0 || w <= 0 || h <= 0) return fail(c, VPL_E_INVALID, "bad batch/image size n=%d w=%d h=%d", n, w, h);
  if (n > c->cfg.max_batch) return fail(c, VPL_E_CAPACITY, "batch %d > max_batch %d", n, c->cfg.max_batch);
  if (w > c->cfg.max_width || h > c->cfg.max_height)
    return fail(c, VPL_E_CAPACITY, "image %dx%d larger than the context's %dx%d", w, h, c->cfg.max_width,
                c->cfg.max_height);
  if (num_octaves < 1 || num_octaves > c->cfg.max_octaves)
    return fail(c, VPL_E_CAPACITY, "numOctaves %d outside 1..%d", num_octaves, c->cfg.max_octaves);
  if (num_octaves > 1 && scale != 2)
    return fail(c, VPL_E_INVALID, "scale must be 2 with more than one octave (pyrDown), got %d", scale);
  if (scale < 1) return fail(c, VPL_E_INVALID, "scale must be >= 1");
  int wo, ho, ws, hs;
  octave_geom(w, h, num_octaves - 1, wo, ho, ws, hs);
  if (ws < 4 || hs < 4) return fail(c, VPL_E_INVALID, "image too small for %d octaves", num_octaves);
  return VPL_OK;
}

// ---- pipeline pieces (all asynchronous on the slot stream) -------------------
void stage_images(Slot& s, const uint8_t* const* imgs, int n, int w, int h, size_t stride) {
  for (int f = 0; f < n; ++f) {
    uint8_t* dst = s.h_img + (size_t)f * w * h;
    if (stride == (size_t)w) memcpy(dst, imgs[f], (size_t)w * h);
    else
      for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * w, imgs[f] + (size_t)y * stride, (size_t)w);
  }
}

void run_pyramid(VplContext* c, Slot& s, int do_blur) {
  StageTimer t(c, s, VPL_STAGE_PYRAMID);
  launch_blur5_sobel(s.d_img, s.oct[0].pyr, s.oct[0].grad, s.w, s.h, s.n, do_blur, s.stream);
  t.launches(1);
  int wo = s.w, ho = s.h;
  for (int o = 1; o < s.num_octaves; ++o) {
    launch_pyrdown(s.oct[o - 1].pyr, s.oct[o].pyr, wo, ho, s.n, s.stream);
    wo /= 2; ho /= 2;
    launch_sobel(s.oct[o].pyr, s.oct[o].grad, wo, ho, s.n, s.stream);
    t.launches(2);
  }
}

void fill_engine_args(VplContext* c, Slot& s, EngineArgs& a) {
  memset(&a, 0, sizeof(a));
  a.num_octaves = s.num_octaves;
  a.cand_cap = c->cand_cap;
  a.batch = s.n;
  a.lc = c->lc;
  a.overflow = s.d_flags;
  a.lgam = c->d_lgam;
  a.lgam_n = c->lgam_n;
  a.lut = c->d_cssn_lut;
  a.ring_cap = c->engine_ring_cap;  // 0 = default; vpl_debug_set_engine_ring_cap forces small rings (tests of the fallbacks)
  for (int o = 0; o < s.num_octaves; ++o) {
    int wo, ho, ws, hs;
    octave_geom(s.w, s.h, o, wo, ho, ws, hs);
    EngineOct& e = a.oct[o];
    e.ang = s.oct[o].ang; e.ord = s.oct[o].ord; e.n_ord = s.oct[o].n_ord; e.pix = s.oct[o].pix;
    e.tag = s.oct[o].spec_tag; e.arena = s.oct[o].spec_arena; e.desc = s.oct[o].eng_desc; e.rects = s.oct[o].eng_rects;
    e.reg = s.oct[o].reg; e.cand = s.oct[o].cand; e.n_cand = s.oct[o].n_cand;
    e.ws = ws; e.hs = hs;
    e.log_nt = 5 * (log10((double)ws) + log10((double)hs)) / 2 + log10(11.0);
    e.min_reg_size = (int)(-e.log_nt / log10(c->lc.p));
  }
}

// LSD on the pyramid already in s.oct[*].pyr -> candidates validated (cand, n_cand).  Two halves: everything up to the
// pseudo-ordering (streaming kernels), then the region engine and the NFA validation; a group submit puts a barrier
// between them so that the engine launches of the group's slots start together.
void run_lsd_front(VplContext* c, Slot& s) {
  cudaMemsetAsync(s.d_flags, 0, 2 * sizeof(int), s.stream);
  int wo[kMaxOctaves], ho[kMaxOctaves], ws[kMaxOctaves], hs[kMaxOctaves];
  for (int o = 0; o < s.num_octaves; ++o) {
    octave_geom(s.w, s.h, o, wo[o], ho[o], ws[o], hs[o]);
    cudaMemsetAsync(s.oct[o].maxq, 0, (size_t)s.n * sizeof(unsigned int), s.stream);
  }
  {
    StageTimer t(c, s, VPL_STAGE_SCALE);
    for (int o = 0; o < s.num_octaves; ++o)
      launch_scale08(s.oct[o].pyr, s.oct[o].scl, wo[o], ho[o], ws[o], hs[o], s.n, s.stream);
    t.launches(s.num_octaves);
  }
  {
    StageTimer t(c, s, VPL_STAGE_ANGLE);
    for (int o = 0; o < s.num_octaves; ++o)
      launch_ll_angle(s.oct[o].scl, s.oct[o].ang, s.oct[o].pix, c->d_cssn_lut, s.oct[o].maxq, ws[o], hs[o], s.n, c->lc.rho, s.stream);
    t.launches(s.num_octaves);
  }
  {
    StageTimer t(c, s, VPL_STAGE_ORDER);
    for (int o = 0; o < s.num_octaves; ++o)
      launch_order(s.oct[o].scl, s.oct[o].maxq, s.oct[o].ord, s.oct[o].n_ord, s.oct[o].reg,
                   (size_t)ws[o] * hs[o] * sizeof(uint32_t), ws[o], hs[o], s.n, c->lc.rho, s.stream);
    t.launches(s.num_octaves);
  }
}

void run_lsd_back(VplContext* c, Slot& s) {
  EngineArgs a;
  fill_engine_args(c, s, a);
  {
    StageTimer t(c, s, VPL_STAGE_REGION);
    // default: the warp-cooperative engine; vpl_debug_set_engine(ctx, 1): the speculative one (its parking buffers
    // are sized for kSpecMaxBatch frames)
    if (c->engine_kind == 1 && s.n <= kSpecMaxBatch && s.oct[0].spec_tag) launch_region_engine_spec(a, s.stream);
    else launch_region_engine(a, s.stream);
    t.launches(1);
  }
  {
    StageTimer t(c, s, VPL_STAGE_NFA);
    launch_rect_nfa(a, s.stream);
    t.launches(1);
  }
}

void run_lsd(VplContext* c, Slot& s) {
  run_lsd_front(c, s);
  run_lsd_back(c, s);
}

void run_pack(VplContext* c, Slot& s) {
  StageTimer t(c, s, VPL_STAGE_PACK);
  PackArgs p;
  memset(&p, 0, sizeof(p));
  p.num_octaves = s.num_octaves;
  p.scale = s.scale;
  p.cand_cap = c->cand_cap;
  for (int o = 0; o < s.num_octaves; ++o) {
    int wo, ho, ws, hs;
    octave_geom(s.w, s.h, o, wo, ho, ws, hs);
    p.cand[o] = s.oct[o].cand; p.n_cand[o] = s.oct[o].n_cand;
    p.w[o] = wo; p.h[o] = ho;
  }
  launch_pack_keylines(p, s.d_kl, s.d_counts, s.d_flags + 1, c->cfg.max_lines, s.n, s.stream);
  t.launches(1);
}

void run_lbd(VplContext* c, Slot& s, int num_octaves, float* fdesc = nullptr) {
  StageTimer t(c, s, VPL_STAGE_LBD);
  LbdArgs a;
  memset(&a, 0, sizeof(a));
  a.num_octaves = num_octaves;
  for (int o = 0; o < num_octaves; ++o) {
    int wo, ho, ws, hs;
    octave_geom(s.w, s.h, o, wo, ho, ws, hs);
    a.grad[o] = s.oct[o].grad; a.w[o] = wo; a.h[o] = ho;
  }
  launch_lbd(a, s.d_kl, s.d_counts, c->cfg.max_lines, s.d_desc, fdesc, s.n, s.stream);
  t.launches(1);
}

__global__ void fill_nomatch_kernel(VplDMatch* m, const int* counts, int k) {
  int n = counts[0] * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    VplDMatch d;
    d.queryIdx = i / k; d.trainIdx = -1; d.imgIdx = 0; d.distance = 3.402823466e+38f;
    m[i] = d;
  }
}

// exclusive prefix of counts[0..n) -> offsets[0..n] (single block; n <= a few thousand)
__global__ void offsets_kernel(const int* __restrict__ counts, int n, int* __restrict__ offsets) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  int sum = 0;
  for (int i = tid * per; i < min(n, (tid + 1) * per); ++i) sum += counts[i];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) { int v = s_part[i]; s_part[i] = run; run += v; }
    offsets[n] = run;
  }
  __syncthreads();
  int run = s_part[tid];
  for (int i = tid * per; i < min(n, (tid + 1) * per); ++i) { offsets[i] = run; run += counts[i]; }
}

// frame f's rows -> dense arrays at offsets[f]; everything is moved as 32-bit words
__global__ void compact_outputs_kernel(const uint32_t* __restrict__ kl, const uint32_t* __restrict__ desc,
                                       const uint32_t* __restrict__ match, const int* __restrict__ counts,
                                       const int* __restrict__ offsets, int cap, int k, uint32_t* __restrict__ kl_d,
                                       uint32_t* __restrict__ desc_d, uint32_t* __restrict__ match_d) {
  const int f = blockIdx.x, c = counts[f];
  const size_t off = (size_t)offsets[f];
  const uint32_t* a = kl + (size_t)f * cap * 17;
  uint32_t* ad = kl_d + off * 17;
  for (int i = threadIdx.x; i < c * 17; i += blockDim.x) ad[i] = a[i];
  const uint32_t* b = desc + (size_t)f * cap * 8;
  uint32_t* bd = desc_d + off * 8;
  for (int i = threadIdx.x; i < c * 8; i += blockDim.x) bd[i] = b[i];
  if (k > 0) {
    const uint32_t* m = match + (size_t)f * cap * k * 4;
    uint32_t* md = match_d + off * k * 4;
    for (int i = threadIdx.x; i < c * k * 4; i += blockDim.x) md[i] = m[i];
  }
}

// The fused path of one slot in two halves (see run_lsd_front): pyramid .. pseudo-ordering, then engine .. match.
void enqueue_frontend_front(VplContext* c, int slot) {
  Slot& s = c->slots[slot];
  if (c->cfg.profile) {
    // the slot's stage events are about to be re-recorded: bank the previous batch's times
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) {
      cudaStreamSynchronize(s.stream);
      harvest_times(c, s);
    }
  }
  run_pyramid(c, s, c->cfg.blur_first ? 1 : 0);
  run_lsd_front(c, s);
}

int enqueue_frontend_back(VplContext* c, int slot, int k, int chain) {
  Slot& s = c->slots[slot];
  const int cap = c->cfg.max_lines;
  run_lsd_back(c, s);
  run_pack(c, s);
  if (!c->cfg.blur_first) run_pyramid(c, s, 1);  // BinaryDescriptor always blurs its own pyramid
  run_lbd(c, s, s.num_octaves);
  if (k > 0) {
    StageTimer t(c, s, VPL_STAGE_MATCH);
    if (s.n > 1) {
      launch_hamming_knn(s.d_desc + (size_t)cap * 32, s.d_counts + 1, cap, s.d_desc, s.d_counts, cap, s.n - 1, k,
                         s.d_match + (size_t)cap * k, s.stream);
      t.launches(1);
    }
    if (chain && c->have_prev) {
      // frame 0 against the last frame of the previously submitted batch
      Slot& p = c->slots[c->prev_slot];
      if (c->prev_slot != slot) {
        cudaStreamWaitEvent(s.stream, p.last_ready, 0);
        p.last_consumer = slot;
      }
      launch_hamming_knn(s.d_desc, s.d_counts, cap, p.d_last_desc, p.d_last_count, cap, 1, k, s.d_match, s.stream);
    } else {
      fill_nomatch_kernel<<<8, 256, 0, s.stream>>>(s.d_match, s.d_counts, k);
    }
    t.launches(1);
    cudaEventRecord(s.match_done, s.stream);
  }
  // keep the last frame's descriptors for the next batch (after our own chain match
  // has read the previous content, and after any other slot that still reads it)
  if (s.last_consumer >= 0) {
    cudaStreamWaitEvent(s.stream, c->slots[s.last_consumer].match_done, 0);
    s.last_consumer = -1;
  }
  cudaMemcpyAsync(s.d_last_desc, s.d_desc + (size_t)(s.n - 1) * cap * 32, (size_t)cap * 32,
                  cudaMemcpyDeviceToDevice, s.stream);
  cudaMemcpyAsync(s.d_last_count, s.d_counts + (s.n - 1), sizeof(int), cudaMemcpyDeviceToDevice, s.stream);
  cudaEventRecord(s.last_ready, s.stream);
  c->prev_slot = slot;
  c->have_prev = true;
  s.k = k;
  return VPL_OK;
}

int enqueue_frontend(VplContext* c, int slot, int k, int chain) {
  enqueue_frontend_front(c, slot);
  return enqueue_frontend_back(c, slot, k, chain);
}

// The same for several slots at once: the first halves of all, then -- behind a barrier over the group's streams -- the
// second halves, so that the region-engine launches start together and share the SMs with each other (a latency-bound
// kernel that wants every warp slot) instead of with another slot's streaming kernels.
int enqueue_frontend_group(VplContext* c, const int* slots, int n_slots, int k, const int* chain) {
  for (int i = 0; i < n_slots; ++i) {
    Slot& s = c->slots[slots[i]];
    enqueue_frontend_front(c, slots[i]);
    cudaEventRecord(s.front_done, s.stream);
  }
  for (int i = 0; i < n_slots; ++i) {
    Slot& s = c->slots[slots[i]];
    for (int j = 0; j < n_slots; ++j)
      if (j != i) cudaStreamWaitEvent(s.stream, c->slots[slots[j]].front_done, 0);
    int r = enqueue_frontend_back(c, slots[i], k, chain[i]);
    if (r) return r;
  }
  return VPL_OK;
}

int enqueue_download(VplContext* c, Slot& s, bool kl, bool desc, bool match) {
  StageTimer t(c, s, VPL_STAGE_D2H);
  const size_t cap = (size_t)c->cfg.max_lines;
  cudaMemcpyAsync(s.h_counts, s.d_counts, (size_t)s.n * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(s.h_flags, s.d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  if (kl) cudaMemcpyAsync(s.h_kl, s.d_kl, (size_t)s.n * cap * sizeof(VplKeyLine), cudaMemcpyDeviceToHost, s.stream);
  if (desc) cudaMemcpyAsync(s.h_desc, s.d_desc, (size_t)s.n * cap * 32, cudaMemcpyDeviceToHost, s.stream);
  if (match && s.k > 0)
    cudaMemcpyAsync(s.h_match, s.d_match, (size_t)s.n * cap * s.k * sizeof(VplDMatch), cudaMemcpyDeviceToHost,
                    s.stream);
  return VPL_OK;
}

// compaction + download of counts/offsets; the dense rows themselves are fetched at collect,
// once their total size is known
void enqueue_dense_download(VplContext* c, Slot& s, Slot::FrontRes& r) {
  StageTimer t(c, s, VPL_STAGE_D2H);
  const int cap = c->cfg.max_lines;
  offsets_kernel<<<1, 1024, 0, s.stream>>>(s.d_counts, s.n, r.d_offsets);
  compact_outputs_kernel<<<s.n, 256, 0, s.stream>>>(
      reinterpret_cast<const uint32_t*>(s.d_kl), reinterpret_cast<const uint32_t*>(s.d_desc),
      reinterpret_cast<const uint32_t*>(s.d_match), s.d_counts, r.d_offsets, cap, s.k,
      reinterpret_cast<uint32_t*>(r.d_kl_dense), reinterpret_cast<uint32_t*>(r.d_desc_dense),
      reinterpret_cast<uint32_t*>(r.d_match_dense));
  t.launches(2);
  cudaMemcpyAsync(r.h_counts, s.d_counts, (size_t)s.n * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(r.h_offsets, r.d_offsets, (size_t)(s.n + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(r.h_flags, s.d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  r.n = s.n; r.k = s.k;
  s.dense = true;
}

bool in_registered_range(const VplContext* c, const uint8_t* p, size_t bytes) {
  for (auto& r : c->pinned)
    if (p >= r.first && p + bytes <= r.first + r.second) return true;
  return false;
}

// remap + CLAHE of n frames at `src` into d_img (readImage, line_feature_tracker.cpp:62-68)
int run_preprocess(VplContext* c, Slot& s, const uint8_t* src, int n, int w, int h) {
  StageTimer t(c, s, VPL_STAGE_PREPROC);
  const uint8_t* cur = src;
  if (c->pre_remap) {
    launch_remap(src, s.d_pre, c->d_mapx, c->d_mapy, c->d_wtab, w, h, n, s.stream);
    t.launches(1);
    cur = s.d_pre;
  }
  if (c->pre_clip > 0.0) {
    launch_clahe(cur, s.d_img, s.d_lut, w, h, c->pre_clip, c->pre_tiles, n, s.stream);  // in place when cur == d_img
    t.launches(2);
  } else if (cur != s.d_img) {
    CK(c, cudaMemcpyAsync(s.d_img, cur, (size_t)n * w * h, cudaMemcpyDeviceToDevice, s.stream));
  }
  return VPL_OK;
}

// what follows the host->device copy of a batch: the optional pre-processing of the frames now in d_img
int after_upload(VplContext* c, Slot& s, int n, int w, int h) {
  if (c->pre_remap || c->pre_clip > 0.0) {
    if (w != c->pre_w || h != c->pre_h)
      return fail(c, VPL_E_INVALID, "pre-processing was configured for %dx%d images, got %dx%d", c->pre_w, c->pre_h, w, h);
    if (s.d_raw)  // keep the raw frames: the fused path can be re-run on them
      CK(c, cudaMemcpyAsync(s.d_raw, s.d_img, (size_t)n * w * h, cudaMemcpyDeviceToDevice, s.stream));
    return run_preprocess(c, s, s.d_img, n, w, h);
  }
  return VPL_OK;
}

int upload(VplContext* c, Slot& s, const uint8_t* const* imgs, int n, int w, int h, size_t stride) {
  if (stride < (size_t)w) return fail(c, VPL_E_INVALID, "stride %zu < width %d", stride, w);
  for (int f = 0; f < n; ++f)
    if (!imgs || !imgs[f]) return fail(c, VPL_E_INVALID, "null image pointer at frame %d", f);
  // frames that are contiguous and lie in memory pinned through vpl_host_register go to
  // the device directly; anything else is staged through the slot's pinned buffer
  bool direct = (stride == (size_t)w) && in_registered_range(c, imgs[0], (size_t)n * w * h);
  for (int f = 1; direct && f < n; ++f) direct = (imgs[f] == imgs[0] + (size_t)f * w * h);
  if (!direct) stage_images(s, imgs, n, w, h, stride);
  {
    StageTimer t(c, s, VPL_STAGE_H2D);
    CK(c, cudaMemcpyAsync(s.d_img, direct ? imgs[0] : s.h_img, (size_t)n * w * h, cudaMemcpyHostToDevice, s.stream));
  }
  return after_upload(c, s, n, w, h);
}

int finish(VplContext* c, Slot& s) {
  CK(c, cudaStreamSynchronize(s.stream));
  CK(c, cudaGetLastError());
  harvest_times(c, s);
  s.in_flight = false;
  s.kind = BK_NONE;
  return VPL_OK;
}

// The oldest uncollected front-end batch of a slot: wait for ITS end (a younger batch may be queued behind it on the
// slot's stream), mark it collected.  The dense rows are then fetched on the context's download stream.
int finish_front(VplContext* c, Slot& s, Slot::FrontRes*& r) {
  r = &s.res[s.pending_gen[s.gen_col & 1]];
  CK(c, cudaEventSynchronize(r->done));
  CK(c, cudaGetLastError());
  ++s.gen_col;
  if (s.gen_sub == s.gen_col) {  // nothing queued behind it: the slot is idle, its stage events can be read
    harvest_times(c, s);
    s.in_flight = false;
    s.kind = BK_NONE;
  }
  if (!c->down_stream) CK(c, cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
  return VPL_OK;
}

void copy_rows(void* dst, const void* src, const int* counts, int n, size_t cap_src, size_t cap_dst, size_t elem,
               size_t per_line = 1) {
  for (int f = 0; f < n; ++f) {
    size_t cnt = (size_t)counts[f] * per_line;
    memcpy((char*)dst + (size_t)f * cap_dst * per_line * elem, (const char*)src + (size_t)f * cap_src * per_line * elem,
           cnt * elem);
  }
}


// ---- EDLines ---------------------------------------------------------------------
EdGeom ed_geom(int w, int h, const VplEDLineParam& p) {
  EdGeom G;
  memset(&G, 0, sizeof(G));
  G.w = w; G.h = h;
  G.scan = p.scanIntervals;
  G.nW = w > 2 ? (w - 2 + G.scan - 1) / G.scan : 0;
  G.nH = h > 2 ? (h - 2 + G.scan - 1) / G.scan : 0;
  G.bm_words = (G.nW * G.nH + 31) / 32 + 1;
  G.cap_px = (int)((unsigned)w * (unsigned)h / 5);
  G.cap_edges = G.cap_px / 20;
  G.part_cap = (G.cap_px > p.minLineLen ? G.cap_px : p.minLineLen) + 2;
  G.nslots = 2 * G.cap_px / p.minLineLen + 1;
  G.min_len = p.minLineLen;
  G.logNT = 2.0 * (log10((double)(unsigned)w) + log10((double)(unsigned)h));
  return G;
}

void ed_free(Slot& s) {
  cudaFree(s.ed.gmap); cudaFree(s.ed.bitmap); cudaFree(s.ed.n_anchor); cudaFree(s.ed.first); cudaFree(s.ed.second);
  cudaFree(s.ed.xy); cudaFree(s.ed.sid); cudaFree(s.ed.n_chain); cudaFree(s.ed.n_px); cudaFree(s.ed.status);
  cudaFree(s.ed.slots); cudaFree(s.ed.slot_valid); cudaFree(s.d_lines);
  cudaFreeHost(s.h_lines); cudaFreeHost(s.h_ed_status);
  s.ed = EdBuffers{};
  s.d_lines = nullptr; s.h_lines = nullptr; s.h_ed_status = nullptr;
}

// gradient map -> anchors -> edge chains -> lines, all asynchronous on the slot stream
void run_edlines(VplContext* c, Slot& s) {
  const VplEDLineParam& p = c->edp;
  const EdGeom G = ed_geom(s.w, s.h, p);
  const int cap = c->cfg.max_lines;
  cudaMemsetAsync(s.d_flags + 2, 0, sizeof(int), s.stream);
  {
    StageTimer t(c, s, VPL_STAGE_ED_GRAD);
    // ed.cpp:82-83: the 5x5 blur only when the caller's image is not smoothed yet (-> oct[0].pyr);
    // then Sobel pair + gradient map + anchors in one pass (ed.cpp:125-164)
    const uint8_t* smooth = s.d_img;
    if (!s.ed_smoothed) {
      launch_blur5_sobel(s.d_img, s.oct[0].pyr, s.oct[0].grad, s.w, s.h, s.n, 1, s.stream);
      smooth = s.oct[0].pyr;
      t.launches(1);
    }
    launch_ed_grad_anchor(smooth, s.oct[0].grad, s.ed, G, (int)(short)p.gradientThreshold,
                          (int)(unsigned char)p.anchorThreshold, s.n, s.stream);
    t.launches(1);
  }
  {
    StageTimer t(c, s, VPL_STAGE_ED_WALK);
    launch_ed_walk(s.ed, G, s.n, s.stream);
    t.launches(1);
  }
  {
    StageTimer t(c, s, VPL_STAGE_ED_FIT);
    launch_ed_fit(s.ed, G, p.lineFitErrThreshold, s.oct[0].grad, c->d_lgam, s.n, s.stream);
    launch_ed_compact(s.ed, G, s.d_lines, s.d_counts, cap, s.d_flags + 2, s.n, s.stream);
    t.launches(2);
  }
}

void ed_enqueue_download(VplContext* c, Slot& s) {
  StageTimer t(c, s, VPL_STAGE_D2H);
  const size_t cap = (size_t)c->cfg.max_lines;
  cudaMemcpyAsync(s.h_counts, s.d_counts, (size_t)s.n * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(s.h_ed_status, s.ed.status, (size_t)s.n * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(s.h_flags, s.d_flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  cudaMemcpyAsync(s.h_lines, s.d_lines, (size_t)s.n * cap * sizeof(VplLine), cudaMemcpyDeviceToHost, s.stream);
  s.last_d2h_bytes = (int64_t)((size_t)s.n * cap * sizeof(VplLine) + (size_t)s.n * 8 + 12);
}

int ed_check(VplContext* c, int n, int w, int h, int smoothed) {
  int r = check_dims(c, n, w, h, 1, 1);
  if (r) return r;
  if (!c->ed_ready) return fail(c, VPL_E_INVALID, "call vpl_edlines_configure first");
  if (w < 3 || h < 3 || w > 65535 || h > 65535) return fail(c, VPL_E_INVALID, "EDLines: image size %dx%d unsupported", w, h);
  if (!smoothed && !(c->edp.ksize == 5 && c->edp.sigma == 1.0f))
    return fail(c, VPL_E_INVALID, "EDLines: smoothed=0 needs ksize 5 / sigma 1 (got %d / %g)", c->edp.ksize, (double)c->edp.sigma);
  return VPL_OK;
}

int ed_deliver(VplContext* c, Slot& s, VplLine* lines, int32_t* counts, int cap, int32_t* status) {
  if (s.h_flags[2]) return fail(c, VPL_E_CAPACITY, "a frame produced more than max_lines=%d lines", c->cfg.max_lines);
  for (int f = 0; f < s.n; ++f) {
    if (s.h_counts[f] > cap) return fail(c, VPL_E_CAPACITY, "frame %d has %d lines > cap %d", f, s.h_counts[f], cap);
    counts[f] = s.h_counts[f];
    if (status) status[f] = s.h_ed_status[f];
  }
  copy_rows(lines, s.h_lines, s.h_counts, s.n, (size_t)c->cfg.max_lines, (size_t)cap, sizeof(VplLine));
  return VPL_OK;
}


// ---- line matching -----------------------------------------------------------------
KltGeom klt_geom(int w, int h, int max_level) {
  KltGeom G;
  memset(&G, 0, sizeof(G));
  G.pad = 13;
  G.padx = 16;
  int cw = w, ch = h, level;
  size_t off = 0;
  if (max_level > kKltMaxLevels - 1) max_level = kKltMaxLevels - 1;
  for (level = 0; level <= max_level; ++level) {
    G.w[level] = cw; G.h[level] = ch; G.stride[level] = (G.padx + cw + G.pad + 15) & ~15;
    G.img_off[level] = off; G.deriv_off[level] = off;
    off += (size_t)G.stride[level] * (ch + 2 * G.pad);
    off = (off + 63) & ~(size_t)63;
    int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
    // cv::buildOpticalFlowPyramid stops when the next level would not exceed the window
    if (nw <= G.pad || nh <= G.pad || level == max_level) break;
    cw = nw; ch = nh;
  }
  G.top = level > max_level ? max_level : level;
  G.img_frame = off; G.deriv_frame = off;
  return G;
}

LmParams lm_params(const VplLineMatchParam& p) {
  LmParams P;
  memset(&P, 0, sizeof(P));
  P.step = p.step; P.closest = p.closest_line_threshold; P.ratio = p.line_matching_ratio;
  P.dist_ratio = p.line_distance_error_ratio; P.klt_err = p.klt_error_threshold;
  // KLT's constructor clamps and squares the criteria (klt.cpp:22-36)
  P.max_count = p.max_count < 0 ? 0 : p.max_count > 100 ? 100 : p.max_count;
  double e = p.epsilon < 0 ? 0 : p.epsilon > 10 ? 10 : p.epsilon;
  P.eps2 = e * e;
  P.min_eig = p.min_eig;
  P.topo_dist = p.topo_distance_threshold; P.topo_len = p.topo_length_ratio; P.topo_viol = p.topo_violation_ratio;
  P.illum = p.illumination_adapt != 0; P.topo = p.topological_filter != 0;
  return P;
}

void lm_free(Slot& s) {
  cudaFree(s.lm_pyr); cudaFree(s.lm_deriv); cudaFree(s.lm.kps); cudaFree(s.lm.nxt); cudaFree(s.lm.status);
  cudaFree(s.lm.err); cudaFree(s.lm.kp2line); cudaFree(s.lm.kp_start); cudaFree(s.lm.n_kp); cudaFree(s.lm.r2c);
  cudaFree(s.lm.matched); cudaFree(s.lm.pair_off); cudaFree(s.lm.queue); cudaFree(s.d_lm_counts);
  cudaFreeHost(s.h_r2c);
  s.lm_pyr = nullptr; s.lm_deriv = nullptr; s.lm = LmBuffers{}; s.h_r2c = nullptr; s.d_lm_counts = nullptr;
}

// pyramids of the n frames in d_img, anchors, LK, votes: pair p = (frame p*pstride, the next one)
void run_linematch(VplContext* c, Slot& s, const int* d_counts, int n_frames, int n_pairs, int pstride) {
  const KltGeom G = klt_geom(s.w, s.h, c->lmp.max_level);
  const LmParams P = lm_params(c->lmp);
  const int cap = c->cfg.max_lines;
  s.lm.overflow = s.d_flags + 3;
  cudaMemsetAsync(s.d_flags + 3, 0, sizeof(int), s.stream);
  s.lm_pairs = n_pairs; s.lm_pstride = pstride;
  if (n_pairs <= 0) return;
  {
    StageTimer t(c, s, VPL_STAGE_LM_PYRAMID);
    launch_klt_pyramid(s.d_img, s.lm_pyr, s.lm_deriv, G, s.w, s.h, n_frames, s.stream);
    t.launches(2 * (G.top + 1));
  }
  {
    StageTimer t(c, s, VPL_STAGE_LM_TRACK);
    launch_lm_anchors(s.d_lines, d_counts, cap, s.lm, P, pstride, n_pairs, s.stream);
    launch_klt_track(s.lm_pyr, s.lm_deriv, G, s.lm, P, pstride, n_pairs, s.stream);
    t.launches(3 + G.top);
  }
  {
    StageTimer t(c, s, VPL_STAGE_LM_VOTE);
    launch_lm_vote(s.d_lines, d_counts, cap, s.lm, P, pstride, n_pairs, s.stream);
    t.launches(1);
  }
}

void vp_free(Slot& s) {
  cudaFree(s.vp.para); cudaFree(s.vp.length); cudaFree(s.vp.orient); cudaFree(s.vp.vp1); cudaFree(s.vp.pairs);
  cudaFree(s.vp.rng); cudaFree(s.vp.status); cudaFree(s.vp.grid); cudaFree(s.vp.grid_new);
  cudaFree(s.vp.part_best); cudaFree(s.vp.part_idx); cudaFree(s.vp.best_idx); cudaFree(s.vp.lx);
  cudaFree(s.d_vp_lines); cudaFree(s.d_vp_all); cudaFree(s.d_vp_n); cudaFree(s.d_vp_seeds); cudaFree(s.d_vps);
  cudaFree(s.d_vp_idx); cudaFree(s.d_line_vps); cudaFree(s.d_vp_ids); cudaFree(s.d_cloud);
  cudaFreeHost(s.h_vp_ids); cudaFreeHost(s.h_cloud);
  s.d_vp_ids = nullptr; s.d_cloud = nullptr; s.h_vp_ids = nullptr; s.h_cloud = nullptr;
  cudaFreeHost(s.h_vp_lines); cudaFreeHost(s.h_vp_all); cudaFreeHost(s.h_vp_n); cudaFreeHost(s.h_vp_seeds);
  cudaFreeHost(s.h_vps); cudaFreeHost(s.h_vp_idx); cudaFreeHost(s.h_line_vps); cudaFreeHost(s.h_vp_status);
  s.vp = VpBuffers{};
  s.d_vp_lines = s.d_vp_all = nullptr; s.d_vp_n = nullptr; s.d_vp_seeds = nullptr; s.d_vps = nullptr;
  s.d_vp_idx = nullptr; s.d_line_vps = nullptr;
  s.h_vp_lines = s.h_vp_all = nullptr; s.h_vp_n = nullptr; s.h_vp_seeds = nullptr; s.h_vps = nullptr;
  s.h_vp_idx = nullptr; s.h_line_vps = nullptr; s.h_vp_status = nullptr;
}

// the stage on the n frames whose lines sit in d_vp_lines / d_vp_all
void run_vp_on(VplContext* c, Slot& s, const VplLine* lines, const int* n_lines, const VplLine* all, const int* n_all,
               int n_frames, int frame_count0) {
  const int cap = c->cfg.max_lines;
  {
    StageTimer t(c, s, VPL_STAGE_VP_PREP);
    launch_vp_prepare(lines, n_lines, cap, s.d_vp_seeds, s.vp, c->vpp, n_frames, s.stream);
    t.launches(1);
  }
  {
    StageTimer t(c, s, VPL_STAGE_VP_VOTE);
    launch_vp_vote(n_lines, cap, s.vp, c->vpp, n_frames, s.stream);
    t.launches(2);
  }
  {
    StageTimer t(c, s, VPL_STAGE_VP_SCORE);
    launch_vp_score(s.vp, c->vpp, n_frames, s.stream);
    t.launches(1);
  }
  {
    StageTimer t(c, s, VPL_STAGE_VP_CLASSIFY);
    launch_vp_classify(all, n_all, cap, frame_count0, s.vp, c->vpp, n_frames, s.d_vps, s.d_vp_idx, s.d_line_vps, s.stream);
    t.launches(1);
  }
}
void run_vp(VplContext* c, Slot& s) {
  const int B = c->cfg.max_batch;
  run_vp_on(c, s, s.d_vp_lines, s.d_vp_n, s.vp_same ? s.d_vp_lines : s.d_vp_all, s.vp_same ? s.d_vp_n : s.d_vp_n + B, s.vp_n,
            s.vp_fc0);
}

int lm_check(VplContext* c) {
  if (!c) return VPL_E_INVALID;
  if (!c->lm_ready) return fail(c, VPL_E_INVALID, "call vpl_linematch_configure first");
  return VPL_OK;
}

}  // namespace

extern "C" {

const char* vpl_version(void) { return "vplines_b200 0.1 (sm_100a)"; }

int vpl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void vpl_default_config(VplConfig* cfg) {
  if (!cfg) return;
  cfg->device = 0;
  cfg->max_width = 752;
  cfg->max_height = 480;
  cfg->max_octaves = 1;
  cfg->max_lines = 2048;
  cfg->max_batch = 64;
  cfg->num_slots = 2;
  cfg->blur_first = 1;
  cfg->profile = 0;
  cfg->lsd_path = 1;
}

const char* vpl_last_error(const VplContext* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void vpl_destroy(VplContext* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  for (auto& r : c->pinned) cudaHostUnregister((void*)r.first);
  cudaFree(c->d_mapx); cudaFree(c->d_mapy); cudaFree(c->d_wtab);
  cudaFree(c->d_lgam);
  cudaFree(c->d_cssn_lut);
  cudaFree(c->d_fdesc);
  cudaFree(c->d_vp_lambda);
  for (Slot& s : c->slots) {
    cudaFree(s.d_img); cudaFree(s.d_pre); cudaFree(s.d_lut); cudaFree(s.d_raw);
    cudaFree(s.d_img_next);
    if (s.uploaded) cudaEventDestroy(s.uploaded);
    {
      Slot::FrontRes& r1 = s.res[1];
      cudaFree(r1.d_offsets); cudaFree(r1.d_kl_dense); cudaFree(r1.d_desc_dense); cudaFree(r1.d_match_dense);
      cudaFreeHost(r1.h_counts); cudaFreeHost(r1.h_offsets); cudaFreeHost(r1.h_flags);
      if (r1.done) cudaEventDestroy(r1.done);
    }
    for (int o = 0; o < kMaxOctaves; ++o) {
      OctBuf& b = s.oct[o];
      cudaFree(b.pyr); cudaFree(b.grad); cudaFree(b.scl); cudaFree(b.ang); cudaFree(b.pix); cudaFree(b.spec_tag); cudaFree(b.spec_arena); cudaFree(b.eng_desc); cudaFree(b.eng_rects); cudaFree(b.ord);
      cudaFree(b.n_ord); cudaFree(b.reg); cudaFree(b.cand); cudaFree(b.n_cand); cudaFree(b.maxq);
    }
    cudaFree(s.d_kl); cudaFree(s.d_counts); cudaFree(s.d_desc); cudaFree(s.d_match); cudaFree(s.d_last_desc);
    cudaFree(s.d_last_count); cudaFree(s.d_flags); cudaFree(s.d_seg); cudaFree(s.d_seg_count);
    cudaFree(s.d_offsets); cudaFree(s.d_kl_dense); cudaFree(s.d_desc_dense); cudaFree(s.d_match_dense);
    ed_free(s);
    lm_free(s);
    vp_free(s);
    cudaFreeHost(s.h_offsets);
    cudaFreeHost(s.h_img); cudaFreeHost(s.h_kl); cudaFreeHost(s.h_counts); cudaFreeHost(s.h_desc);
    cudaFreeHost(s.h_match); cudaFreeHost(s.h_flags);
    if (s.done) cudaEventDestroy(s.done);
    if (s.last_ready) cudaEventDestroy(s.last_ready);
    if (s.match_done) cudaEventDestroy(s.match_done);
    if (s.front_done) cudaEventDestroy(s.front_done);
    for (int i = 0; i < VPL_NUM_STAGES; ++i)
      for (int j = 0; j < 2; ++j)
        if (s.ev[i][j]) cudaEventDestroy(s.ev[i][j]);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (c->up_stream) cudaStreamDestroy(c->up_stream);
  if (c->down_stream) cudaStreamDestroy(c->down_stream);
  if (c->mark) cudaEventDestroy(c->mark);
  delete c;
}

int vpl_create(const VplConfig* cfg, VplContext** out) {
  if (!cfg || !out) return fail(nullptr, VPL_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->max_width < 8 || cfg->max_height < 8 || cfg->max_octaves < 1 || cfg->max_octaves > kMaxOctaves ||
      cfg->max_lines < 1 || cfg->max_lines > (1 << 20) - 1 || cfg->max_batch < 1 || cfg->num_slots < 1 ||
      cfg->num_slots > 4)
    return fail(nullptr, VPL_E_INVALID, "bad configuration");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, VPL_E_NODEVICE,
                "no usable CUDA device (count=%d, requested %d, %s): libvplines_b200 has no CPU path", ndev,
                cfg->device, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  VplContext* c = new VplContext();
  c->cfg = *cfg;
  c->cand_cap = 4 * cfg->max_lines;
  memset(c->stage_ms, 0, sizeof(c->stage_ms));
  memset(c->stage_launches, 0, sizeof(c->stage_launches));
  c->lc.prec = VPL_PI * 22.5 / 180;
  c->lc.p = 22.5 / 180;
  c->lc.rho = 2.0 / sin(c->lc.prec);
#define CKC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      int r_ = fail(nullptr, VPL_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));          \
      vpl_destroy(c);                                                                              \
      return r_;                                                                                   \
    }                                                                                              \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  // refuse to run if the build contracted a*b+c (the exact float sequences need -fmad=false)
  {
    float* d = nullptr;
    CKC(cudaMalloc((void**)&d, sizeof(float)));
    // (1+2^-12)^2 = 1 + 2^-11 + 2^-24 rounds to 1 + 2^-11: unfused a*b+c is 0, fused 2^-24
    const float a = 1.0f + 1.0f / 4096.0f, cc = -(1.0f + 1.0f / 2048.0f);
    fma_probe_kernel<<<1, 1>>>(a, a, cc, d);
    float r = 1.0f;
    CKC(cudaMemcpy(&r, d, sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (r != 0.0f) {
      vpl_destroy(c);
      return fail(nullptr, VPL_E_INVALID, "library built with FMA contraction (needs nvcc -fmad=false)");
    }
  }
  lbd_init_tables();
  {
    // a rectangle holds at most every pixel of the largest 0.8-scaled image
    c->lgam_n = (int)((double)cfg->max_width * cfg->max_height * 0.64) + 2 * (cfg->max_width + cfg->max_height) + 16;
    std::vector<double> tab((size_t)c->lgam_n, 0.0);
    for (int m = 1; m < c->lgam_n; ++m) tab[m] = host_log_gamma((double)m);
    CKC(cudaMalloc((void**)&c->d_lgam, tab.size() * sizeof(double)));
    CKC(cudaMemcpy(c->d_lgam, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (cfg->lsd_path) {
      CKC(cudaMalloc((void**)&c->d_cssn_lut, (size_t)kLutN * kLutN * sizeof(float2)));
      launch_cssn_lut(c->d_cssn_lut, 0);
      CKC(cudaGetLastError());
    }
  }
  const size_t B = (size_t)cfg->max_batch, cap = (size_t)cfg->max_lines;
  const size_t P0 = (size_t)cfg->max_width * cfg->max_height;
  c->slots.resize(cfg->num_slots);
  for (Slot& s : c->slots) {
    memset(s.ev, 0, sizeof(s.ev));
    memset(s.ev_used, 0, sizeof(s.ev_used));
    CKC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&s.last_ready, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&s.match_done, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&s.front_done, cudaEventDisableTiming));
    for (int i = 0; i < VPL_NUM_STAGES; ++i)
      for (int j = 0; j < 2; ++j) CKC(cudaEventCreate(&s.ev[i][j]));
    CKC(dmalloc(&s.d_img, B * P0));
    CKC(dmalloc(&s.d_pre, B * P0));
    CKC(dmalloc(&s.d_lut, B * 256 * 256));
    for (int o = 0; o < cfg->max_octaves; ++o) {
      // worst case over aspect ratios with the same area: size by area with slack
      size_t Po = (P0 >> (2 * o)) + 64;
      size_t So = (size_t)((double)Po * 0.64) + 2 * (size_t)(cfg->max_width + cfg->max_height) + 64;
      OctBuf& b = s.oct[o];
      CKC(dmalloc(&b.pyr, B * Po));
      CKC(dmalloc(&b.grad, B * Po));
      if (!cfg->lsd_path) continue;  // EDLines / KLT front end only: pyramid image + Sobel pair of octave 0 suffice
      CKC(dmalloc(&b.scl, B * So));
      CKC(dmalloc(&b.ang, B * So));
      CKC(dmalloc(&b.pix, B * So));
      CKC(dmalloc(&b.ord, B * So));
      CKC(dmalloc(&b.reg, B * So));
      CKC(dmalloc(&b.n_ord, B));
      CKC(dmalloc(&b.cand, B * c->cand_cap));
      CKC(dmalloc(&b.n_cand, B));
      CKC(dmalloc(&b.maxq, B));
    }
    CKC(dmalloc(&s.d_counts, B));
    CKC(dmalloc(&s.d_flags, 4));
    CKC(hmalloc(&s.h_img, B * P0));
    CKC(hmalloc(&s.h_counts, B));
    CKC(hmalloc(&s.h_flags, 4));
    if (!cfg->lsd_path) continue;
    CKC(dmalloc(&s.d_kl, B * cap));
    CKC(dmalloc(&s.d_desc, B * cap * 32));
    CKC(dmalloc(&s.d_match, B * cap * c->max_k));
    CKC(dmalloc(&s.d_last_desc, cap * 32));
    CKC(dmalloc(&s.d_last_count, 1));
    CKC(dmalloc(&s.d_offsets, B + 1));
    CKC(dmalloc(&s.d_kl_dense, B * cap));
    CKC(dmalloc(&s.d_desc_dense, B * cap * 32));
    CKC(dmalloc(&s.d_match_dense, B * cap * c->max_k));
    CKC(hmalloc(&s.h_offsets, B + 1));
    {
      Slot::FrontRes& r0 = s.res[0];
      r0.done = s.done; r0.d_offsets = s.d_offsets; r0.d_kl_dense = s.d_kl_dense; r0.d_desc_dense = s.d_desc_dense;
      r0.d_match_dense = s.d_match_dense; r0.h_counts = s.h_counts; r0.h_offsets = s.h_offsets; r0.h_flags = s.h_flags;
    }
    CKC(dmalloc(&s.d_seg, (size_t)c->cand_cap));
    CKC(dmalloc(&s.d_seg_count, 1));
    CKC(hmalloc(&s.h_kl, B * cap));
    CKC(hmalloc(&s.h_desc, B * cap * 32));
    CKC(hmalloc(&s.h_match, B * cap * c->max_k));
  }
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = c;
  return VPL_OK;
}

// ---- fused path ---------------------------------------------------------------
// the second result generation of a slot (see Slot::FrontRes), allocated when a batch is first queued behind another
static int ensure_second_generation(VplContext* c, Slot& s) {
  Slot::FrontRes& r = s.res[1];
  if (r.done) return VPL_OK;
  const size_t B = (size_t)c->cfg.max_batch, cap = (size_t)c->cfg.max_lines;
  CK(c, dmalloc(&r.d_offsets, B + 1));
  CK(c, dmalloc(&r.d_kl_dense, B * cap));
  CK(c, dmalloc(&r.d_desc_dense, B * cap * 32));
  CK(c, dmalloc(&r.d_match_dense, B * cap * c->max_k));
  CK(c, hmalloc(&r.h_counts, B));
  CK(c, hmalloc(&r.h_offsets, B + 1));
  CK(c, hmalloc(&r.h_flags, 4));
  CK(c, cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming));
  return VPL_OK;
}

// vpl_frontend_submit up to the kernels: argument checks, then the frames (uploaded now, or taken from the buffer that
// vpl_frontend_upload filled) and the optional pre-processing
static int submit_prepare(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                          int scale, int num_octaves, int k) {
  int r = check_dims(c, n, w, h, num_octaves, scale, true);
  if (r) return r;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (k < 0 || k > c->max_k) return fail(c, VPL_E_INVALID, "k=%d outside 0..%d", k, c->max_k);
  if (n == 0) return fail(c, VPL_E_INVALID, "empty batch");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (s.in_flight) {
    // one more front-end batch may queue behind an uncollected one if its frames were uploaded ahead (the slot's
    // staging buffer may still feed the batch in flight) -- it writes the slot's other result generation
    if (!(s.kind == BK_FRONTEND && !imgs && s.gen_sub - s.gen_col < 2))
      return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
    int r1 = ensure_second_generation(c, s);
    if (r1) return r1;
  }
  if (!imgs) {
    // the frames were staged by vpl_frontend_upload: make its buffer the slot's input
    if (s.staged_n == 0) return fail(c, VPL_E_INVALID, "imgs == NULL but slot %d holds no uploaded batch (vpl_frontend_upload)", slot);
    if (s.staged_n != n || s.staged_w != w || s.staged_h != h)
      return fail(c, VPL_E_INVALID, "slot %d holds an uploaded batch of %d frames %dx%d, submit asks for %d frames %dx%d",
                  slot, s.staged_n, s.staged_w, s.staged_h, n, w, h);
  }
  s.n = n; s.w = w; s.h = h; s.num_octaves = num_octaves; s.scale = scale; s.k = k;
  if (imgs) return upload(c, s, imgs, n, w, h, stride);
  std::swap(s.d_img, s.d_img_next);
  s.staged_n = 0;
  CK(c, cudaStreamWaitEvent(s.stream, s.uploaded, 0));
  return after_upload(c, s, n, w, h);
}

static int submit_finish(VplContext* c, Slot& s) {
  // generation 0 unless the slot's previous batch is still uncollected (then the one it did not write)
  const int gi = (s.in_flight && s.kind == BK_FRONTEND) ? 1 - s.pending_gen[s.gen_col & 1] : 0;
  Slot::FrontRes& r = s.res[gi];
  enqueue_dense_download(c, s, r);
  CK(c, cudaEventRecord(r.done, s.stream));
  s.pending_gen[s.gen_sub & 1] = gi;
  ++s.gen_sub;
  s.in_flight = true;
  s.kind = BK_FRONTEND;
  s.resident = BK_FRONTEND;
  CK(c, cudaGetLastError());
  return VPL_OK;
}

int vpl_frontend_submit(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                        int scale, int num_octaves, int k, int chain) {
  int r = submit_prepare(c, slot, imgs, n, w, h, stride, scale, num_octaves, k);
  if (r) return r;
  r = enqueue_frontend(c, slot, k, chain);
  if (r) return r;
  return submit_finish(c, c->slots[slot]);
}

int vpl_frontend_submit_group(VplContext* c, int n_slots, const int* slots, const int* n, int w, int h, int scale,
                              int num_octaves, int k, const int* chain) {
  if (!c) return VPL_E_INVALID;
  if (n_slots < 1 || n_slots > (int)c->slots.size() || !slots || !n || !chain)
    return fail(c, VPL_E_INVALID, "bad group of %d slots", n_slots);
  for (int i = 0; i < n_slots; ++i)
    for (int j = 0; j < i; ++j)
      if (slots[i] == slots[j]) return fail(c, VPL_E_INVALID, "slot %d twice in the group", slots[i]);
  // check everything before touching anything: a group is submitted whole or not at all
  for (int i = 0; i < n_slots; ++i) {
    if (slots[i] < 0 || slots[i] >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slots[i]);
    const Slot& s = c->slots[slots[i]];
    if (s.in_flight && !(s.kind == BK_FRONTEND && s.gen_sub - s.gen_col < 2))
      return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slots[i]);
    if (s.staged_n == 0 || s.staged_n != n[i] || s.staged_w != w || s.staged_h != h)
      return fail(c, VPL_E_INVALID, "slot %d holds no uploaded batch of %d frames %dx%d (vpl_frontend_upload)", slots[i], n[i], w, h);
  }
  for (int i = 0; i < n_slots; ++i) {
    int r = submit_prepare(c, slots[i], nullptr, n[i], w, h, (size_t)w, scale, num_octaves, k);
    if (r) return r;
  }
  int r = enqueue_frontend_group(c, slots, n_slots, k, chain);
  if (r) return r;
  for (int i = 0; i < n_slots; ++i) {
    r = submit_finish(c, c->slots[slots[i]]);
    if (r) return r;
  }
  return VPL_OK;
}

int vpl_frontend_upload(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (n <= 0 || n > c->cfg.max_batch || w <= 0 || h <= 0 || w > c->cfg.max_width || h > c->cfg.max_height)
    return fail(c, VPL_E_INVALID, "upload of %d frames %dx%d outside the context's limits (%d frames %dx%d)", n, w, h,
                c->cfg.max_batch, c->cfg.max_width, c->cfg.max_height);
  if (stride != (size_t)w) return fail(c, VPL_E_INVALID, "vpl_frontend_upload takes dense frames (stride == width)");
  for (int f = 0; f < n; ++f)
    if (!imgs || !imgs[f]) return fail(c, VPL_E_INVALID, "null image pointer at frame %d", f);
  for (int f = 1; f < n; ++f)
    if (imgs[f] != imgs[0] + (size_t)f * w * h)
      return fail(c, VPL_E_INVALID, "vpl_frontend_upload takes frames that are contiguous in memory (frame %d is not)", f);
  // the copy runs while the slot's staging buffer may still feed the batch in flight: it has to come straight from
  // the caller's memory, which therefore has to be pinned
  if (!in_registered_range(c, imgs[0], (size_t)n * w * h))
    return fail(c, VPL_E_INVALID, "vpl_frontend_upload takes frames in memory pinned through vpl_host_register");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (!c->up_stream) CK(c, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
  if (!s.d_img_next) {
    CK(c, cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
    CK(c, dmalloc(&s.d_img_next, (size_t)c->cfg.max_batch * c->cfg.max_width * c->cfg.max_height));
  }
  if (s.staged_n) return fail(c, VPL_E_INVALID, "slot %d already holds an uploaded batch: submit it first", slot);
  // d_img_next was the input of the batch before the latest one submitted on this slot: collected (hence finished) as
  // long as at most one batch is uncollected
  if (s.kind == BK_FRONTEND && s.gen_sub - s.gen_col >= 2)
    return fail(c, VPL_E_INVALID, "slot %d has two uncollected batches: collect the older one before uploading the next", slot);
  CK(c, cudaMemcpyAsync(s.d_img_next, imgs[0], (size_t)n * w * h, cudaMemcpyHostToDevice, c->up_stream));
  CK(c, cudaEventRecord(s.uploaded, c->up_stream));
  s.staged_n = n; s.staged_w = w; s.staged_h = h;
  return VPL_OK;
}

int vpl_frontend_collect(VplContext* c, int slot, VplKeyLine* keylines, int32_t* counts, int cap, uint8_t* desc,
                         VplDMatch* matches) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_FRONTEND) return fail(c, VPL_E_INVALID, "slot %d has no front-end batch in flight", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  Slot::FrontRes* g = nullptr;
  int r = finish_front(c, s, g);
  if (r) return r;
  const int n = g->n, k = g->k;
  if (g->h_flags[0]) return fail(c, VPL_E_CAPACITY, "a frame produced more than %d LSD candidates", c->cand_cap);
  if (g->h_flags[1]) return fail(c, VPL_E_CAPACITY, "a frame produced more than max_lines=%d keylines", c->cfg.max_lines);
  if (counts) {
    for (int f = 0; f < n; ++f) {
      if (g->h_counts[f] > cap) return fail(c, VPL_E_CAPACITY, "frame %d has %d keylines > cap %d", f, g->h_counts[f], cap);
      counts[f] = g->h_counts[f];
    }
  }
  // dense rows: now that the total is known, fetch exactly that much and scatter it into the
  // caller's frame-major layout
  const size_t total = (size_t)g->h_offsets[n];
  if (total > 0) {
    cudaStream_t ds = c->down_stream;
    if (keylines) CK(c, cudaMemcpyAsync(s.h_kl, g->d_kl_dense, total * sizeof(VplKeyLine), cudaMemcpyDeviceToHost, ds));
    if (desc) CK(c, cudaMemcpyAsync(s.h_desc, g->d_desc_dense, total * 32, cudaMemcpyDeviceToHost, ds));
    if (matches && k > 0)
      CK(c, cudaMemcpyAsync(s.h_match, g->d_match_dense, total * k * sizeof(VplDMatch), cudaMemcpyDeviceToHost, ds));
    CK(c, cudaStreamSynchronize(ds));
  }
  for (int f = 0; f < n; ++f) {
    const size_t off = (size_t)g->h_offsets[f], cnt = (size_t)g->h_counts[f];
    if (keylines) memcpy(keylines + (size_t)f * cap, s.h_kl + off, cnt * sizeof(VplKeyLine));
    if (desc) memcpy(desc + (size_t)f * cap * 32, s.h_desc + off * 32, cnt * 32);
    if (matches && k > 0) memcpy(matches + (size_t)f * cap * k, s.h_match + off * k, cnt * k * sizeof(VplDMatch));
  }
  s.last_d2h_bytes = (int64_t)(total * (sizeof(VplKeyLine) + 32 + (size_t)k * sizeof(VplDMatch)) + (2 * (size_t)n + 3) * sizeof(int));
  return VPL_OK;
}

int vpl_frontend_collect_dense(VplContext* c, int slot, int32_t* counts, VplKeyLine* keylines, uint8_t* desc,
                               VplDMatch* matches, int64_t cap_total, int64_t* total_out) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_FRONTEND) return fail(c, VPL_E_INVALID, "slot %d has no front-end batch in flight", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  Slot::FrontRes* g = nullptr;
  int r = finish_front(c, s, g);
  if (r) return r;
  const int n = g->n, k = g->k;
  if (g->h_flags[0]) return fail(c, VPL_E_CAPACITY, "a frame produced more than %d LSD candidates", c->cand_cap);
  if (g->h_flags[1]) return fail(c, VPL_E_CAPACITY, "a frame produced more than max_lines=%d keylines", c->cfg.max_lines);
  const size_t total = (size_t)g->h_offsets[n];
  if (total_out) *total_out = (int64_t)total;
  if ((int64_t)total > cap_total) return fail(c, VPL_E_CAPACITY, "%zu keylines in the batch > cap_total %lld", total, (long long)cap_total);
  if (counts) memcpy(counts, g->h_counts, (size_t)n * sizeof(int));
  struct Out { void* user; const void* dev; void* stage; size_t bytes; };
  const Out outs[3] = {{keylines, g->d_kl_dense, s.h_kl, total * sizeof(VplKeyLine)},
                       {desc, g->d_desc_dense, s.h_desc, total * 32},
                       {(k > 0) ? matches : nullptr, g->d_match_dense, s.h_match, total * (size_t)k * sizeof(VplDMatch)}};
  bool direct[3] = {false, false, false};
  for (int i = 0; i < 3; ++i) {
    if (!outs[i].user || outs[i].bytes == 0) continue;
    direct[i] = in_registered_range(c, (const uint8_t*)outs[i].user, outs[i].bytes);
    CK(c, cudaMemcpyAsync(direct[i] ? outs[i].user : outs[i].stage, outs[i].dev, outs[i].bytes, cudaMemcpyDeviceToHost,
                          c->down_stream));
  }
  CK(c, cudaStreamSynchronize(c->down_stream));
  for (int i = 0; i < 3; ++i)
    if (outs[i].user && outs[i].bytes && !direct[i]) memcpy(outs[i].user, outs[i].stage, outs[i].bytes);
  s.last_d2h_bytes = (int64_t)(total * (sizeof(VplKeyLine) + 32 + (size_t)k * sizeof(VplDMatch)) + (2 * (size_t)n + 3) * sizeof(int));
  return VPL_OK;
}

int vpl_frontend_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride, int scale,
                       int num_octaves, int k, int chain, VplKeyLine* keylines, int32_t* counts, int cap,
                       uint8_t* desc, VplDMatch* matches) {
  if (!c) return VPL_E_INVALID;
  // alternate slots so that chaining can read the previous batch's last frame
  int slot = 0;
  if (c->slots.size() > 1 && c->have_prev) slot = (c->prev_slot + 1) % (int)c->slots.size();
  int r = vpl_frontend_submit(c, slot, imgs, n, w, h, stride, scale, num_octaves, k, chain);
  if (r) return r;
  return vpl_frontend_collect(c, slot, keylines, counts, cap, desc, matches);
}

int vpl_frontend_run_resident(VplContext* c, int slot, int k) {
  if (!c) return VPL_E_INVALID;
  if (!c->cfg.lsd_path) return fail(c, VPL_E_INVALID, "this context was created with lsd_path = 0");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (k < 0 || k > c->max_k) return fail(c, VPL_E_INVALID, "k=%d outside 0..%d", k, c->max_k);
  if (s.n <= 0 || s.resident != BK_FRONTEND)
    return fail(c, VPL_E_INVALID, "slot %d holds no front-end frames: submit+collect a front-end batch first", slot);
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight", slot);
  int r = check_dims(c, s.n, s.w, s.h, s.num_octaves, s.scale, true);
  if (r) return r;
  CK(c, cudaSetDevice(c->cfg.device));
  r = enqueue_frontend(c, slot, k, 0);
  if (r) return r;
  CK(c, cudaGetLastError());
  return VPL_OK;
}

int vpl_frontend_run_resident_group(VplContext* c, int n_slots, const int* slots, int k) {
  if (!c) return VPL_E_INVALID;
  if (!c->cfg.lsd_path) return fail(c, VPL_E_INVALID, "this context was created with lsd_path = 0");
  if (n_slots < 1 || n_slots > (int)c->slots.size() || !slots) return fail(c, VPL_E_INVALID, "bad group of %d slots", n_slots);
  if (k < 0 || k > c->max_k) return fail(c, VPL_E_INVALID, "k=%d outside 0..%d", k, c->max_k);
  int chain[64] = {0};
  if (n_slots > 64) return fail(c, VPL_E_INVALID, "group of %d slots", n_slots);
  for (int i = 0; i < n_slots; ++i) {
    if (slots[i] < 0 || slots[i] >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slots[i]);
    for (int j = 0; j < i; ++j)
      if (slots[i] == slots[j]) return fail(c, VPL_E_INVALID, "slot %d twice in the group", slots[i]);
    Slot& s = c->slots[slots[i]];
    if (s.n <= 0 || s.resident != BK_FRONTEND)
      return fail(c, VPL_E_INVALID, "slot %d holds no front-end frames: submit+collect a front-end batch first", slots[i]);
    if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight", slots[i]);
    int r = check_dims(c, s.n, s.w, s.h, s.num_octaves, s.scale, true);
    if (r) return r;
  }
  CK(c, cudaSetDevice(c->cfg.device));
  int r = enqueue_frontend_group(c, slots, n_slots, k, chain);
  if (r) return r;
  CK(c, cudaGetLastError());
  return VPL_OK;
}

int vpl_set_preprocess(VplContext* c, const float* mapx, const float* mapy, int w, int h, double clahe_clip,
                       int clahe_tiles) {
  if (!c) return VPL_E_INVALID;
  if (w <= 0 || h <= 0 || w > c->cfg.max_width || h > c->cfg.max_height)
    return fail(c, VPL_E_CAPACITY, "pre-processing size %dx%d outside the context's %dx%d", w, h, c->cfg.max_width, c->cfg.max_height);
  if ((mapx == nullptr) != (mapy == nullptr)) return fail(c, VPL_E_INVALID, "give both maps or neither");
  if (clahe_clip > 0.0 && (clahe_tiles < 1 || clahe_tiles > 16)) return fail(c, VPL_E_INVALID, "CLAHE grid must be 1..16 tiles per side");
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots) CK(c, cudaStreamSynchronize(s.stream));
  cudaFree(c->d_mapx); cudaFree(c->d_mapy);
  c->d_mapx = c->d_mapy = nullptr;
  c->pre_remap = false;
  if (mapx) {
    const size_t bytes = (size_t)w * h * sizeof(float);
    CK(c, cudaMalloc((void**)&c->d_mapx, bytes));
    CK(c, cudaMalloc((void**)&c->d_mapy, bytes));
    CK(c, cudaMemcpy(c->d_mapx, mapx, bytes, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_mapy, mapy, bytes, cudaMemcpyHostToDevice));
    if (!c->d_wtab) {
      // OpenCV's bilinear weight table (initInterTab2D): 32x32 x 4 weights summing to 32768
      std::vector<uint16_t> tab(32 * 32 * 4);
      for (int i = 0; i < 32; ++i) {
        float fy = (float)i / 32.0f;
        for (int j = 0; j < 32; ++j) {
          float fx = (float)j / 32.0f;
          float wy[2] = {1.0f - fy, fy}, wx[2] = {1.0f - fx, fx};
          int it[4], isum = 0;
          for (int k = 0; k < 4; ++k) {
            volatile float v = wy[k >> 1] * wx[k & 1];
            long r = lrintf(v * 32768.0f);
            it[k] = (int)std::min<long>(32767, std::max<long>(-32768, r));
            isum += it[k];
          }
          if (isum != 32768) {
            int diff = isum - 32768, kmax = 0, kmin = 0;
            for (int k = 1; k < 4; ++k) {
              if (it[k] > it[kmax]) kmax = k;
              if (it[k] < it[kmin]) kmin = k;
            }
            if (diff < 0) it[kmax] -= diff;
            else it[kmin] -= diff;
          }
          for (int k = 0; k < 4; ++k) tab[(size_t)(i * 32 + j) * 4 + k] = (uint16_t)it[k];
        }
      }
      CK(c, cudaMalloc((void**)&c->d_wtab, tab.size() * sizeof(uint16_t)));
      CK(c, cudaMemcpy(c->d_wtab, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    c->pre_remap = true;
  }
  c->pre_clip = clahe_clip > 0.0 ? clahe_clip : 0.0;
  c->pre_tiles = clahe_tiles;
  c->pre_w = w; c->pre_h = h;
  return VPL_OK;
}

int vpl_preprocess_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride, uint8_t* out) {
  int r = check_dims(c, n, w, h, 1, 2);
  if (r) return r;
  if (!out) return fail(c, VPL_E_INVALID, "null output");
  if (n == 0) return VPL_OK;
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  s.n = n; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 2; s.k = 0;
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  CK(c, cudaMemcpyAsync(s.h_img, s.d_img, (size_t)n * w * h, cudaMemcpyDeviceToHost, s.stream));
  r = finish(c, s);
  if (r) return r;
  memcpy(out, s.h_img, (size_t)n * w * h);
  return VPL_OK;
}

int vpl_host_register(VplContext* c, const void* ptr, size_t bytes) {
  if (!c || !ptr || bytes == 0) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable));
  c->pinned.emplace_back((const uint8_t*)ptr, bytes);
  return VPL_OK;
}

int vpl_host_unregister(VplContext* c, const void* ptr) {
  if (!c || !ptr) return VPL_E_INVALID;
  for (size_t i = 0; i < c->pinned.size(); ++i)
    if (c->pinned[i].first == (const uint8_t*)ptr) {
      // no batch may still be reading from the range
      for (Slot& s : c->slots) cudaStreamSynchronize(s.stream);
      cudaHostUnregister(const_cast<void*>(ptr));
      c->pinned.erase(c->pinned.begin() + (long)i);
      return VPL_OK;
    }
  return fail(c, VPL_E_INVALID, "pointer was not registered");
}

int64_t vpl_last_d2h_bytes(const VplContext* c, int slot) {
  if (!c || slot < 0 || slot >= (int)c->slots.size()) return 0;
  return c->slots[slot].last_d2h_bytes;
}

int vpl_sync(VplContext* c) {
  if (!c) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots) {
    CK(c, cudaStreamSynchronize(s.stream));
    harvest_times(c, s);
  }
  CK(c, cudaGetLastError());
  return VPL_OK;
}

// ---- the three OpenCV-shaped calls ----------------------------------------------
int vpl_lsd_detect_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride, int scale,
                         int num_octaves, VplKeyLine* keylines, int32_t* counts, int cap) {
  int r = check_dims(c, n, w, h, num_octaves, scale, true);
  if (r) return r;
  if (n == 0) return VPL_OK;
  if (!keylines || !counts) return fail(c, VPL_E_INVALID, "null output");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  s.n = n; s.w = w; s.h = h; s.num_octaves = num_octaves; s.scale = scale; s.k = 0;
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  run_pyramid(c, s, c->cfg.blur_first ? 1 : 0);
  run_lsd(c, s);
  run_pack(c, s);
  enqueue_download(c, s, true, false, false);
  r = finish(c, s);
  if (r) return r;
  if (s.h_flags[0]) return fail(c, VPL_E_CAPACITY, "a frame produced more than %d LSD candidates", c->cand_cap);
  if (s.h_flags[1]) return fail(c, VPL_E_CAPACITY, "a frame produced more than max_lines=%d keylines", c->cfg.max_lines);
  for (int f = 0; f < n; ++f) {
    if (s.h_counts[f] > cap) return fail(c, VPL_E_CAPACITY, "frame %d has %d keylines > cap %d", f, s.h_counts[f], cap);
    counts[f] = s.h_counts[f];
  }
  copy_rows(keylines, s.h_kl, s.h_counts, n, (size_t)c->cfg.max_lines, (size_t)cap, sizeof(VplKeyLine));
  return VPL_OK;
}

int vpl_lbd_compute_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                          const VplKeyLine* keylines, const int32_t* counts, int cap, uint8_t* desc) {
  if (!c) return VPL_E_INVALID;
  if (!keylines || !counts || !desc) return fail(c, VPL_E_INVALID, "null argument");
  int max_oct = 0;
  for (int f = 0; f < n; ++f) {
    if (counts[f] < 0 || counts[f] > cap || counts[f] > c->cfg.max_lines)
      return fail(c, VPL_E_CAPACITY, "frame %d: %d keylines (cap %d, max_lines %d)", f, counts[f], cap, c->cfg.max_lines);
    for (int i = 0; i < counts[f]; ++i) {
      const VplKeyLine& k = keylines[(size_t)f * cap + i];
      if (k.octave < 0) return fail(c, VPL_E_INVALID, "negative octave");
      if (k.octave > max_oct) max_oct = k.octave;
      if (k.numOfPixels < 0 || k.numOfPixels > 32767) return fail(c, VPL_E_INVALID, "numOfPixels out of range");
    }
  }
  int r = check_dims(c, n, w, h, max_oct + 1, 2, true);
  if (r) return r;
  if (n == 0) return VPL_OK;
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  s.n = n; s.w = w; s.h = h; s.num_octaves = max_oct + 1; s.scale = 2; s.k = 0;
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  const size_t mc = (size_t)c->cfg.max_lines;
  copy_rows(s.h_kl, keylines, counts, n, (size_t)cap, mc, sizeof(VplKeyLine));
  memcpy(s.h_counts, counts, (size_t)n * sizeof(int));
  CK(c, cudaMemcpyAsync(s.d_kl, s.h_kl, (size_t)n * mc * sizeof(VplKeyLine), cudaMemcpyHostToDevice, s.stream));
  CK(c, cudaMemcpyAsync(s.d_counts, s.h_counts, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  run_pyramid(c, s, 1);
  run_lbd(c, s, s.num_octaves);
  CK(c, cudaMemsetAsync(s.d_flags, 0, 2 * sizeof(int), s.stream));
  enqueue_download(c, s, false, true, false);
  r = finish(c, s);
  if (r) return r;
  copy_rows(desc, s.h_desc, counts, n, mc, (size_t)cap, 32);
  return VPL_OK;
}

int vpl_lbd_compute_float_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                                const VplKeyLine* keylines, const int32_t* counts, int cap, float* fdesc) {
  // BinaryDescriptor::compute(..., returnFloatDescr = true): the 72-float LBD vectors.
  if (!c) return VPL_E_INVALID;
  if (!keylines || !counts || !fdesc) return fail(c, VPL_E_INVALID, "null argument");
  const size_t mc = (size_t)c->cfg.max_lines;
  std::vector<uint8_t> bytes((size_t)std::max(n, 1) * cap * 32);
  if (!c->d_fdesc) {  // 72 floats per line: a scratch of its own, allocated at the first use of this call
    CK(c, cudaSetDevice(c->cfg.device));
    CK(c, dmalloc(&c->d_fdesc, (size_t)c->cfg.max_batch * mc * 72));
  }
  // run the byte path first (validates arguments, uploads, builds the pyramid), then re-run LBD with the float sink
  int r = vpl_lbd_compute_batch(c, imgs, n, w, h, stride, keylines, counts, cap, bytes.data());
  if (r) return r;
  if (n == 0) return VPL_OK;
  Slot& s = c->slots[0];
  float* d_f = c->d_fdesc;
  run_lbd(c, s, s.num_octaves, d_f);
  std::vector<float> tmp((size_t)n * mc * 72);
  CK(c, cudaMemcpyAsync(tmp.data(), d_f, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
  r = finish(c, s);
  if (r) return r;
  for (int f = 0; f < n; ++f)
    memcpy(fdesc + (size_t)f * cap * 72, tmp.data() + (size_t)f * mc * 72, (size_t)counts[f] * 72 * sizeof(float));
  return VPL_OK;
}

int vpl_match_batch(VplContext* c, const uint8_t* q, const int32_t* nq, int cap_q, const uint8_t* t,
                    const int32_t* nt, int cap_t, int n_pairs, int k, VplDMatch* out) {
  if (!c) return VPL_E_INVALID;
  if (!c->cfg.lsd_path) return fail(c, VPL_E_INVALID, "this context was created with lsd_path = 0");
  if (!q || !nq || !t || !nt || !out) return fail(c, VPL_E_INVALID, "null argument");
  if (k < 1 || k > c->max_k) return fail(c, VPL_E_INVALID, "k=%d outside 1..%d", k, c->max_k);
  if (n_pairs < 0 || cap_q < 1 || cap_t < 1) return fail(c, VPL_E_INVALID, "bad sizes");
  if (n_pairs == 0) return VPL_OK;
  // the train set reuses the descriptor buffer of slot 0, the query set that of the
  // last slot (or the second half of the same buffer if there is one slot)
  const size_t mc = (size_t)c->cfg.max_lines, B = (size_t)c->cfg.max_batch;
  if ((size_t)cap_q > mc * B / std::max<size_t>(1, (size_t)n_pairs) ||
      (size_t)cap_t > mc * B / std::max<size_t>(1, (size_t)n_pairs))
    return fail(c, VPL_E_CAPACITY, "n_pairs*cap exceeds max_batch*max_lines");
  if (n_pairs > c->cfg.max_batch) return fail(c, VPL_E_CAPACITY, "n_pairs %d > max_batch", n_pairs);
  if (cap_t >= (1 << 20)) return fail(c, VPL_E_CAPACITY, "train set too large");
  for (int p = 0; p < n_pairs; ++p)
    if (nq[p] < 0 || nq[p] > cap_q || nt[p] < 0 || nt[p] > cap_t) return fail(c, VPL_E_INVALID, "bad counts in pair %d", p);
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  // device scratch: q -> d_desc, t -> reuse d_kl region (B*cap*68 B >= B*cap*32 B), counts -> d_counts / n_ord
  uint8_t* d_q = s.d_desc;
  uint8_t* d_t = reinterpret_cast<uint8_t*>(s.d_kl);
  int* d_nq = s.d_counts;
  int* d_nt = s.oct[0].n_ord;
  const size_t qb = (size_t)n_pairs * cap_q * 32, tb = (size_t)n_pairs * cap_t * 32;
  memcpy(s.h_desc, q, qb);
  memcpy(s.h_kl, t, tb);
  memcpy(s.h_counts, nq, (size_t)n_pairs * sizeof(int));
  CK(c, cudaMemcpyAsync(d_q, s.h_desc, qb, cudaMemcpyHostToDevice, s.stream));
  CK(c, cudaMemcpyAsync(d_t, s.h_kl, tb, cudaMemcpyHostToDevice, s.stream));
  CK(c, cudaMemcpyAsync(d_nq, s.h_counts, (size_t)n_pairs * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  CK(c, cudaStreamSynchronize(s.stream));
  memcpy(s.h_counts, nt, (size_t)n_pairs * sizeof(int));
  CK(c, cudaMemcpyAsync(d_nt, s.h_counts, (size_t)n_pairs * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  {
    StageTimer tm(c, s, VPL_STAGE_MATCH);
    launch_hamming_knn(d_q, d_nq, cap_q, d_t, d_nt, cap_t, n_pairs, k, s.d_match, s.stream);
    tm.launches(1);
  }
  s.m_pairs = n_pairs; s.m_cap_q = cap_q; s.m_cap_t = cap_t;
  const size_t ob = (size_t)n_pairs * cap_q * k * sizeof(VplDMatch);
  CK(c, cudaMemcpyAsync(s.h_match, s.d_match, ob, cudaMemcpyDeviceToHost, s.stream));
  int r = finish(c, s);
  if (r) return r;
  for (int p = 0; p < n_pairs; ++p)
    memcpy(out + (size_t)p * cap_q * k, s.h_match + (size_t)p * cap_q * k, (size_t)nq[p] * k * sizeof(VplDMatch));
  return VPL_OK;
}

int vpl_match_run_resident(VplContext* c, int k) {
  if (!c) return VPL_E_INVALID;
  Slot& s = c->slots[0];
  if (s.m_pairs <= 0) return fail(c, VPL_E_INVALID, "no descriptor sets resident: call vpl_match_batch first");
  if (k < 1 || k > c->max_k) return fail(c, VPL_E_INVALID, "k=%d outside 1..%d", k, c->max_k);
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  CK(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  StageTimer tm(c, s, VPL_STAGE_MATCH);
  launch_hamming_knn(s.d_desc, s.d_counts, s.m_cap_q, reinterpret_cast<uint8_t*>(s.d_kl), s.oct[0].n_ord, s.m_cap_t, s.m_pairs, k,
                     s.d_match, s.stream);
  tm.launches(1);
  return VPL_OK;
}

int vpl_debug_popc_peak(VplContext* c, double* popc32_per_s) {
  if (!c || !popc32_per_s) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  cudaDeviceProp prop;
  CK(c, cudaGetDeviceProperties(&prop, c->cfg.device));
  unsigned* d = nullptr;
  CK(c, cudaMalloc((void**)&d, 256));
  cudaEvent_t e0, e1;
  CK(c, cudaEventCreate(&e0));
  CK(c, cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first pass warms up
    cudaEventRecord(e0, 0);
    launch_popc_peak(d, blocks, iters, 0);
    cudaEventRecord(e1, 0);
    CK(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  c->launches += 4;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *popc32_per_s = best;
  return VPL_OK;
}


// ---- EDLines: the reference's real detector -------------------------------------
void vpl_edlines_default_param(VplEDLineParam* p) {
  if (!p) return;
  // the tracker node's parameters with the EuRoC configuration
  // (feature_tracker/src/line_feature_tracker_node.cpp:203, config/euroc/euroc_config.yaml:84,87)
  p->ksize = 5; p->sigma = 1.0f; p->gradientThreshold = 30; p->anchorThreshold = 5; p->scanIntervals = 2;
  p->minLineLen = 35; p->lineFitErrThreshold = 1.8;
}

int vpl_edlines_configure(VplContext* c, const VplEDLineParam* p) {
  if (!c || !p) return VPL_E_INVALID;
  if (p->scanIntervals < 1 || p->minLineLen < 2 || !(p->lineFitErrThreshold >= 0) || p->gradientThreshold < 0 ||
      p->gradientThreshold > 32000 || p->anchorThreshold < 0 || p->anchorThreshold > 255)
    return fail(c, VPL_E_INVALID, "bad EDLineParam");
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots)
    if (s.in_flight) return fail(c, VPL_E_INVALID, "vpl_edlines_configure while a batch is in flight");
  CK(c, cudaDeviceSynchronize());
  c->ed_ready = false;
  c->edp = *p;
  const EdGeom G = ed_geom(c->cfg.max_width, c->cfg.max_height, *p);
  const size_t B = (size_t)c->cfg.max_batch, P0 = (size_t)c->cfg.max_width * c->cfg.max_height;
  // buffer sizes must dominate every smaller image: all EdGeom sizes are monotone in w and h
  // except the scan grid of a shape with the same area but another aspect; size it by area
  const size_t bm_words = (P0 / ((size_t)p->scanIntervals * p->scanIntervals) + 2 * (c->cfg.max_width + c->cfg.max_height)) / 32 + 8;
  for (Slot& s : c->slots) {
    ed_free(s);
    CK(c, dmalloc(&s.ed.gmap, B * P0));
    CK(c, dmalloc(&s.ed.bitmap, B * bm_words));
    CK(c, dmalloc(&s.ed.n_anchor, B));
    CK(c, dmalloc(&s.ed.first, B * (size_t)G.part_cap));
    CK(c, dmalloc(&s.ed.second, B * (size_t)G.part_cap));
    CK(c, dmalloc(&s.ed.xy, B * 2 * (size_t)G.cap_px));
    CK(c, dmalloc(&s.ed.sid, B * ((size_t)G.cap_edges + 2)));
    CK(c, dmalloc(&s.ed.n_chain, B));
    CK(c, dmalloc(&s.ed.n_px, B));
    CK(c, dmalloc(&s.ed.status, B));
    CK(c, dmalloc(&s.ed.slots, B * (size_t)G.nslots));
    CK(c, dmalloc(&s.ed.slot_valid, B * (size_t)G.nslots));
    CK(c, dmalloc(&s.d_lines, B * (size_t)c->cfg.max_lines));
    CK(c, hmalloc(&s.h_lines, B * (size_t)c->cfg.max_lines));
    CK(c, hmalloc(&s.h_ed_status, B));
  }
  c->ed_ready = true;
  return VPL_OK;
}

int vpl_edlines_submit(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                       int smoothed) {
  int r = ed_check(c, n, w, h, smoothed);
  if (r) return r;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (n == 0) return fail(c, VPL_E_INVALID, "empty batch");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  s.n = n; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 1; s.k = 0; s.ed_smoothed = smoothed ? 1 : 0;
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  run_edlines(c, s);
  ed_enqueue_download(c, s);
  CK(c, cudaEventRecord(s.done, s.stream));
  s.in_flight = true;
  s.kind = BK_EDLINES;
  s.resident = BK_EDLINES;
  return VPL_OK;
}

int vpl_edlines_collect(VplContext* c, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* status) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_EDLINES) return fail(c, VPL_E_INVALID, "slot %d has no EDLines batch in flight", slot);
  if (!lines || !counts) return fail(c, VPL_E_INVALID, "null output");
  CK(c, cudaSetDevice(c->cfg.device));
  int r = finish(c, s);
  if (r) return r;
  return ed_deliver(c, s, lines, counts, cap, status);
}

int vpl_edlines_detect_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                             int smoothed, VplLine* lines, int32_t* counts, int cap, int32_t* status) {
  int r = ed_check(c, n, w, h, smoothed);
  if (r) return r;
  if (n == 0) return VPL_OK;
  r = vpl_edlines_submit(c, 0, imgs, n, w, h, stride, smoothed);
  if (r) return r;
  return vpl_edlines_collect(c, 0, lines, counts, cap, status);
}

int vpl_edlines_run_resident(VplContext* c, int slot) {
  if (!c) return VPL_E_INVALID;
  if (!c->ed_ready) return fail(c, VPL_E_INVALID, "call vpl_edlines_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (s.n <= 0) return fail(c, VPL_E_INVALID, "slot %d holds no frames", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  run_edlines(c, s);
  return VPL_OK;
}

int vpl_debug_edge_chains(VplContext* c, int frame, uint32_t* xy, int cap_px, uint32_t* sid, int cap_chains,
                          int32_t* n_px, int32_t* n_chains) {
  if (!c || !n_px || !n_chains) return VPL_E_INVALID;
  if (!c->ed_ready) return fail(c, VPL_E_INVALID, "call vpl_edlines_configure first");
  Slot& s = c->slots[0];
  if (frame < 0 || frame >= s.n) return fail(c, VPL_E_INVALID, "frame %d outside the last batch", frame);
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaStreamSynchronize(s.stream));
  const EdGeom G = ed_geom(s.w, s.h, c->edp);
  int np = 0, nc = 0;
  CK(c, cudaMemcpy(&np, s.ed.n_px + frame, sizeof(int), cudaMemcpyDeviceToHost));
  CK(c, cudaMemcpy(&nc, s.ed.n_chain + frame, sizeof(int), cudaMemcpyDeviceToHost));
  *n_px = np; *n_chains = nc;
  if (np > cap_px || nc > cap_chains) return fail(c, VPL_E_CAPACITY, "edge chains: %d px / %d chains exceed the buffers", np, nc);
  if (xy && np) CK(c, cudaMemcpy(xy, s.ed.xy + (size_t)frame * 2 * G.cap_px, (size_t)np * 4, cudaMemcpyDeviceToHost));
  if (sid) {
    if (nc) CK(c, cudaMemcpy(sid, s.ed.sid + (size_t)frame * (G.cap_edges + 2), ((size_t)nc + 1) * 4, cudaMemcpyDeviceToHost));
    else sid[0] = 0;
  }
  return VPL_OK;
}


// ---- LineMatching::Matching: the reference's real matcher ------------------------
void vpl_linematch_default_param(VplLineMatchParam* p) {
  if (!p) return;
  p->step = 10; p->closest_line_threshold = 0.5f; p->line_matching_ratio = 0.4f; p->line_distance_error_ratio = 3.f;
  p->klt_error_threshold = 40.f;                                  // line_matching.h:14-18
  p->max_level = 3; p->max_count = 30; p->epsilon = 0.001; p->min_eig = 1e-4f;  // line_matching.cpp:14
  p->topo_distance_threshold = 15.f; p->topo_length_ratio = 0.2f; p->topo_violation_ratio = 0.05f;  // line_matching.h:45-47
  p->illumination_adapt = 1; p->topological_filter = 1;          // line_feature_tracker.cpp:307-308
  p->max_anchors = 4096;
}

int vpl_linematch_configure(VplContext* c, const VplLineMatchParam* p) {
  if (!c || !p) return VPL_E_INVALID;
  if (p->step < 1 || p->max_level < 0 || p->max_level > kKltMaxLevels - 1 || p->max_anchors < 1 || p->max_count < 0)
    return fail(c, VPL_E_INVALID, "bad VplLineMatchParam");
  if (c->cfg.max_lines > 4096) return fail(c, VPL_E_INVALID, "line matching supports max_lines <= 4096");
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots)
    if (s.in_flight) return fail(c, VPL_E_INVALID, "vpl_linematch_configure while a batch is in flight");
  CK(c, cudaDeviceSynchronize());
  c->lm_ready = false;
  c->lmp = *p;
  // pyramid size by area with slack: any w x h inside the context's limits must fit
  const KltGeom G = klt_geom(c->cfg.max_width, c->cfg.max_height, p->max_level);
  const size_t frame = G.img_frame + 4096;
  const size_t B = (size_t)c->cfg.max_batch, cap = (size_t)c->cfg.max_lines, ck = (size_t)p->max_anchors;
  for (Slot& s : c->slots) {
    lm_free(s);
    CK(c, dmalloc(&s.lm_pyr, B * frame));
    CK(c, dmalloc(&s.lm_deriv, B * frame));
    s.lm.cap_kp = p->max_anchors;
    CK(c, dmalloc(&s.lm.kps, B * ck));
    CK(c, dmalloc(&s.lm.nxt, B * ck));
    CK(c, dmalloc(&s.lm.status, B * ck));
    CK(c, dmalloc(&s.lm.err, B * ck));
    CK(c, dmalloc(&s.lm.kp2line, B * ck));
    CK(c, dmalloc(&s.lm.kp_start, B * (cap + 1)));
    CK(c, dmalloc(&s.lm.n_kp, B));
    CK(c, dmalloc(&s.lm.r2c, B * cap));
    CK(c, dmalloc(&s.lm.matched, B));
    CK(c, dmalloc(&s.lm.pair_off, B + 1));
    CK(c, dmalloc(&s.lm.queue, (size_t)kKltMaxLevels));
    CK(c, dmalloc(&s.d_lm_counts, B));
    CK(c, hmalloc(&s.h_r2c, B * cap));
    if (!s.d_lines) {  // line storage is shared with the EDLines detector
      CK(c, dmalloc(&s.d_lines, B * cap));
      CK(c, hmalloc(&s.h_lines, B * cap));
      CK(c, hmalloc(&s.h_ed_status, B));
    }
  }
  c->lm_ready = true;
  return VPL_OK;
}

int vpl_linematch_batch(VplContext* c, const uint8_t* const* imgs_ref, const uint8_t* const* imgs_cur, int n_pairs,
                        int w, int h, size_t stride, const VplLine* lines_ref, const int32_t* n_ref,
                        const VplLine* lines_cur, const int32_t* n_cur, int cap, int32_t* ref_to_cur) {
  int r = lm_check(c);
  if (r) return r;
  r = check_dims(c, 2 * n_pairs, w, h, 1, 1);
  if (r) return r;
  if (n_pairs == 0) return VPL_OK;
  if (!imgs_ref || !imgs_cur || !lines_ref || !lines_cur || !n_ref || !n_cur || !ref_to_cur)
    return fail(c, VPL_E_INVALID, "null argument");
  const int mcap = c->cfg.max_lines;
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  for (int p = 0; p < n_pairs; ++p)
    if (n_ref[p] < 0 || n_cur[p] < 0 || n_ref[p] > cap || n_cur[p] > cap || n_ref[p] > mcap || n_cur[p] > mcap)
      return fail(c, VPL_E_CAPACITY, "pair %d: %d / %d lines exceed cap %d or max_lines %d", p, n_ref[p], n_cur[p], cap, mcap);
  // frames interleaved (ref, cur) per pair; lines staged the same way
  std::vector<const uint8_t*> ptrs((size_t)2 * n_pairs);
  for (int p = 0; p < n_pairs; ++p) { ptrs[2 * p] = imgs_ref[p]; ptrs[2 * p + 1] = imgs_cur[p]; }
  s.n = 2 * n_pairs; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 1; s.k = 0;
  r = upload(c, s, ptrs.data(), 2 * n_pairs, w, h, stride);
  if (r) return r;
  for (int p = 0; p < n_pairs; ++p) {
    memcpy(s.h_lines + (size_t)(2 * p) * mcap, lines_ref + (size_t)p * cap, (size_t)n_ref[p] * sizeof(VplLine));
    memcpy(s.h_lines + (size_t)(2 * p + 1) * mcap, lines_cur + (size_t)p * cap, (size_t)n_cur[p] * sizeof(VplLine));
    s.h_counts[2 * p] = n_ref[p]; s.h_counts[2 * p + 1] = n_cur[p];
  }
  CK(c, cudaMemcpyAsync(s.d_lines, s.h_lines, (size_t)2 * n_pairs * mcap * sizeof(VplLine), cudaMemcpyHostToDevice, s.stream));
  CK(c, cudaMemcpyAsync(s.d_lm_counts, s.h_counts, (size_t)2 * n_pairs * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  run_linematch(c, s, s.d_lm_counts, 2 * n_pairs, n_pairs, 2);
  CK(c, cudaMemcpyAsync(s.h_r2c, s.lm.r2c, (size_t)n_pairs * mcap * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(c, cudaMemcpyAsync(s.h_flags, s.d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  r = finish(c, s);
  if (r) return r;
  if (s.h_flags[3]) return fail(c, VPL_E_CAPACITY, "a pair needs more than max_anchors=%d anchor points", c->lmp.max_anchors);
  for (int p = 0; p < n_pairs; ++p) {
    for (int i = 0; i < n_ref[p]; ++i) ref_to_cur[(size_t)p * cap + i] = s.h_r2c[(size_t)p * mcap + i];
  }
  return VPL_OK;
}

int vpl_debug_linematch_points(VplContext* c, int pair, float* kps_ref, float* kps_cur, uint8_t* status, float* err,
                               int32_t* kp2line_cur, int cap, int32_t* n) {
  int r = lm_check(c);
  if (r) return r;
  if (!n) return VPL_E_INVALID;
  Slot& s = c->slots[0];
  if (pair < 0 || pair >= s.lm_pairs) return fail(c, VPL_E_INVALID, "pair %d outside the last match", pair);
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaStreamSynchronize(s.stream));
  int nk = 0;
  CK(c, cudaMemcpy(&nk, s.lm.n_kp + pair, sizeof(int), cudaMemcpyDeviceToHost));
  *n = nk;
  if (nk > cap) return fail(c, VPL_E_CAPACITY, "%d anchors > cap %d", nk, cap);
  const size_t o = (size_t)pair * s.lm.cap_kp;
  if (kps_ref) CK(c, cudaMemcpy(kps_ref, s.lm.kps + o, (size_t)nk * 8, cudaMemcpyDeviceToHost));
  if (kps_cur) CK(c, cudaMemcpy(kps_cur, s.lm.nxt + o, (size_t)nk * 8, cudaMemcpyDeviceToHost));
  if (status) CK(c, cudaMemcpy(status, s.lm.status + o, (size_t)nk, cudaMemcpyDeviceToHost));
  if (err) CK(c, cudaMemcpy(err, s.lm.err + o, (size_t)nk * 4, cudaMemcpyDeviceToHost));
  if (kp2line_cur) CK(c, cudaMemcpy(kp2line_cur, s.lm.kp2line + o, (size_t)nk * 4, cudaMemcpyDeviceToHost));
  return VPL_OK;
}

// ---- vanishing points -----------------------------------------------------------------------------
int vpl_vp_configure(VplContext* c, float f, float cx, float cy) {
  if (!c) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots)
    if (s.in_flight) return fail(c, VPL_E_INVALID, "vpl_vp_configure while a batch is in flight");
  CK(c, cudaDeviceSynchronize());
  c->vp_ready = false;
  c->vpp.f = f; c->vpp.ppx = cx; c->vpp.ppy = cy;
  {  // getVPHypVia2Lines' iteration count, as the reference computes it (vanishing_point_detection.cpp:93-97)
    double noiseRatio = 0.5;
    double p = 1.0 / 3.0 * pow(1.0 - noiseRatio, 2);
    double confEfficience = 0.9999;
    c->vpp.it = (int)(log(1 - confEfficience) / log(1.0 - p));
  }
  c->vpp.max_draws = 1000000;
  const size_t B = (size_t)c->cfg.max_batch, cap = (size_t)c->cfg.max_lines, it = (size_t)c->vpp.it;
  if (!c->d_vp_lambda) CK(c, dmalloc(&c->d_vp_lambda, 720));
  launch_vp_lambda(c->d_vp_lambda, 0);
  c->launches += 1;
  CK(c, cudaDeviceSynchronize());
  for (Slot& s : c->slots) {
    vp_free(s);
    CK(c, dmalloc(&s.vp.para, B * cap * 3));
    CK(c, dmalloc(&s.vp.length, B * cap));
    CK(c, dmalloc(&s.vp.orient, B * cap));
    CK(c, dmalloc(&s.vp.vp1, B * it * 3));
    CK(c, dmalloc(&s.vp.pairs, B * it * 2));
    CK(c, dmalloc(&s.vp.rng, B * 33));
    CK(c, dmalloc(&s.vp.status, 2 * B));
    s.vp.flags = s.vp.status + B;
    CK(c, dmalloc(&s.vp.grid, B * kVpCells));
    CK(c, dmalloc(&s.vp.grid_new, B * kVpCells));
    s.vp.lambda_sc = c->d_vp_lambda;
    CK(c, dmalloc(&s.vp.part_best, B * kVpMaxSplits));
    CK(c, dmalloc(&s.vp.part_idx, B * kVpMaxSplits));
    CK(c, dmalloc(&s.vp.best_idx, B));
    CK(c, dmalloc(&s.vp.lx, B * cap));
    CK(c, dmalloc(&s.d_vp_lines, B * cap));
    CK(c, dmalloc(&s.d_vp_all, B * cap));
    CK(c, dmalloc(&s.d_vp_n, 2 * B));
    CK(c, cudaMemset(s.d_vp_n, 0, (size_t)2 * B * sizeof(int)));
    CK(c, dmalloc(&s.d_vp_seeds, B));
    CK(c, dmalloc(&s.d_vps, B * 9));
    CK(c, dmalloc(&s.d_vp_idx, B * cap));
    CK(c, dmalloc(&s.d_line_vps, B * cap * 4));
    CK(c, hmalloc(&s.h_vp_lines, B * cap));
    CK(c, hmalloc(&s.h_vp_all, B * cap));
    CK(c, hmalloc(&s.h_vp_n, 2 * B));
    CK(c, hmalloc(&s.h_vp_seeds, B));
    CK(c, hmalloc(&s.h_vps, B * 9));
    CK(c, hmalloc(&s.h_vp_idx, B * cap));
    CK(c, hmalloc(&s.h_line_vps, B * cap * 4));
    CK(c, hmalloc(&s.h_vp_status, 2 * B));
    CK(c, dmalloc(&s.d_vp_ids, B * cap));
    CK(c, dmalloc(&s.d_cloud, B * cap * 10));
    CK(c, hmalloc(&s.h_vp_ids, B * cap));
    CK(c, hmalloc(&s.h_cloud, B * cap * 10));
    CK(c, cudaMemset(s.vp.status, 0, 2 * B * sizeof(int)));
  }
  c->vp_ready = true;
  return VPL_OK;
}

int vpl_vp_submit(VplContext* c, int slot, const VplLine* lines, const int32_t* n_lines, const VplLine* all_lines,
                  const int32_t* n_all, int n_frames, int cap, const uint32_t* seeds, int frame_count0) {
  if (!c) return VPL_E_INVALID;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (n_frames <= 0 || n_frames > c->cfg.max_batch) return fail(c, VPL_E_CAPACITY, "n_frames %d outside 1..max_batch %d", n_frames, c->cfg.max_batch);
  if (!lines || !n_lines || !seeds || cap < 0 || (all_lines && !n_all)) return fail(c, VPL_E_INVALID, "null argument");
  const int mcap = c->cfg.max_lines, B = c->cfg.max_batch;
  for (int i = 0; i < n_frames; ++i) {
    const int na = all_lines ? n_all[i] : n_lines[i];
    if (n_lines[i] < 0 || na < 0 || n_lines[i] > cap || na > cap || n_lines[i] > mcap || na > mcap)
      return fail(c, VPL_E_CAPACITY, "frame %d: %d / %d lines exceed cap %d or max_lines %d", i, n_lines[i], na, cap, mcap);
  }
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  s.vp_n = n_frames; s.vp_fc0 = frame_count0; s.vp_same = all_lines == nullptr;
  {
    StageTimer t(c, s, VPL_STAGE_H2D);
    for (int i = 0; i < n_frames; ++i) {
      memcpy(s.h_vp_lines + (size_t)i * mcap, lines + (size_t)i * cap, (size_t)n_lines[i] * sizeof(VplLine));
      s.h_vp_n[i] = n_lines[i];
      s.h_vp_seeds[i] = seeds[i];
      if (all_lines) {
        memcpy(s.h_vp_all + (size_t)i * mcap, all_lines + (size_t)i * cap, (size_t)n_all[i] * sizeof(VplLine));
        s.h_vp_n[B + i] = n_all[i];
      }
    }
    // only the rows in use travel: one copy per frame would be launch-bound, so whole rows up to the longest count
    CK(c, cudaMemcpyAsync(s.d_vp_lines, s.h_vp_lines, (size_t)n_frames * mcap * sizeof(VplLine), cudaMemcpyHostToDevice, s.stream));
    if (all_lines)
      CK(c, cudaMemcpyAsync(s.d_vp_all, s.h_vp_all, (size_t)n_frames * mcap * sizeof(VplLine), cudaMemcpyHostToDevice, s.stream));
    CK(c, cudaMemcpyAsync(s.d_vp_n, s.h_vp_n, (size_t)2 * B * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CK(c, cudaMemcpyAsync(s.d_vp_seeds, s.h_vp_seeds, (size_t)n_frames * sizeof(unsigned), cudaMemcpyHostToDevice, s.stream));
  }
  run_vp(c, s);
  {
    StageTimer t(c, s, VPL_STAGE_D2H);
    CK(c, cudaMemcpyAsync(s.h_vps, s.d_vps, (size_t)n_frames * 9 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CK(c, cudaMemcpyAsync(s.h_vp_idx, s.d_vp_idx, (size_t)n_frames * mcap * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(c, cudaMemcpyAsync(s.h_vp_status, s.vp.status, (size_t)2 * B * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  }
  CK(c, cudaEventRecord(s.done, s.stream));
  s.in_flight = true;
  s.kind = BK_VP;
  s.resident = BK_VP;
  s.vp_src_lines = s.vp_same ? s.d_vp_lines : s.d_vp_all;
  s.vp_src_n = s.vp_same ? s.d_vp_n : s.d_vp_n + B;
  s.vp_src_hn = s.vp_same ? s.h_vp_n : s.h_vp_n + B;
  return VPL_OK;
}

int vpl_vp_collect(VplContext* c, int slot, int cap, double* vps, int32_t* vp_idx, double* line_vps, int32_t* status) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_VP) return fail(c, VPL_E_INVALID, "slot %d has no vanishing-point batch in flight", slot);
  if (!vps || !vp_idx) return fail(c, VPL_E_INVALID, "null output");
  CK(c, cudaSetDevice(c->cfg.device));
  const int mcap = c->cfg.max_lines, B = c->cfg.max_batch, n = s.vp_n;
  if (line_vps)  // the per-line Vector4d copies are optional and 32 bytes per line: fetched only on request
    CK(c, cudaMemcpyAsync(s.h_line_vps, s.d_line_vps, (size_t)n * mcap * 4 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  int r = finish(c, s);
  if (r) return r;
  memcpy(vps, s.h_vps, (size_t)n * 9 * sizeof(double));
  const int* na = s.vp_same ? s.h_vp_n : s.h_vp_n + B;
  for (int i = 0; i < n; ++i) {
    if (na[i] > cap) return fail(c, VPL_E_CAPACITY, "frame %d: %d lines > cap %d", i, na[i], cap);
    memcpy(vp_idx + (size_t)i * cap, s.h_vp_idx + (size_t)i * mcap, (size_t)na[i] * sizeof(int));
    if (line_vps) memcpy(line_vps + (size_t)i * cap * 4, s.h_line_vps + (size_t)i * mcap * 4, (size_t)na[i] * 4 * sizeof(double));
    if (status) status[i] = s.h_vp_status[i] != 0 ? s.h_vp_status[i] : (s.h_vp_status[B + i] & 1);
  }
  return VPL_OK;
}

int vpl_vp_detect_batch(VplContext* c, const VplLine* lines, const int32_t* n_lines, const VplLine* all_lines,
                        const int32_t* n_all, int n_frames, int cap, const uint32_t* seeds, int frame_count0,
                        double* vps, int32_t* vp_idx, double* line_vps, int32_t* status) {
  if (!c) return VPL_E_INVALID;
  if (n_frames == 0) return c->vp_ready ? VPL_OK : fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  int r = vpl_vp_submit(c, 0, lines, n_lines, all_lines, n_all, n_frames, cap, seeds, frame_count0);
  if (r) return r;
  return vpl_vp_collect(c, 0, cap, vps, vp_idx, line_vps, status);
}

int vpl_vp_pack_cloud(VplContext* c, int slot, const int32_t* line_ids, int cap, float fx, float fy, float cx, float cy,
                      int num_of_cam, int cam, float* cloud) {
  if (!c) return VPL_E_INVALID;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
  if (s.vp_n <= 0 || !s.vp_src_lines || !s.vp_src_n || !s.vp_src_hn)
    return fail(c, VPL_E_INVALID, "slot %d holds no collected vanishing-point batch", slot);
  if (!line_ids || !cloud || num_of_cam < 1 || cam < 0 || cam >= num_of_cam) return fail(c, VPL_E_INVALID, "bad argument");
  CK(c, cudaSetDevice(c->cfg.device));
  const int mcap = c->cfg.max_lines, n = s.vp_n;
  const int* na = s.vp_src_hn;  // counts of the line set the last vanishing-point run classified
  for (int i = 0; i < n; ++i) {
    if (na[i] < 0 || na[i] > mcap) return fail(c, VPL_E_INVALID, "frame %d: stale line count %d", i, na[i]);
    if (na[i] > cap) return fail(c, VPL_E_CAPACITY, "frame %d: %d lines > cap %d", i, na[i], cap);
    memcpy(s.h_vp_ids + (size_t)i * mcap, line_ids + (size_t)i * cap, (size_t)na[i] * sizeof(int));
  }
  CK(c, cudaMemcpyAsync(s.d_vp_ids, s.h_vp_ids, (size_t)n * mcap * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  {
    StageTimer t(c, s, VPL_STAGE_VP_CLASSIFY);
    launch_vp_cloud(s.vp_src_lines, s.vp_src_n, mcap, s.d_vp_ids,
                    s.d_line_vps, fx, fy, cx, cy, num_of_cam, cam, s.d_cloud, n, s.stream);
    t.launches(1);
  }
  CK(c, cudaMemcpyAsync(s.h_cloud, s.d_cloud, (size_t)n * mcap * 10 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
  int r = finish(c, s);
  if (r) return r;
  for (int i = 0; i < n; ++i)
    memcpy(cloud + (size_t)i * cap * 10, s.h_cloud + (size_t)i * mcap * 10, (size_t)na[i] * 10 * sizeof(float));
  return VPL_OK;
}

int vpl_vp_run_resident(VplContext* c, int slot) {
  if (!c) return VPL_E_INVALID;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (s.vp_n <= 0 || s.resident != BK_VP) return fail(c, VPL_E_INVALID, "slot %d holds no vanishing-point line sets", slot);
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  run_vp(c, s);
  return VPL_OK;
}

int vpl_debug_vp(VplContext* c, int frame, double* grid, int32_t* best_idx, int32_t* pairs) {
  if (!c) return VPL_E_INVALID;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  Slot& s = c->slots[0];
  if (frame < 0 || frame >= s.vp_n) return fail(c, VPL_E_INVALID, "frame %d outside the last batch", frame);
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaStreamSynchronize(s.stream));
  if (grid) CK(c, cudaMemcpy(grid, s.vp.grid_new + (size_t)frame * kVpCells, kVpCells * sizeof(double), cudaMemcpyDeviceToHost));
  if (best_idx) CK(c, cudaMemcpy(best_idx, s.vp.best_idx + frame, sizeof(int), cudaMemcpyDeviceToHost));
  if (pairs) CK(c, cudaMemcpy(pairs, s.vp.pairs + (size_t)frame * c->vpp.it * 2, (size_t)c->vpp.it * 2 * sizeof(int), cudaMemcpyDeviceToHost));
  return VPL_OK;
}

int vpl_debug_vp_scores(VplContext* c, int frame, double* scores) {
  if (!c || !scores) return VPL_E_INVALID;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  Slot& s = c->slots[0];
  if (frame < 0 || frame >= s.vp_n) return fail(c, VPL_E_INVALID, "frame %d outside the last batch", frame);
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaStreamSynchronize(s.stream));
  const size_t n = (size_t)c->vpp.it * 360;
  double* d = nullptr;
  CK(c, cudaMalloc((void**)&d, n * sizeof(double)));
  CK(c, cudaMemsetAsync(d, 0, n * sizeof(double), s.stream));
  launch_vp_scores_debug(s.vp, c->vpp, s.vp_n, frame, d, s.stream);
  c->launches += 1;
  cudaError_t e = cudaMemcpyAsync(scores, d, n * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(c, VPL_E_CUDA, "vpl_debug_vp_scores: %s", cudaGetErrorString(e));
  return VPL_OK;
}

static int linefront_collect_impl(VplContext* c, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur);

// ---- fused: readImage's line pipeline for n consecutive frames ------------------------------------
// (remap + CLAHE) -> EDline -> Matching(f-1, f) -> run_vanishing_point_detection on each frame's own lines
static int ri_enqueue(VplContext* c, Slot& s) {
  run_edlines(c, s);
  run_linematch(c, s, s.d_counts, s.n, s.n - 1, 1);
  run_vp_on(c, s, s.d_lines, s.d_counts, s.d_lines, s.d_counts, s.n, s.vp_fc0);
  return VPL_OK;
}

int vpl_readimage_submit(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                         int smoothed, const uint32_t* seeds, int frame_count0) {
  int r = lm_check(c);
  if (r) return r;
  r = ed_check(c, n, w, h, smoothed);
  if (r) return r;
  if (!c->vp_ready) return fail(c, VPL_E_INVALID, "call vpl_vp_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (n == 0 || !seeds) return fail(c, VPL_E_INVALID, "empty batch or no seeds");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  if ((c->pre_remap || c->pre_clip > 0.0) && !s.d_raw)
    CK(c, dmalloc(&s.d_raw, (size_t)c->cfg.max_batch * c->cfg.max_width * c->cfg.max_height));
  s.n = n; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 1; s.k = 0; s.ed_smoothed = smoothed ? 1 : 0;
  s.vp_n = n; s.vp_fc0 = frame_count0; s.vp_same = true;
  memcpy(s.h_vp_seeds, seeds, (size_t)n * sizeof(uint32_t));
  CK(c, cudaMemcpyAsync(s.d_vp_seeds, s.h_vp_seeds, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  ri_enqueue(c, s);
  ed_enqueue_download(c, s);
  {
    StageTimer t(c, s, VPL_STAGE_D2H);
    const int B = c->cfg.max_batch, mcap = c->cfg.max_lines;
    if (n > 1) {
      CK(c, cudaMemcpyAsync(s.h_r2c, s.lm.r2c, (size_t)(n - 1) * mcap * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      s.last_d2h_bytes += (int64_t)(n - 1) * mcap * sizeof(int);
    }
    CK(c, cudaMemcpyAsync(s.h_vps, s.d_vps, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CK(c, cudaMemcpyAsync(s.h_vp_idx, s.d_vp_idx, (size_t)n * mcap * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(c, cudaMemcpyAsync(s.h_vp_status, s.vp.status, (size_t)2 * B * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    s.last_d2h_bytes += (int64_t)n * (72 + (int64_t)mcap * 4) + 2 * B * 4;
  }
  CK(c, cudaMemcpyAsync(s.h_flags, s.d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(c, cudaEventRecord(s.done, s.stream));
  s.in_flight = true;
  s.kind = BK_READIMAGE;
  s.resident = BK_READIMAGE;
  s.vp_src_lines = s.d_lines;   // the stage ran on each frame's own detected lines
  s.vp_src_n = s.d_counts;
  s.vp_src_hn = s.h_counts;
  return VPL_OK;
}

int vpl_readimage_collect(VplContext* c, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur, double* vps,
                          int32_t* vp_idx, int32_t* vp_status) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_READIMAGE) return fail(c, VPL_E_INVALID, "slot %d has no readImage batch in flight", slot);
  if (!vps || !vp_idx) return fail(c, VPL_E_INVALID, "null output");
  const int n = s.n, mcap = c->cfg.max_lines, B = c->cfg.max_batch;
  int r = linefront_collect_impl(c, slot, lines, counts, cap, prev_to_cur);
  if (r) return r;
  memcpy(vps, s.h_vps, (size_t)n * 9 * sizeof(double));
  for (int i = 0; i < n; ++i) {
    memcpy(vp_idx + (size_t)i * cap, s.h_vp_idx + (size_t)i * mcap, (size_t)std::min(counts[i], cap) * sizeof(int));
    if (vp_status) vp_status[i] = s.h_vp_status[i] != 0 ? s.h_vp_status[i] : (s.h_vp_status[B + i] & 1);
  }
  return VPL_OK;
}

int vpl_readimage_run_resident(VplContext* c, int slot) {
  int r = lm_check(c);
  if (r) return r;
  if (!c->ed_ready || !c->vp_ready) return fail(c, VPL_E_INVALID, "configure EDLines, line matching and the vanishing points first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (s.n <= 0 || s.vp_n != s.n || s.resident != BK_READIMAGE) return fail(c, VPL_E_INVALID, "slot %d holds no readImage batch", slot);
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  if ((c->pre_remap || c->pre_clip > 0.0) && s.d_raw) {
    r = run_preprocess(c, s, s.d_raw, s.n, s.w, s.h);
    if (r) return r;
  }
  return ri_enqueue(c, s);
}

// ---- fused: EDline on every frame + Matching(frame f-1, frame f) ---------------------------------
int vpl_linefront_submit(VplContext* c, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                         int smoothed) {
  int r = lm_check(c);
  if (r) return r;
  r = ed_check(c, n, w, h, smoothed);
  if (r) return r;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (n == 0) return fail(c, VPL_E_INVALID, "empty batch");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[slot];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot %d still in flight: collect it first", slot);
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  s.n = n; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 1; s.k = 0; s.ed_smoothed = smoothed ? 1 : 0;
  r = upload(c, s, imgs, n, w, h, stride);
  if (r) return r;
  run_edlines(c, s);
  run_linematch(c, s, s.d_counts, n, n - 1, 1);
  ed_enqueue_download(c, s);
  if (n > 1) {
    StageTimer t(c, s, VPL_STAGE_D2H);
    CK(c, cudaMemcpyAsync(s.h_r2c, s.lm.r2c, (size_t)(n - 1) * c->cfg.max_lines * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    s.last_d2h_bytes += (int64_t)(n - 1) * c->cfg.max_lines * sizeof(int);
  }
  CK(c, cudaMemcpyAsync(s.h_flags, s.d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(c, cudaEventRecord(s.done, s.stream));
  s.in_flight = true;
  s.kind = BK_LINEFRONT;
  s.resident = BK_LINEFRONT;
  return VPL_OK;
}

// delivery shared by vpl_linefront_collect and vpl_readimage_collect (the caller has checked the batch kind)
static int linefront_collect_impl(VplContext* c, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur) {
  Slot& s = c->slots[slot];
  if (!lines || !counts || !prev_to_cur) return fail(c, VPL_E_INVALID, "null output");
  CK(c, cudaSetDevice(c->cfg.device));
  int r = finish(c, s);
  if (r) return r;
  r = ed_deliver(c, s, lines, counts, cap, nullptr);
  if (r) return r;
  if (s.h_flags[3]) return fail(c, VPL_E_CAPACITY, "a frame pair needs more than max_anchors=%d anchor points", c->lmp.max_anchors);
  const int mcap = c->cfg.max_lines;
  for (int i = 0; i < cap; ++i) prev_to_cur[i] = -1;
  for (int f = 1; f < s.n; ++f) {
    int32_t* row = prev_to_cur + (size_t)f * cap;
    const int np = s.h_counts[f - 1];
    memcpy(row, s.h_r2c + (size_t)(f - 1) * mcap, (size_t)np * sizeof(int));
    for (int i = np; i < cap; ++i) row[i] = -1;
  }
  return VPL_OK;
}

int vpl_linefront_collect(VplContext* c, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur) {
  if (!c) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (!s.in_flight || s.kind != BK_LINEFRONT) return fail(c, VPL_E_INVALID, "slot %d has no line front-end batch in flight", slot);
  return linefront_collect_impl(c, slot, lines, counts, cap, prev_to_cur);
}

int vpl_linefront_batch(VplContext* c, const uint8_t* const* imgs, int n, int w, int h, size_t stride, int smoothed,
                        VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur) {
  int r = lm_check(c);
  if (r) return r;
  r = ed_check(c, n, w, h, smoothed);
  if (r) return r;
  if (n == 0) return VPL_OK;
  r = vpl_linefront_submit(c, 0, imgs, n, w, h, stride, smoothed);
  if (r) return r;
  return vpl_linefront_collect(c, 0, lines, counts, cap, prev_to_cur);
}

int vpl_linefront_run_resident(VplContext* c, int slot) {
  int r = lm_check(c);
  if (r) return r;
  if (!c->ed_ready) return fail(c, VPL_E_INVALID, "call vpl_edlines_configure first");
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  Slot& s = c->slots[slot];
  if (s.n <= 0) return fail(c, VPL_E_INVALID, "slot %d holds no frames", slot);
  CK(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.profile) {
    bool any = false;
    for (int i = 0; i < VPL_NUM_STAGES; ++i) any |= s.ev_used[i];
    if (any) { cudaStreamSynchronize(s.stream); harvest_times(c, s); }
  }
  run_edlines(c, s);
  run_linematch(c, s, s.d_counts, s.n, s.n - 1, 1);
  return VPL_OK;
}

// ---- raw stages for the parity tests -------------------------------------------
int vpl_lsd_raw(VplContext* c, const uint8_t* img, int w, int h, size_t stride, VplSegment* out, int32_t* count,
                int cap) {
  int r = check_dims(c, 1, w, h, 1, 1, true);
  if (r) return r;
  if (!img || !out || !count) return fail(c, VPL_E_INVALID, "null argument");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  s.n = 1; s.w = w; s.h = h; s.num_octaves = 1; s.scale = 1; s.k = 0;
  const uint8_t* one[1] = {img};
  r = upload(c, s, one, 1, w, h, stride);
  if (r) return r;
  run_pyramid(c, s, 0);  // no pyramid blur: plain cv::LineSegmentDetector on the image
  run_lsd(c, s);
  launch_pack_segments(s.oct[0].cand, s.oct[0].n_cand, c->cand_cap, s.d_seg, s.d_seg_count, c->cand_cap, 1, s.stream);
  c->launches += 1;
  int n = 0;
  CK(c, cudaMemcpyAsync(&n, s.d_seg_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(c, cudaMemcpyAsync(s.h_flags, s.d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  r = finish(c, s);
  if (r) return r;
  if (s.h_flags[0]) return fail(c, VPL_E_CAPACITY, "more than %d LSD candidates", c->cand_cap);
  *count = n;
  if (n > cap) return fail(c, VPL_E_CAPACITY, "%d segments > cap %d", n, cap);
  CK(c, cudaMemcpy(out, s.d_seg, (size_t)n * sizeof(VplSegment), cudaMemcpyDeviceToHost));
  return VPL_OK;
}

int vpl_debug_stage(VplContext* c, int which, const uint8_t* img, int w, int h, size_t stride, void* out,
                    size_t out_bytes, int32_t* out_w, int32_t* out_h) {
  int r = check_dims(c, 1, w, h, which == 1 ? 2 : 1, 2, true);
  if (r) return r;
  if (!img || !out || !out_w || !out_h) return fail(c, VPL_E_INVALID, "null argument");
  if (which == 1 && c->cfg.max_octaves < 2) return fail(c, VPL_E_CAPACITY, "pyrDown stage needs max_octaves >= 2");
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  if (s.in_flight) return fail(c, VPL_E_INVALID, "slot 0 in flight");
  s.n = 1; s.w = w; s.h = h; s.num_octaves = (which == 1) ? 2 : 1; s.scale = 2; s.k = 0;
  const uint8_t* one[1] = {img};
  r = upload(c, s, one, 1, w, h, stride);
  if (r) return r;
  int wo, ho, ws, hs;
  octave_geom(w, h, 0, wo, ho, ws, hs);
  const void* src = nullptr;
  size_t bytes = 0;
  int n_ord = 0;
  switch (which) {
    case 0:  // GaussianBlur 5x5
      run_pyramid(c, s, 1);
      src = s.oct[0].pyr; bytes = (size_t)w * h; *out_w = w; *out_h = h;
      break;
    case 1:  // pyrDown of the UNBLURRED image
      s.num_octaves = 2;
      run_pyramid(c, s, 0);
      src = s.oct[1].pyr; bytes = (size_t)(w / 2) * (h / 2); *out_w = w / 2; *out_h = h / 2;
      break;
    case 2:  // Sobel of the unblurred image
      run_pyramid(c, s, 0);
      src = s.oct[0].grad; bytes = (size_t)w * h * sizeof(short2); *out_w = w; *out_h = h;
      break;
    case 3:
    case 4:
    case 5:
      run_pyramid(c, s, 0);
      CK(c, cudaMemsetAsync(s.oct[0].maxq, 0, sizeof(unsigned int), s.stream));
      launch_scale08(s.oct[0].pyr, s.oct[0].scl, w, h, ws, hs, 1, s.stream);
      c->launches += 1;
      if (which == 3) { src = s.oct[0].scl; bytes = (size_t)ws * hs; *out_w = ws; *out_h = hs; break; }
      launch_ll_angle(s.oct[0].scl, s.oct[0].ang, s.oct[0].pix, c->d_cssn_lut, s.oct[0].maxq, ws, hs, 1, c->lc.rho, s.stream);
      c->launches += 1;
      if (which == 4) { src = s.oct[0].ang; bytes = (size_t)ws * hs * sizeof(float); *out_w = ws; *out_h = hs; break; }
      launch_order(s.oct[0].scl, s.oct[0].maxq, s.oct[0].ord, s.oct[0].n_ord, s.oct[0].reg, (size_t)ws * hs * sizeof(uint32_t), ws,
                   hs, 1, c->lc.rho, s.stream);
      c->launches += 1;
      CK(c, cudaMemcpyAsync(&n_ord, s.oct[0].n_ord, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CK(c, cudaStreamSynchronize(s.stream));
      src = s.oct[0].ord; bytes = (size_t)n_ord * sizeof(int); *out_w = n_ord; *out_h = 1;
      break;
    default:
      return fail(c, VPL_E_INVALID, "unknown stage %d", which);
  }
  r = finish(c, s);
  if (r) return r;
  if (bytes > out_bytes) return fail(c, VPL_E_CAPACITY, "output needs %zu bytes, %zu given", bytes, out_bytes);
  CK(c, cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
  return VPL_OK;
}

int vpl_debug_candidates(VplContext* c, double* out, int32_t* count, int cap) {
  if (!c || !out || !count) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  Slot& s = c->slots[0];
  int n = 0;
  CK(c, cudaMemcpy(&n, s.oct[0].n_cand, sizeof(int), cudaMemcpyDeviceToHost));
  *count = n;
  if (n > cap) n = cap;
  std::vector<RectCand> tmp((size_t)n);
  CK(c, cudaMemcpy(tmp.data(), s.oct[0].cand, (size_t)n * sizeof(RectCand), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) {
    const RectCand& r = tmp[i];
    const double v[16] = {r.x1, r.y1, r.x2, r.y2, r.width, r.x, r.y, r.theta, r.dx, r.dy, r.prec, r.p,
                          r.nfa, (double)r.accepted, 0, 0};
    memcpy(out + (size_t)16 * i, v, sizeof(v));
  }
  return VPL_OK;
}

#ifdef VPL_DEBUG_NFA
extern "C" int vpl_debug_set_nfa_cand(int c) { vpl::debug_set_cand(c); return 0; }
#endif

// ---- measurement -------------------------------------------------------------------
int vpl_get_stage_times(VplContext* c, double* ms, int64_t* launches) {
  if (!c) return VPL_E_INVALID;
  for (int i = 0; i < VPL_NUM_STAGES; ++i) {
    if (ms) ms[i] = c->stage_ms[i];
    if (launches) launches[i] = c->stage_launches[i];
  }
  return VPL_OK;
}

int vpl_debug_set_engine_ring_cap(VplContext* c, int entries_per_lane) {
  if (!c) return VPL_E_INVALID;
  if (entries_per_lane != 0 && entries_per_lane < 64) return fail(c, VPL_E_INVALID, "ring capacity must be 0 (default) or >= 64 entries");
  c->engine_ring_cap = entries_per_lane;
  return VPL_OK;
}

int vpl_debug_set_engine(VplContext* c, int kind) {
  if (!c) return VPL_E_INVALID;
  if (kind < 0 || kind > 1) return fail(c, VPL_E_INVALID, "engine kind must be 0 (default, warp-cooperative) or 1 (speculative)");
  if (kind == 1) {
    if (!c->cfg.lsd_path) return fail(c, VPL_E_INVALID, "this context was created with lsd_path = 0");
    CK(c, cudaSetDevice(c->cfg.device));
    // the speculative engine keeps owner tags, a two-list arena and its parking buffers per frame: allocated here,
    // for batches of at most kSpecMaxBatch frames
    const size_t sb = std::min<size_t>((size_t)c->cfg.max_batch, (size_t)kSpecMaxBatch);
    const size_t P0 = (size_t)c->cfg.max_width * c->cfg.max_height;
    for (Slot& s : c->slots)
      for (int o = 0; o < c->cfg.max_octaves; ++o) {
        OctBuf& b = s.oct[o];
        if (b.spec_tag) continue;
        size_t Po = (P0 >> (2 * o)) + 64;
        size_t So = (size_t)((double)Po * 0.64) + 2 * (size_t)(c->cfg.max_width + c->cfg.max_height) + 64;
        CK(c, dmalloc(&b.spec_tag, sb * So));
        CK(c, dmalloc(&b.spec_arena, sb * So * 2));
        CK(c, dmalloc(&b.eng_desc, sb * 32 * kEngQ));
        CK(c, dmalloc(&b.eng_rects, sb * 32 * kEngQ));
      }
  }
  c->engine_kind = kind;
  return VPL_OK;
}

int vpl_set_profile(VplContext* c, int on) {
  if (!c) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  for (Slot& s : c->slots) {  // bank what the slots' events still hold before the switch
    CK(c, cudaStreamSynchronize(s.stream));
    harvest_times(c, s);
  }
  c->cfg.profile = on ? 1 : 0;
  return VPL_OK;
}

int vpl_reset_stage_times(VplContext* c) {
  if (!c) return VPL_E_INVALID;
  memset(c->stage_ms, 0, sizeof(c->stage_ms));
  memset(c->stage_launches, 0, sizeof(c->stage_launches));
  return VPL_OK;
}

int vpl_debug_mark(VplContext* c) {
  if (!c) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  if (!c->mark) CK(c, cudaEventCreate(&c->mark));
  CK(c, cudaEventRecord(c->mark, c->slots[0].stream));
  return VPL_OK;
}

int vpl_debug_timeline(VplContext* c, int slot, double* start_ms, double* end_ms) {
  if (!c || !start_ms || !end_ms) return VPL_E_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return fail(c, VPL_E_INVALID, "bad slot %d", slot);
  if (!c->mark) return fail(c, VPL_E_INVALID, "vpl_debug_mark has not been called");
  Slot& s = c->slots[slot];
  CK(c, cudaSetDevice(c->cfg.device));
  CK(c, cudaStreamSynchronize(s.stream));
  for (int i = 0; i < VPL_NUM_STAGES; ++i) {
    start_ms[i] = end_ms[i] = -1.0;
    if (!s.ev_used[i]) continue;
    float a = 0, b = 0;
    if (cudaEventElapsedTime(&a, c->mark, s.ev[i][0]) == cudaSuccess &&
        cudaEventElapsedTime(&b, c->mark, s.ev[i][1]) == cudaSuccess) {
      start_ms[i] = a; end_ms[i] = b;
    }
  }
  cudaGetLastError();
  return VPL_OK;
}

int vpl_debug_nfa_stats(VplContext* c, uint64_t* out4) {
  if (!c || !out4) return VPL_E_INVALID;
  CK(c, cudaSetDevice(c->cfg.device));
  unsigned long long v[4];
  nfa_check_stats(v);
  for (int i = 0; i < 4; ++i) out4[i] = v[i];
  return VPL_OK;
}

int64_t vpl_kernel_launches(const VplContext* c) { return c ? c->launches : 0; }

}  // extern "C"
