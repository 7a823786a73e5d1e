/*
 * orc_threads.c -- CPU oracle: run the whole front end (LSDDetector::detect ->
 * BinaryDescriptor::compute -> brute-force match against the previous frame)
 * over a frame sequence on T host threads, for the CPU baseline of bench.py.
 * TEST / BASELINE INFRASTRUCTURE ONLY (see vpl_oracle.h).  Frames are split in
 * contiguous chunks with a one-frame halo, the same partition the multi-GPU
 * driver uses.
 */
#define _GNU_SOURCE
#include "vpl_oracle.h"
#include <malloc.h>
#include <pthread.h>
#include <stdlib.h>

typedef struct {
  const uint8_t* frames;
  int f0, f1, w, h, num_octaves, max_lines;
  int64_t total;
} Job;

static void* worker(void* a) {
  Job* j = (Job*)a;
  /* halo: also describe frame f0-1 so that pair (f0-1, f0) is matched here */
  int start = j->f0 > 0 ? j->f0 - 1 : 0;
  j->total = orc_frontend_sequence(j->frames + (size_t)start * j->w * j->h, j->f1 - start, j->w,
                                   j->h, j->num_octaves, j->max_lines);
  return NULL;
}

int64_t orc_frontend_sequence_mt(const uint8_t* frames, int n_frames, int w, int h,
                                 int num_octaves, int max_lines, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_frames) n_threads = n_frames;
  /* the per-frame scratch images are MBs each: keep them in the per-thread arenas instead of
   * mmap/munmap per frame, which serialises the threads in the kernel (a fair CPU baseline) */
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  pthread_t* th = (pthread_t*)malloc((size_t)n_threads * sizeof(pthread_t));
  Job* jobs = (Job*)malloc((size_t)n_threads * sizeof(Job));
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].frames = frames;
    jobs[t].f0 = (int)((int64_t)n_frames * t / n_threads);
    jobs[t].f1 = (int)((int64_t)n_frames * (t + 1) / n_threads);
    jobs[t].w = w; jobs[t].h = h; jobs[t].num_octaves = num_octaves; jobs[t].max_lines = max_lines;
    jobs[t].total = 0;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    total += jobs[t].total;
  }
  free(th); free(jobs);
  return total;
}

/* EDLines over a frame sequence on n_threads host threads (frames are independent). */
typedef struct {
  const uint8_t* frames;
  int f0, f1, w, h, smoothed;
  const OrcEDLineParam* p;
  int64_t total;
} EdJob;

static void* ed_worker(void* a) {
  EdJob* j = (EdJob*)a;
  OrcLine* out = (OrcLine*)malloc(sizeof(OrcLine) * 8192);
  for (int f = j->f0; f < j->f1; ++f)
    j->total += orc_edline_detect(j->frames + (size_t)f * j->w * j->h, j->w, j->h, j->p, j->smoothed, out, 8192,
                                  NULL, NULL, NULL, NULL);
  free(out);
  return NULL;
}

int64_t orc_edline_sequence_mt(const uint8_t* frames, int n_frames, int w, int h, const OrcEDLineParam* p,
                               int smoothed, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  pthread_t* th = (pthread_t*)malloc((size_t)n_threads * sizeof(pthread_t));
  EdJob* jobs = (EdJob*)malloc((size_t)n_threads * sizeof(EdJob));
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].frames = frames;
    jobs[t].f0 = (int)((int64_t)n_frames * t / n_threads);
    jobs[t].f1 = (int)((int64_t)n_frames * (t + 1) / n_threads);
    jobs[t].w = w; jobs[t].h = h; jobs[t].smoothed = smoothed; jobs[t].p = p; jobs[t].total = 0;
    pthread_create(&th[t], NULL, ed_worker, &jobs[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    total += jobs[t].total;
  }
  free(th); free(jobs);
  return total;
}

/* The tracker's per-frame hot loop (EDline on every frame + Matching(prev, cur) on every consecutive
 * pair, feature_tracker/src/line_feature_tracker.cpp:87, :115) over n_threads host threads:
 * contiguous chunks with a one-frame halo.  Returns the number of matched lines. */
typedef struct {
  const uint8_t* frames;
  int f0, f1, w, h, smoothed;
  const OrcEDLineParam* p;
  int64_t total;
} LfJob;

static void* lf_worker(void* a) {
  LfJob* j = (LfJob*)a;
  enum { CAP = 8192 };
  OrcLine* prev = (OrcLine*)malloc(sizeof(OrcLine) * CAP);
  OrcLine* cur = (OrcLine*)malloc(sizeof(OrcLine) * CAP);
  int32_t* r2c = (int32_t*)malloc(sizeof(int32_t) * CAP);
  OrcLineMatchParam mp;
  orc_lm_default_param(&mp);
  int n_prev = 0;
  int start = j->f0 > 0 ? j->f0 - 1 : 0;
  size_t fsz = (size_t)j->w * j->h;
  for (int f = start; f < j->f1; ++f) {
    int n_cur = orc_edline_detect(j->frames + f * fsz, j->w, j->h, j->p, j->smoothed, cur, CAP, NULL, NULL, NULL, NULL);
    if (n_cur > CAP) n_cur = CAP;
    if (f > start &&
        orc_line_matching(j->frames + (f - 1) * fsz, j->frames + f * fsz, j->w, j->h, prev, n_prev, cur, n_cur, &mp, 1, 1,
                          r2c, NULL, NULL, NULL, NULL, NULL, 0, NULL))
      for (int i = 0; i < n_prev; ++i) j->total += r2c[i] >= 0;
    OrcLine* t = prev; prev = cur; cur = t;
    n_prev = n_cur;
  }
  free(prev); free(cur); free(r2c);
  return NULL;
}

int64_t orc_linefront_sequence_mt(const uint8_t* frames, int n_frames, int w, int h, const OrcEDLineParam* p,
                                  int smoothed, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  pthread_t* th = (pthread_t*)malloc((size_t)n_threads * sizeof(pthread_t));
  LfJob* jobs = (LfJob*)malloc((size_t)n_threads * sizeof(LfJob));
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].frames = frames;
    jobs[t].f0 = (int)((int64_t)n_frames * t / n_threads);
    jobs[t].f1 = (int)((int64_t)n_frames * (t + 1) / n_threads);
    jobs[t].w = w; jobs[t].h = h; jobs[t].smoothed = smoothed; jobs[t].p = p; jobs[t].total = 0;
    pthread_create(&th[t], NULL, lf_worker, &jobs[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    total += jobs[t].total;
  }
  free(th); free(jobs);
  return total;
}
