#!/usr/bin/env python
"""bench.py -- line front-end throughput (LSD + LBD + Hamming match), frames/s.

Workload (BASELINE.json configs[1], "C2"): synthetic EuRoC-shaped 752x480 mono frames,
single-octave LSDDetector::detect + BinaryDescriptor::compute + consecutive-frame
match (k=1).  One STEP = one batch of --batch frames through the whole hot path.

  value  whole-job frames/s with the batches already resident in HBM (kernels only,
         results left in HBM), the slots' batches enqueued as groups
         (vpl_frontend_run_resident_group);
  e2e    the same through the reference-facing C ABI with HOST buffers, everything inside
         the timed region: frames in pinned caller memory copied to the device ahead of
         their turn (vpl_frontend_upload), the slots' batches submitted as a group
         (vpl_frontend_submit_group) before the previous group is collected, dense rows
         downloaded into pinned caller buffers (vpl_frontend_collect_dense);
  roofline      dominant kernel (the LSD region engine): algorithmic bytes / its mean
                launch duration (CUDA events on its own stream) vs the measured HBM peak;
  cpu_baseline  the CPU oracle (port of the path, oracle/) on the host cores, bounded sample.

`--impl reference` times the CPU implementation of the path alone (the reference's
own line_descriptor binary cannot be built here: no OpenCV C++ / contrib; the oracle
port stands in).

Multi-GPU: one process per GPU (torchrun), frames sharded by contiguous range with a
one-frame halo, no data-path collective; timing is max over ranks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "C2_euroc_752x480"
METRIC = "line front-end frames/sec (LSD+LBD+match) at 752x480"
# other parity/bench cases of BASELINE.json (not the headline): --workload C1|C3|C4
WORKLOADS = {
    "C2": dict(name="C2_euroc_752x480", w=752, h=480, octaves=1, k=1, max_lines=1024),
    "C1": dict(name="C1_mh04_real_752x480", w=752, h=480, octaves=1, k=1, max_lines=2048),
    "C3": dict(name="C3_d455_1280x720", w=1280, h=720, octaves=2, k=2, max_lines=2048),
    "C4": dict(name="C4_manhattan_1920x1080", w=1920, h=1080, octaves=1, k=2, max_lines=6144),
    # BASELINE.json configs[4]: ONE sequence of 1280x720 frames (the C3 generator, seed + 1), 2 octaves, knnMatch k=2,
    # strong-sharded by contiguous frame range + 1-frame halo over the ranks; results hashed per frame and compared
    # with the same sequence run on one GPU
    "C5": dict(name="C5_sharded_sequence_1280x720", w=1280, h=720, octaves=2, k=2, max_lines=2048, frames="C3", sharded=True),
    # the detector the reference really runs (EDLineDetector::EDline with the tracker
    # node's parameters, smoothed=true) on the C2 frames / on the reference's bundled frames
    "E1": dict(name="E1_edlines_euroc_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C2"),
    "E1r": dict(name="E1r_edlines_mh04_real_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C1"),
    # SURVEY 8f-1 + 8f-2: the reference's whole per-frame line front end as it really runs
    # (readImage: EDline on every frame + LineMatching::Matching(prev, cur), line_feature_tracker.cpp:87, :115)
    "E2": dict(name="E2_edlines_kltmatch_euroc_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C2", match=True),
    "E2r": dict(name="E2r_edlines_kltmatch_mh04_real_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C1", match=True),
    # SURVEY 8f-1..4 chained: readImage's whole line pipeline on raw frames in one pass over the device
    # (remap + CLAHE -> EDline -> Matching(prev, cur) -> vanishing points on each frame's lines), vpl_readimage_*
    "R1": dict(name="R1_readimage_line_pipeline_mh04_real_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C1",
               match=True, full=True),
    "R1s": dict(name="R1s_readimage_line_pipeline_euroc_752x480", w=752, h=480, octaves=1, k=0, max_lines=512, frames="C2",
                match=True, full=True),
    # the vanishing-point stage readImage runs on every frame's lines after matching
    # (vanishing_point_detection::run_vanishing_point_detection, line_feature_tracker.cpp:233-262); the line sets
    # are the EDLines of the C2 / mh04 frames
    "V1": dict(name="V1_vanishing_points_euroc_752x480", w=752, h=480, octaves=1, k=0, max_lines=256, frames="C2", vp=True),
    "V1r": dict(name="V1r_vanishing_points_mh04_real_752x480", w=752, h=480, octaves=1, k=0, max_lines=256, frames="C1", vp=True),
    # SURVEY 8d / C4's matcher alone: brute-force Hamming kNN, 2000 x 2000 256-bit codes per frame pair, k = 2
    "M4": dict(name="M4_hamming_knn_2000x2000_k2", w=64, h=64, octaves=1, k=2, max_lines=2000, lq=2000, lt=2000, matcher=True),
}
M_METRIC = "brute-force Hamming kNN matching frame-pairs/sec (BinaryDescriptorMatcher::knnMatch, 2000x2000 256-bit codes, k=2)"
VP_METRIC = "vanishing-point stage frames/sec (vanishing_point_detection::run_vanishing_point_detection) on 752x480 line sets"
EUROC_CAM = (461.6, 363.0, 248.1)  # fx, cx, cy of config/euroc/euroc_config.yaml
RI_METRIC = ("reference line pipeline frames/sec (LineFeatureTracker::readImage: remap + CLAHE, EDLines, KLT line matching, "
             "vanishing points) at 752x480")
LF_METRIC = "reference line front-end frames/sec (EDLines + KLT line matching, LineFeatureTracker::readImage) at 752x480"
ED_METRIC = "EDLines line detection frames/sec (the reference's EDLineDetector::EDline) at 752x480"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(n_unique, seed, workload="C2"):
    if WORKLOADS[workload].get("sharded"):
        seed += 1  # C5 = the C3 generator with seed + 1
    workload = WORKLOADS[workload].get("frames", workload)
    if workload == "C1":  # the reference's bundled EuRoC MH_04 frames (tests/golden/mh04_frames.npz)
        return np.load(os.path.join(ROOT, "tests", "golden", "mh04_frames.npz"))["frames"]
    synth = importlib.import_module("vplines-slam_b200.synth")
    return synth.config_sequence(WORKLOADS[workload]["name"], n_unique, seed=seed)


def tile_frames(unique, n):
    reps = (n + len(unique) - 1) // len(unique)
    return np.ascontiguousarray(np.concatenate([unique] * reps)[:n])


def lsd_metric(wl):
    return METRIC.replace("752x480", "%dx%d" % (wl["w"], wl["h"]))


def lsd_config(args, wl, world):
    """The `config` object of the LSD/LBD/match workloads: the same for the B200 arm and for --impl reference
    (the reference arm runs a bounded sample of this workload; how many frames it actually ran is in its
    cpu_baseline.sample and `sample_frames_per_step`)."""
    W, H = wl["w"], wl["h"]
    return {"workload": wl["name"], "frames_per_step": args.batch, "width": W, "height": H, "octaves": wl["octaves"],
            "match_k": wl["k"], "unique_frames": args.unique, "slots": args.slots,
            "max_lines": args.max_lines or wl["max_lines"],
            "e2e_pipeline": ("frames uploaded ahead on a copy stream" if args.upload_ahead else "upload at submit") +
                            ("; the slots' batches submitted back to back" if args.e2e_together else "; one slot at a time"),
            "parallelism": f"frames x{world}" + (f" (rank r's frames: the {args.unique} distinct frames rotated by r*{args.unique}//{world})"
                                                 if world > 1 else ""),
            "l2": "inputs per step (%.0f MB) and per-step working set exceed the 126 MB L2" % (args.batch * W * H / 1e6)}


def verify_dense_step(frames, prev_frame, counts, kl, desc, mt, K, octaves, check):
    """Parity of one collected step against the CPU oracle chain (LSDDetector::detect -> BinaryDescriptor::compute ->
    match(frame, previous frame)): `frames` are the step's frames, `prev_frame` the frame frame 0 was matched
    against (None: no chaining), counts/kl/desc/mt the dense outputs of vpl_frontend_collect_dense, `check` the
    frame indices to compare.  KeyLines (17 fields), descriptors, match indices and distances must be bit-equal.
    -> list of (frame, what) mismatches."""
    from oracle import oracle as O
    n = len(frames)
    off = np.concatenate([[0], np.cumsum(np.asarray(counts[:n], np.int64))])
    cache = {}

    def orc(img, key):
        if key not in cache:
            ekl = O.lsd_detector_detect(img, 2, octaves)
            cache[key] = (ekl, O.lbd_compute(img, ekl))
        return cache[key]

    bad = []
    for f in check:
        ekl, ed = orc(frames[f], f)
        a, b = int(off[f]), int(off[f + 1])
        if b - a != len(ekl) or kl[a:b].tobytes() != ekl.tobytes():
            bad.append((f, "keylines"))
            continue
        if not np.array_equal(desc[a:b], ed):
            bad.append((f, "descriptors"))
            continue
        if K <= 0:
            continue
        m = mt[a:b].reshape(b - a, K)
        prev = frames[f - 1] if f > 0 else prev_frame
        if prev is None:
            if not ((m["trainIdx"] == -1).all() and (m["queryIdx"][:, 0] == np.arange(b - a)).all()):
                bad.append((f, "no-match fill"))
            continue
        pkl, pd = orc(prev, f - 1 if f > 0 else "prev")
        if len(pd) < K or len(ed) == 0:
            continue
        idx, dist = O.hamming_knn(ed, pd, K)
        if not (np.array_equal(m["trainIdx"], idx) and np.array_equal(m["distance"].astype(np.int32), dist)
                and (m["queryIdx"] == np.arange(b - a)[:, None]).all()):
            bad.append((f, "matches"))
    return bad


def cv2_crosscheck(unique, octaves=1, k=1, n=8):
    """BASELINE.md 3.2 / SURVEY 8(d): the stages cv2 4.13 covers (GaussianBlur 5x5 + LSD_REFINE_ADV per octave +
    BFMatcher on 32-byte codes; it has no LBD), 1 thread and default threads, frames/s on `n` frames."""
    try:
        import cv2
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)}
    out = {"version": cv2.__version__, "frames": n,
           "stages": "GaussianBlur(5x5,1) + createLineSegmentDetector(LSD_REFINE_ADV).detect per octave + "
                     "BFMatcher(NORM_HAMMING).knnMatch on L x 32 random bytes (cv2 has no LBD)"}
    rng = np.random.default_rng(0)
    default_threads = cv2.getNumThreads()
    for label, th in (("one_thread", 1), ("default_threads", default_threads)):
        cv2.setNumThreads(th)
        det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        prev = None
        t = time.time()
        for i in range(n):
            img = unique[i % len(unique)]
            b = cv2.GaussianBlur(img, (5, 5), 1)
            L = 0
            for o in range(octaves):
                seg = det.detect(b)[0]
                L += 0 if seg is None else len(seg)
                if o + 1 < octaves:
                    b = cv2.pyrDown(b)
            d = rng.integers(0, 256, (max(L, 1), 32), dtype=np.uint8)
            if prev is not None:
                bf.knnMatch(d, prev, k=k)
            prev = d
        out[label + "_frames_per_s"] = n / (time.time() - t)
    out["default_threads"] = default_threads
    cv2.setNumThreads(default_threads)
    return out


def cpu_baseline(unique, seconds=12.0, threads=None, octaves=1, name=WORKLOAD):
    """The oracle (CPU port of the path) on the host cores: bounded sample of the same frames."""
    from oracle import oracle as O
    O.build()
    threads = threads or os.cpu_count() or 1
    t = time.time()
    O.frontend_sequence(unique[:4], num_octaves=octaves, max_lines=16384, threads=1)
    per_frame = max((time.time() - t) / 4, 1e-3)
    n = int(max(threads * 2, min(seconds / per_frame * threads, 4096)))
    frames = tile_frames(unique, n)
    t = time.time()
    total = O.frontend_sequence(frames, num_octaves=octaves, max_lines=16384, threads=threads)
    dt = time.time() - t
    return {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{n} frames of {name} ({total} keylines) in {dt:.1f}s; CPU oracle (C port of cv2-4.13 LSD + "
                      f"opencv_contrib-3.4 LBD + brute-force Hamming), {threads} pthreads, contiguous chunks + 1-frame halo",
            "single_thread_frames_per_s": 1.0 / per_frame}


def ed_cpu_baseline(unique, seconds=10.0, threads=None, name="E1", match=False):
    """EDLines on the host cores: the reference's own edline_detector.cpp (oracle/_ref, built in the
    authoring container against oracle/cvshim) when that library travelled here, else the oracle port."""
    from oracle import oracle as O
    O.build()
    use_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_linefront.so"))
    threads = threads or os.cpu_count() or 1
    run = O.linefront_sequence if match else O.edline_sequence
    t = time.time()
    run(unique[:8], threads=1, use_ref=use_ref)
    per_frame = max((time.time() - t) / 8, 1e-4)
    n = int(max(threads * 4, min(seconds / per_frame * threads, 16384)))
    frames = tile_frames(unique, n)
    t = time.time()
    total = run(frames, threads=threads, use_ref=use_ref)
    dt = time.time() - t
    src = "edline_detector.cpp" + (" + line_matching.cpp + lk_tracker_invoker_2d.cpp" if match else "")
    what = (f"the reference's own line_matching/src/{src} compiled against oracle/cvshim (OpenCV stand-in)"
            if use_ref else f"CPU oracle port of {src} (oracle/orc_edlines.c, orc_linematch.c)")
    return {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{n} frames of {name} ({total} {'matched lines' if match else 'lines'}) in {dt:.1f}s; {what}, "
                      f"one detector{'/matcher' if match else ''} per thread, {threads} threads over contiguous frame "
                      f"chunks{' with a one-frame halo' if match else ''}",
            "single_thread_frames_per_s": 1.0 / per_frame}


def run_reference_edlines(args):
    from oracle import oracle as O
    O.build()
    wl = WORKLOADS[args.workload]
    unique = make_frames(min(args.unique, 32), args.seed, args.workload)
    use_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_linefront.so"))
    threads = os.cpu_count() or 1
    match = bool(wl.get("match"))
    run = O.linefront_sequence if match else O.edline_sequence
    t = time.time()
    run(unique[:8], threads=1, use_ref=use_ref)
    per_frame = max((time.time() - t) / 8, 1e-4)
    n = int(max(threads * 4, min(5.0 / per_frame * threads, 16384)))
    frames = tile_frames(unique, n)
    for _ in range(args.warmup):
        run(frames[:max(threads, n // 4)], threads=threads, use_ref=use_ref)
    t0 = time.time()
    for _ in range(args.steps):
        run(frames, threads=threads, use_ref=use_ref)
    dt = time.time() - t0
    fps = n * args.steps / dt
    kind = "reference" if use_ref else "port"
    line = {"impl": "reference", "metric": LF_METRIC if match else ED_METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/s16/f64", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_step": n,
                       "note": "the reference's own line_matching/src sources compiled against oracle/cvshim" if use_ref
                       else "CPU oracle port (oracle/_ref not present)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": f"{n} frames/step x {args.steps} steps"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_edlines(args, torch, dist, rank, local_rank, world):
    """--workload E1 / E1r: EDLineDetector::EDline (tracker-node parameters, smoothed=true) over a batch;
    E2 / E2r: the same followed by LineMatching::Matching(frame f-1, frame f) for every frame (fused)."""
    vpl = importlib.import_module("vplines_slam_b200")
    capi = vpl.capi
    wl = WORKLOADS[args.workload]
    W, H = wl["w"], wl["h"]
    B, S, cap = args.batch, args.slots, (args.max_lines or wl["max_lines"])
    match = bool(wl.get("match"))
    full = bool(wl.get("full"))  # R1: pre-processing in front, vanishing points behind (vpl_readimage_*)
    unique = make_frames(args.unique, args.seed, args.workload)
    # E2: every step submits its B frames plus the last frame of the previous step in front (one-frame
    # overlap), so that each consecutive pair of the sequence is matched exactly once
    NB = B + 1 if match else B
    host_buf = np.ascontiguousarray(np.concatenate([unique[-1:], tile_frames(unique, B)])) if match else tile_frames(unique, B)
    ctx = capi.Context(device=local_rank, max_width=W, max_height=H, max_octaves=1, max_lines=cap, max_batch=NB,
                       num_slots=S, blur_first=True, profile=True, lsd_path=False)
    param = capi.EDLineParam()
    ctx.edlines_configure(param)
    if match:
        ctx.linematch_configure(capi.LineMatchParam())
    if full:
        ctx.vp_configure(*EUROC_CAM)
        ctx.set_preprocess(*euroc_maps(W, H), clahe_clip=3.0, clahe_tiles=8)
        seeds = (1700000000 + np.arange(NB)).astype(np.uint32)
        vps = [np.zeros((NB, 3, 3), np.float64) for _ in range(S)]
        vp_idx = [np.zeros((NB, cap), np.int32) for _ in range(S)]
        vp_st = [np.zeros(NB, np.int32) for _ in range(S)]
    ctx.host_register(host_buf)
    lines = [np.zeros((NB, cap), capi.LINE_DTYPE) for _ in range(S)]
    counts = [np.zeros(NB, np.int32) for _ in range(S)]
    status = [np.zeros(NB, np.int32) for _ in range(S)]
    p2c = [np.zeros((NB, cap), np.int32) for _ in range(S)] if match else None
    matched = [0]

    def submit(s, frames):
        if full:
            ctx.readimage_submit(s, frames, seeds[:len(frames)], smoothed=True, frame_count0=1)
        elif match:
            ctx.linefront_submit(s, frames, smoothed=True)
        else:
            ctx.edlines_submit(s, frames, smoothed=True)

    def collect(s):
        if full:
            ctx.readimage_collect_into(s, lines[s], counts[s], cap, p2c[s], vps[s], vp_idx[s], vp_st[s])
            matched[0] = int((p2c[s][1:] >= 0).sum())
        elif match:
            ctx.linefront_collect_into(s, lines[s], counts[s], cap, p2c[s])
            matched[0] = int((p2c[s][1:] >= 0).sum())
        else:
            ctx.edlines_collect_into(s, lines[s], counts[s], cap, status[s])

    def resident(s):
        if full:
            ctx.readimage_run_resident(s)
        elif match:
            ctx.linefront_run_resident(s)
        else:
            ctx.edlines_run_resident(s)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_steps(n_steps):
        pending = []
        for i in range(n_steps):
            s = i % S
            if len(pending) == S:
                ps = pending.pop(0)
                collect(ps)
            submit(s, host_buf)
            pending.append(s)
        while pending:
            ps = pending.pop(0)
            collect(ps)
        return int(counts[ps][NB - B:].sum())

    e2e_steps(max(args.warmup, S))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    n_lines = e2e_steps(args.steps)
    d2h_bytes = ctx.last_d2h_bytes((args.steps - 1) % S)
    ctx.sync()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    ctx.edlines_submit(0, host_buf[NB - B:NB - B + 1], smoothed=True)  # one frame's edge pixels, for the byte counts below
    ctx.edlines_collect_into(0, lines[0], counts[0], cap, status[0])
    xy0, _ = ctx.edge_chains(0, W, H)
    submit(0, host_buf)       # leave a full batch resident in slot 0 again
    collect(0)
    anchors = int(len(ctx.linematch_points(0)["status"])) if match else 0

    for w in range(max(args.warmup, S)):
        resident(w % S)
    ctx.sync()
    ctx.reset_stage_times()
    l0 = ctx.kernel_launches()
    barrier()
    ev0.record()
    for i in range(args.steps):
        resident(i % S)
    ctx.sync()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - l0
    stage = ctx.stage_times()
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln[0])
    frames_total = B * args.steps * world
    value = frames_total / (dev_ms * 1e-3)
    e2e_value = frames_total / (e2e_ms * 1e-3)

    # roofline of the dominant EDLines kernel; algorithmic bytes per frame:
    P = W * H
    edge_px = float(len(xy0))  # frame 0 of the last batch (frames are a tiling of `unique`)
    alg = {"ed_grad": 7.0 * P,                                    # Sobel + map + anchors, one pass: r P, w 4P + 2P
           "ed_walk": 20.0 * edge_px,                              # 3 u16 reads + mark + record, re-pack r+w
           "ed_fit": 20.0 * edge_px}                               # chain pixel 3 x 4 B + Sobel pair 2 x 4 B
    launches_of = {"ed_grad": 1, "ed_walk": 1, "ed_fit": 2}
    kernels_of = {"ed_grad": "ed_grad_anchor_kernel", "ed_walk": "ed_walk_kernel", "ed_fit": "ed_fit_kernel+ed_compact_kernel"}
    if match:
        # padded pyramid (4 levels, 13-px border) ~ 1.5 P bytes: written once, read by the next level and by
        # the Scharr pass, which writes 4 B per padded pixel: ~ 1 + 3 x 1.5 + 6 = 11.5 P;
        # tracker: per anchor, level and iteration one 14x14 u8 window (196 B) + once per level the
        # 14x14 Scharr window (784 B) and the reference window: ~ 4 levels x (980 + ~8 x 196) B
        alg.update({"lm_pyramid": 11.5 * P, "lm_track": anchors * 4 * (980.0 + 8 * 196.0), "lm_vote": anchors * 24.0})
        launches_of.update({"lm_pyramid": 8, "lm_track": 6, "lm_vote": 1})
        kernels_of.update({"lm_pyramid": "klt_level0/pyrdown/scharr kernels", "lm_track": "klt_track_kernel x4 levels (+ anchors, offsets)",
                           "lm_vote": "lm_vote_kernel"})
    if full:
        cm = counts[0][NB - B:].astype(np.float64)
        Lm, Np, G = float(cm.mean()), float((cm * (cm - 1) / 2).mean()), 90 * 360 * 8.0
        alg.update({"preproc": 13.0 * P,  # remap: maps 8P + r P + w P; CLAHE: r 2P + w P
                    "vp_prep": 56.0 * Lm + G, "vp_vote": 16.0 * Np + 2 * G, "vp_score": 105 * 360 * 3 * 8.0, "vp_classify": 52.0 * Lm})
        launches_of.update({"preproc": 3, "vp_prep": 1, "vp_vote": 2, "vp_score": 1, "vp_classify": 1})
        kernels_of.update({"preproc": "remap_kernel + clahe_lut_kernel + clahe_apply_kernel", "vp_prep": "vp_prepare_kernel",
                           "vp_vote": "vp_vote_kernel + vp_smooth_kernel", "vp_score": "vp_score_kernel", "vp_classify": "vp_classify_kernel"})
    ed = {k: stage[k] for k in alg}
    dom = max(ed, key=lambda k: ed[k][0])
    peak, peak_kind = measured_peak()
    launches_in_stage = launches_of[dom]
    dur = ed[dom][0] / max(ed[dom][1] / launches_in_stage, 1)
    achieved = alg[dom] * B / (dur * 1e-3) / 1e9 if dur > 0 else 0.0
    tot = max(sum(x[0] for x in stage.values()), 1e-9)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "klt_track_traffic.json")
    if dom == "lm_track" and os.path.exists(tp) and wl["frames"] == "C2":
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (4 level launches), per frame
            traffic = float(json.load(open(tp))["dram_bytes_per_frame"]) * B
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernels_of[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)", "ms_per_launch": dur,
                "algorithmic_bytes_per_launch": alg[dom] * B,
                "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in stage.items() if v[1]},
                "stage_share": {k: round(v[0] / tot, 4) for k, v in stage.items() if v[1]}}
    line = {"metric": RI_METRIC if full else LF_METRIC if match else ED_METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/s16/f32/f64" if match else "u8/s16/f64", "data": "synthetic" if wl["frames"] == "C2" else "reference frames",
            "config": {"workload": wl["name"], "frames_per_step": B, "width": W, "height": H, "slots": S, "max_lines": cap,
                       "unique_frames": len(unique), "parallelism": f"frames x{world}", "smoothed": True,
                       "edline_param": [param.ksize, param.sigma, param.gradientThreshold, param.anchorThreshold,
                                        param.scanIntervals, param.minLineLen, param.lineFitErrThreshold],
                       "lines_per_frame": round(n_lines / B, 1), "edge_px_frame0": int(edge_px),
                       "matched_lines_per_frame": round(matched[0] / B, 1) if match else None,
                       "anchors_pair0": anchors if match else None,
                       "overlap": "each step submits B+1 frames (the previous step's last frame first)" if match else None,
                       "pipeline": ("remap (EuRoC cam0 maps) + CLAHE(3.0, 8x8) -> EDLines -> Matching(f-1, f) -> vanishing points on each "
                                    "frame's own lines (lines == all_lines), one pass over the device; labelled lines/frame %.1f"
                                    % ((vp_idx[0][NB - B:] != 3)[np.arange(cap)[None, :] < counts[0][NB - B:, None]].sum() / B)) if full else None,
                       "l2": "inputs per step (%.0f MB) exceed the 126 MB L2" % (B * W * H / 1e6)},
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": NB * W * H, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = (readimage_cpu_baseline(unique, name=wl["name"]) if full
                                    else ed_cpu_baseline(unique, name=wl["name"], match=match))
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def euroc_maps(w=752, h=480):
    """Radial-tangential undistortion maps of EuRoC cam0 (config/euroc/euroc_config.yaml), what initUndistortRectifyMap gives."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    k1, k2, p1, p2 = -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05
    x = (xx - cx) / fx; y = (yy - cy) / fy; r2 = x * x + y * y; rad = 1 + k1 * r2 + k2 * r2 * r2
    mapx = ((x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)) * fx + cx).astype(np.float32)
    mapy = ((y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y) * fy + cy).astype(np.float32)
    return mapx, mapy


def readimage_cpu_baseline(unique, seconds=8.0, name="R1"):
    """The same four stages on the host cores, each with the reference's own code where it exists: cv2.remap + cv2 CLAHE (what
    readImage calls; the oracle's bit-identical port if cv2 is missing), the reference's EDLines + line matching
    (oracle/_ref/libref_linefront.so, threads), the reference's vanishing-point stage (oracle/_ref/libref_vp.so, processes).
    The stages run one after the other over the same n frames; value = n / the sum of their times."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    use_lf = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_linefront.so"))
    use_vp = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_vp.so"))
    mapx, mapy = euroc_maps()
    try:
        import cv2
        cv2.setNumThreads(1)
        clahe_of = lambda: cv2.createCLAHE(3.0, (8, 8))
        pre_one = lambda f: clahe_of().apply(cv2.remap(f, mapx, mapy, cv2.INTER_LINEAR))
        pre_kind = "cv2.remap + cv2.createCLAHE(3.0, 8x8)"
    except Exception:
        pre_one = lambda f: O.clahe(O.remap_linear(f, mapx, mapy), 3.0, 8)
        pre_kind = "oracle port of remap + CLAHE"
    t = time.time()
    pre_u = [pre_one(f) for f in unique[:4]]
    O.linefront_sequence(np.ascontiguousarray(pre_u), threads=1, use_ref=use_lf)
    per_frame = max((time.time() - t) / 4, 1e-3) * 1.3
    n = int(max(threads * 4, min(seconds / per_frame * threads, 8192)))
    frames = tile_frames(unique, n)
    t0 = time.time()
    with ThreadPoolExecutor(threads) as ex:
        pre = np.ascontiguousarray(list(ex.map(pre_one, frames)))
    t_pre = time.time() - t0
    t0 = time.time()
    matched = O.linefront_sequence(pre, threads=threads, use_ref=use_lf)
    t_lf = time.time() - t0
    # line sets of the distinct pre-processed frames (the detector's output is not returned by the timing entry point)
    sets = [O.edline_detect(f) for f in pre[:len(unique)]]
    cap = max(len(x) for x in sets) + 1
    u_lines = np.zeros((len(sets), cap), O.LINE_DTYPE); u_counts = np.array([len(x) for x in sets], np.int32)
    seeds = np.zeros(len(sets), np.uint32)
    for i, x in enumerate(sets):
        u_lines[i, :len(x)] = x
        for k in range(64):  # a seed on which the reference's code does not read lx[] out of range
            sd = 1700000000 + 1009 * k + i
            if len(x) > 2 and O.vp_detect(x, None, *EUROC_CAM, sd, 1, math_mode=0, details=True)[2]["flags"] == 0:
                seeds[i] = sd
                break
    keep = seeds != 0
    u_lines, u_counts, seeds = u_lines[keep], u_counts[keep], seeds[keep]
    reps = (n + len(u_counts) - 1) // len(u_counts)
    _, t_vp = vp_cpu_run(np.ascontiguousarray(np.concatenate([u_lines] * reps)[:n]), np.ascontiguousarray(np.concatenate([u_counts] * reps)[:n]),
                         np.ascontiguousarray(np.concatenate([seeds] * reps)[:n]), threads, use_vp)
    tot = t_pre + t_lf + t_vp
    return {"value": n / tot, "unit": "frames/s", "cores": threads, "kind": "reference" if (use_lf and use_vp) else "port",
            "sample": f"{n} frames of {name}: {pre_kind} {t_pre:.1f}s ({threads} threads) + the reference's own EDLines + line matching "
                      f"{t_lf:.1f}s ({threads} threads, {matched} matched lines) + the reference's own vanishing-point stage {t_vp:.1f}s "
                      f"({threads} processes), one stage after the other",
            "stage_seconds": {"preprocess": t_pre, "edlines_matching": t_lf, "vanishing_points": t_vp}}


def _vp_chunk(job):
    """One host process: frames one after another through the reference's own vanishing_point_detection.cpp
    (oracle/_ref/libref_vp.so) or, where that library did not travel, the oracle port (libm arithmetic)."""
    lines, counts, seeds, use_ref = job
    from oracle import oracle as O
    _, _, tot = O.vp_sequence(lines, counts, seeds, *EUROC_CAM, frame_count0=1, math_mode=0, use_ref=use_ref)
    return tot


def vp_cpu_run(lines, counts, seeds, procs, use_ref):
    """rand()/srand() are process-global in the reference's code, so the host cores are used as processes."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    n = len(counts)
    cuts = [n * i // procs for i in range(procs + 1)]
    jobs = [(lines[a:b], counts[a:b], seeds[a:b], use_ref) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    # an executor (not a Pool): a worker that dies raises BrokenProcessPool instead of hanging the bench
    with ProcessPoolExecutor(len(jobs), mp_context=mp.get_context("fork")) as pool:
        list(pool.map(int, [0] * len(jobs)))  # workers started before the clock
        t = time.time()
        tot = sum(pool.map(_vp_chunk, jobs, timeout=600))
        dt = time.time() - t
    return tot, dt


def vp_cpu_baseline(u_lines, u_counts, u_seeds, seconds=10.0, name="V1"):
    from oracle import oracle as O
    O.build()
    use_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_vp.so"))
    procs = os.cpu_count() or 1
    per_frame = max(vp_cpu_run(u_lines[:4], u_counts[:4], u_seeds[:4], 1, use_ref)[1] / 4, 1e-4)
    n = int(max(procs * 4, min(seconds / per_frame * procs, 65536)))
    reps = (n + len(u_counts) - 1) // len(u_counts)
    lines = np.ascontiguousarray(np.concatenate([u_lines] * reps)[:n])
    counts = np.ascontiguousarray(np.concatenate([u_counts] * reps)[:n]); seeds = np.ascontiguousarray(np.concatenate([u_seeds] * reps)[:n])
    tot, dt = vp_cpu_run(lines, counts, seeds, procs, use_ref)
    what = ("the reference's own feature_tracker/src/vanishing_point_detection.cpp compiled against oracle/cvshim, its "
            "time(NULL) answered with the frame's seed" if use_ref else "CPU oracle port (oracle/orc_vp.c, libm arithmetic)")
    return {"value": n / dt, "unit": "frames/s", "cores": procs, "kind": "reference" if use_ref else "port",
            "sample": f"{n} line sets of {name} ({tot} labelled lines) in {dt:.1f}s; {what}; one object per process, "
                      f"{procs} processes over contiguous chunks (rand() is process-global)",
            "single_thread_frames_per_s": 1.0 / per_frame}


def vp_inputs(ctx, capi, unique, cap):
    """Line sets = the device's own EDLines of the distinct frames; one seed per set on which the reference does not
    read lx[] out of range (device status 0; every frame of the bench is a later call of its object, frame_count > 0), so
    that its code can be timed."""
    ctx.edlines_configure(capi.EDLineParam())
    sets = []
    for a in range(0, len(unique), 32):
        sets += ctx.edlines_detect_batch(unique[a:a + 32], smoothed=True)
    n = len(sets)
    seeds = np.zeros(n, np.uint32)
    todo = list(range(n))
    base = 1700000000
    for attempt in range(64):
        if not todo:
            break
        cand = np.array([base + 1009 * attempt + i for i in todo], np.uint32)
        st = np.concatenate([ctx.vp_detect_batch([sets[i] for i in todo[a:a + 32]], cand[a:a + 32], frame_count0=1)[2]
                             for a in range(0, len(todo), 32)])
        ok = st == 0
        for j, i in enumerate(todo):
            if ok[j]:
                seeds[i] = cand[j]
        todo = [i for j, i in enumerate(todo) if not ok[j]]
    keep = [i for i in range(n) if seeds[i] and len(sets[i]) > 2]
    u_lines = np.zeros((len(keep), cap), capi.LINE_DTYPE)
    u_counts = np.array([len(sets[i]) for i in keep], np.int32)
    for r, i in enumerate(keep):
        u_lines[r, :u_counts[r]] = sets[i]
    return u_lines, u_counts, seeds[keep]


def run_reference_vp(args):
    """--impl reference --workload V1: needs the line sets, which bench.py takes from the device's EDLines; on a box
    without a GPU the oracle's EDLines (bit-identical) stands in."""
    from oracle import oracle as O
    O.build()
    wl = WORKLOADS[args.workload]
    unique = make_frames(min(args.unique, 32), args.seed, args.workload)
    cap = args.max_lines or wl["max_lines"]
    sets = [O.edline_detect(f) for f in unique]
    u_lines = np.zeros((len(sets), cap), O.LINE_DTYPE); u_counts = np.array([len(s) for s in sets], np.int32)
    seeds = np.zeros(len(sets), np.uint32)
    for i, s in enumerate(sets):
        u_lines[i, :len(s)] = s
        for k in range(64):  # a seed on which the reference's code does not read lx[] out of range
            sd = 1700000000 + 1009 * k + i
            if O.vp_detect(s, None, *EUROC_CAM, sd, 1, math_mode=0, details=True)[2]["flags"] == 0:
                seeds[i] = sd
                break
    keep = (seeds != 0) & (u_counts > 2)
    u_lines, u_counts, seeds = u_lines[keep], u_counts[keep], seeds[keep]
    use_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_vp.so"))
    procs = os.cpu_count() or 1
    per_frame = max(vp_cpu_run(u_lines[:4], u_counts[:4], seeds[:4], 1, use_ref)[1] / 4, 1e-4)
    n = int(max(procs * 4, min(5.0 / per_frame * procs, 65536)))
    reps = (n + len(u_counts) - 1) // len(u_counts)
    lines = np.ascontiguousarray(np.concatenate([u_lines] * reps)[:n])
    counts = np.ascontiguousarray(np.concatenate([u_counts] * reps)[:n]); sd = np.ascontiguousarray(np.concatenate([seeds] * reps)[:n])
    for _ in range(args.warmup):
        vp_cpu_run(lines[:max(procs, n // 4)], counts[:max(procs, n // 4)], sd[:max(procs, n // 4)], procs, use_ref)
    dt = 0.0
    for _ in range(args.steps):
        dt += vp_cpu_run(lines, counts, sd, procs, use_ref)[1]
    fps = n * args.steps / dt
    kind = "reference" if use_ref else "port"
    line = {"impl": "reference", "metric": VP_METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_step": n,
                       "note": "the reference's own vanishing_point_detection.cpp compiled against oracle/cvshim" if use_ref
                       else "CPU oracle port (oracle/_ref not present)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": procs, "kind": kind, "sample": f"{n} line sets/step x {args.steps} steps"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_vp(args, torch, dist, rank, local_rank, world):
    """--workload V1 / V1r: run_vanishing_point_detection on B line sets per step (lines == all_lines, what readImage
    passes unless it found > 2 vertical lines)."""
    vpl = importlib.import_module("vplines_slam_b200")
    capi = vpl.capi
    wl = WORKLOADS[args.workload]
    W, H = wl["w"], wl["h"]
    B, S, cap = args.batch, args.slots, (args.max_lines or wl["max_lines"])
    unique = make_frames(args.unique, args.seed, args.workload)
    ctx = capi.Context(device=local_rank, max_width=W, max_height=H, max_octaves=1, max_lines=cap, max_batch=B,
                       num_slots=S, blur_first=True, profile=True, lsd_path=False)
    ctx.vp_configure(*EUROC_CAM)
    u_lines, u_counts, u_seeds = vp_inputs(ctx, capi, unique, cap)
    reps = (B + len(u_counts) - 1) // len(u_counts)
    lines = np.ascontiguousarray(np.concatenate([u_lines] * reps)[:B])
    counts = np.ascontiguousarray(np.concatenate([u_counts] * reps)[:B])
    seeds = np.ascontiguousarray(np.concatenate([u_seeds] * reps)[:B])
    vps = [np.zeros((B, 3, 3), np.float64) for _ in range(S)]
    idx = [np.zeros((B, cap), np.int32) for _ in range(S)]
    status = [np.zeros(B, np.int32) for _ in range(S)]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_steps(n_steps):
        pending = []
        for i in range(n_steps):
            s = i % S
            if len(pending) == S:
                ps = pending.pop(0)
                ctx.vp_collect_into(ps, cap, vps[ps], idx[ps], status[ps])
            ctx.vp_submit(s, lines, counts, seeds, frame_count0=1)
            pending.append(s)
        while pending:
            ps = pending.pop(0)
            ctx.vp_collect_into(ps, cap, vps[ps], idx[ps], status[ps])
        return ps

    e2e_steps(max(args.warmup, S))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    ps = e2e_steps(args.steps)
    ctx.sync()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    labelled = int(sum((idx[ps][i, :counts[i]] != 3).sum() for i in range(B)))
    assert (status[ps] == 0).all()

    for w in range(max(args.warmup, S)):
        ctx.vp_run_resident(w % S)
    ctx.sync()
    ctx.reset_stage_times()
    l0 = ctx.kernel_launches()
    barrier()
    if args.profile_region:  # ncu --profile-from-start off captures only the resident steps
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for i in range(args.steps):
        ctx.vp_run_resident(i % S)
    ctx.sync()
    ev1.record()
    if args.profile_region:
        torch.cuda.cudart().cudaProfilerStop()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - l0
    stage = ctx.stage_times()
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln[0])
    frames_total = B * args.steps * world
    value = frames_total / (dev_ms * 1e-3)
    e2e_value = frames_total / (e2e_ms * 1e-3)

    # algorithmic bytes per frame: L lines, Np = L (L - 1) / 2 pairs, G = 90 x 360 cells of 8 B
    Lm = float(counts.mean()); Np = float((counts.astype(np.float64) * (counts - 1) / 2).mean()); G = 90 * 360 * 8.0
    alg = {"vp_prep": 16.0 * Lm + 40.0 * Lm + G,            # endpoints in, 5 doubles per line out, grid cleared
           "vp_vote": 16.0 * Np + G + G,                    # one read-modify-write per voting pair; 3x3 pass: r G, w G
           "vp_score": 105 * 360 * 3 * 8.0,                 # three cell reads per hypothesis
           "vp_classify": 16.0 * Lm + 32.0 * Lm + 4.0 * Lm}  # endpoints, angles, labels
    launches_of = {"vp_prep": 1, "vp_vote": 2, "vp_score": 1, "vp_classify": 1}
    kernels_of = {"vp_prep": "vp_prepare_kernel (+ grid memset)", "vp_vote": "vp_vote_kernel + vp_smooth_kernel",
                  "vp_score": "vp_score_kernel", "vp_classify": "vp_classify_kernel"}
    st = {k: stage[k] for k in alg}
    dom = max(st, key=lambda k: st[k][0])
    peak, peak_kind = measured_peak()
    dur = st[dom][0] / max(st[dom][1] / launches_of[dom], 1)
    achieved = alg[dom] * B / (dur * 1e-3) / 1e9 if dur > 0 else 0.0
    tot = max(sum(x[0] for x in stage.values()), 1e-9)
    roofline = {"bound": "hbm", "kernel": kernels_of[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "ms_per_launch": dur, "algorithmic_bytes_per_launch": alg[dom] * B,
                "note": "FP64-arithmetic-bound (double-double atan/acos/sincos per pair and per hypothesis), not a streaming kernel; see DESIGN.md",
                "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in stage.items() if v[1]},
                "stage_share": {k: round(v[0] / tot, 4) for k, v in stage.items() if v[1]}}
    h2d = int(B * cap * capi.LINE_DTYPE.itemsize + 2 * B * 4 + B * 4)
    d2h = int(B * 72 + B * cap * 4 + 2 * B * 4)
    line = {"metric": VP_METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if wl["frames"] == "C2" else "reference frames",
            "config": {"workload": wl["name"], "frames_per_step": B, "slots": S, "max_lines": cap, "unique_line_sets": int(len(u_counts)),
                       "lines_per_frame": round(Lm, 1), "pairs_per_frame": round(Np, 1), "hypotheses_per_frame": 105 * 360,
                       "labelled_lines_per_frame": round(labelled / B, 1), "camera": list(EUROC_CAM), "parallelism": f"frames x{world}",
                       "seeds": "one per line set, chosen so that the reference's own code does not read lx[] out of range (device status 0)",
                       "l2": "sphere grids per step (%.0f MB) exceed the 126 MB L2" % (2 * B * G / 1e6)},
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks}
    if rank == 0:
        line["cpu_baseline"] = (vp_cpu_baseline(u_lines, u_counts, u_seeds, name=wl["name"])
                                if (world == 1 and not args.no_cpu_baseline) else None)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_matcher(args, torch, dist, rank, local_rank, world):
    """--workload M4: BinaryDescriptorMatcher::knnMatch alone on B frame pairs of 2000 x 2000 codes (the matcher of
    BASELINE.json configs[3]); reports the POPC-pipe utilisation SURVEY 8d asks for next to the contract's roofline."""
    vpl = importlib.import_module("vplines_slam_b200")
    capi = vpl.capi
    wl = WORKLOADS[args.workload]
    Lq, Lt, K = wl["lq"], wl["lt"], wl["k"]
    B = min(args.batch, 512)
    rng = np.random.default_rng(args.seed + rank)
    q = rng.integers(0, 256, (B, Lq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (B, Lt, 32), dtype=np.uint8)
    t[:, ::7] = q[:, ::7]  # planted exact matches (distance 0) and, with them, ties between candidates
    ctx = capi.Context(device=local_rank, max_width=wl["w"], max_height=wl["h"], max_octaves=1, max_lines=max(Lq, Lt),
                       max_batch=B, num_slots=1, blur_first=True, profile=True)
    nq, nt = np.full(B, Lq, np.int32), np.full(B, Lt, np.int32)
    res = np.zeros((B, Lq, K), capi.DMATCH_DTYPE)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        ctx.match_batch_into(q, nq, t, nt, K, res)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ctx.match_batch_into(q, nq, t, nt, K, res)
    ctx.sync()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    assert (res["trainIdx"][:, ::7, 0] == np.arange(0, Lq, 7)).all() and (res["distance"][:, ::7, 0] == 0).all()
    for _ in range(max(args.warmup, 1)):
        ctx.match_run_resident(K)
    ctx.sync()
    ctx.reset_stage_times()
    l0 = ctx.kernel_launches()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ctx.match_run_resident(K)
    ctx.sync()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - l0
    stage = ctx.stage_times()
    clocks = sampler.stop()
    popc_peak = ctx.popc_peak()
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(tt[0]), float(tt[1])
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln[0])
    total = B * args.steps * world
    value = total / (dev_ms * 1e-3)
    e2e_value = total / (e2e_ms * 1e-3)
    dur = stage["match"][0] / max(stage["match"][1], 1)
    alg = (32.0 * (Lq + Lt) + 16.0 * K * Lq) * B        # codes in, DMatch out
    popc = 8.0 * Lq * Lt * B                               # popc32 per launch
    peak, peak_kind = measured_peak()
    achieved = alg / (dur * 1e-3) / 1e9 if dur > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "hamming_knn_kernel<2>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "ms_per_launch": dur, "algorithmic_bytes_per_launch": alg,
                "note": "integer-pipe (POPC) bound, not HBM: see popc",
                "popc": {"achieved_popc32_per_s": popc / (dur * 1e-3) if dur > 0 else 0.0, "peak_popc32_per_s": popc_peak,
                         "frac": (popc / (dur * 1e-3) / popc_peak) if dur > 0 and popc_peak > 0 else None,
                         "popc32_per_launch": popc,
                         "peak_source": "vpl_debug_popc_peak: micro-benchmark kernel on this device (8 independent popc chains per thread)"}}
    line = {"metric": M_METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (xor/popc)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "pairs_per_step": B, "queries": Lq, "train": Lt, "k": K, "parallelism": f"pairs x{world}",
                       "codes": "uniform random 256-bit codes, every 7th train code equal to its query (exact matches and ties)",
                       "l2": "descriptor sets per step (%.0f MB) fit the 126 MB L2: the kernel is compute-bound, its inputs are re-read from "
                             "shared-memory tiles" % (B * (Lq + Lt) * 32 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(B * (Lq + Lt) * 32 + 8 * B), "d2h_bytes_per_step": int(B * Lq * K * 16)},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks}
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            from concurrent.futures import ThreadPoolExecutor
            from oracle import oracle as O
            O.build()
            threads = os.cpu_count() or 1
            t0 = time.time(); O.hamming_knn(q[0], t[0], K); per = max(time.time() - t0, 1e-4)
            n = int(max(threads, min(10.0 / per * threads, B)))
            t0 = time.time()
            with ThreadPoolExecutor(threads) as ex:  # the C call releases the GIL
                list(ex.map(lambda i: O.hamming_knn(q[i % B], t[i % B], K), range(n)))
            dt = time.time() - t0
            cb = {"value": n / dt, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                  "sample": f"{n} pairs of {Lq}x{Lt} codes in {dt:.1f}s; oracle brute force (orc_hamming_knn, 64-bit popcount), "
                            f"{threads} threads", "single_thread_pairs_per_s": 1.0 / per}
        line["cpu_baseline"] = cb
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def sharded_config(args, wl, world):
    """`config` of the sharded-sequence workload (C5), the same for both arms."""
    B = min(args.batch, 2368)
    return {"workload": wl["name"], "total_frames": args.total_frames, "frames_per_batch": B, "width": wl["w"],
            "height": wl["h"], "octaves": wl["octaves"], "match_k": wl["k"], "unique_frames": min(args.unique, 48),
            "slots": args.slots, "max_lines": args.max_lines or wl["max_lines"],
            "parallelism": f"frame ranges x{world}, 1-frame halo, no collective on the data path",
            "l2": "every batch (%.0f MB of frames) exceeds the 126 MB L2" % (B * wl["w"] * wl["h"] / 1e6)}


def frame_hashes(counts, kl, desc, mt, n):
    """One 32-bit hash per frame of a collected step's dense outputs (KeyLines, descriptors, matches)."""
    import zlib
    off = np.concatenate([[0], np.cumsum(np.asarray(counts[:n], np.int64))])
    out = np.zeros(n, np.uint32)
    for f in range(n):
        a, b = int(off[f]), int(off[f + 1])
        h = zlib.crc32(kl[a:b].tobytes())
        h = zlib.crc32(desc[a:b].tobytes(), h)
        out[f] = zlib.crc32(mt[a:b].tobytes(), h) & 0xffffffff
    return out


def run_sharded_sequence(args, torch, dist, rank, local_rank, world):
    """--workload C5: ONE sequence of `total` frames (1280x720, 2 octaves, knnMatch k=2), strong scaling: rank g owns
    frames [g*total/G, (g+1)*total/G) and also processes frame start-1 (halo), so that every consecutive pair is
    matched on exactly one GPU; no data-path collective.  A step = every rank's whole shard, batch by batch, host
    buffers in and out through the C ABI.  After the timed region every rank hashes its per-frame results; rank 0
    gathers them (the only collective, outside the timed region) and compares them with the hashes of the same
    sequence run on ONE GPU (rank 0 alone, untimed): `identical_to_1gpu`."""
    vpl = importlib.import_module("vplines_slam_b200")
    capi = vpl.capi
    driver = importlib.import_module("vplines-slam_b200.driver")
    wl = WORKLOADS[args.workload]
    W, H, OCT, K = wl["w"], wl["h"], wl["octaves"], wl["k"]
    cap = args.max_lines or wl["max_lines"]
    total = args.total_frames
    B, S = min(args.batch, 2368), args.slots
    unique = make_frames(min(args.unique, 48), args.seed, args.workload)
    seq_len = len(unique)

    def seq_frames(lo, hi):  # frames [lo, hi) of the global sequence (a tiling of the distinct frames)
        idx = np.arange(lo, hi) % seq_len
        return np.ascontiguousarray(unique[idx])

    ctx = capi.Context(device=local_rank, max_width=W, max_height=H, max_octaves=OCT, max_lines=cap, max_batch=B + 1,
                       num_slots=S, blur_first=True, profile=False)
    rows = (B + 1) * cap
    kl = [np.zeros(rows, capi.KEYLINE_DTYPE) for _ in range(S)]
    counts = [np.zeros(B + 1, np.int32) for _ in range(S)]
    desc = [np.zeros((rows, 32), np.uint8) for _ in range(S)]
    mt = [np.zeros((rows, K), capi.DMATCH_DTYPE) for _ in range(S)]
    for a in kl + desc + mt:
        ctx.host_register(a)

    def run_shard(start, end, halo, shard, want_hashes):
        """frames [start, end) (+ halo frame in front) through the grouped pipeline of the batch API (uploads ahead,
        the slots' batches submitted as a group, group g submitted before group g-1 is collected); shard = pinned
        host array of them."""
        lo = start - halo
        n = end - lo
        hashes = np.zeros(end - start, np.uint32) if want_hashes else None
        lines = 0
        batches = []  # (offset in the shard, frames, leading halo frames to skip)
        i = 0
        while i < n:
            first = not batches
            b_n = min(B + (1 if first and halo else 0), n - i)
            batches.append((i, b_n, 1 if first and halo else 0))
            i += b_n
        groups = [list(range(g, min(g + S, len(batches)))) for g in range(0, len(batches), S)]

        def upload(bi):
            i0, b_n, _ = batches[bi]
            ctx.upload(bi % S, shard[i0:i0 + b_n])

        def submit(grp):
            ctx.submit_group([bi % S for bi in grp], [batches[bi][1] for bi in grp], W, H, scale=2, num_octaves=OCT, k=K,
                             chain=[bi > 0 for bi in grp])

        def collect(bi):
            nonlocal lines
            i0, b_n, skip = batches[bi]
            ps = bi % S
            ctx.collect_dense_into(ps, counts[ps], kl[ps], desc[ps], mt[ps])
            lines += int(counts[ps][skip:b_n].sum())
            if want_hashes:
                hh = frame_hashes(counts[ps], kl[ps], desc[ps], mt[ps], b_n)
                hashes[lo + i0 + skip - start:lo + i0 + b_n - start] = hh[skip:]

        for bi in groups[0]:
            upload(bi)
        submit(groups[0])
        for gi in range(1, len(groups)):
            for bi in groups[gi]:
                upload(bi)
            submit(groups[gi])
            for bi in groups[gi - 1]:
                collect(bi)
        for bi in groups[-1]:
            collect(bi)
        return lines, hashes

    start, end, halo = driver.shard_range(total, rank, world)
    shard = seq_frames(start - halo, end)
    ctx.host_register(shard)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        run_shard(start, end, halo, shard, False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    lines = 0
    for _ in range(args.steps):
        lines, _ = run_shard(start, end, halo, shard, False)
    ctx.sync()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop()
    _, mine = run_shard(start, end, halo, shard, True)  # untimed: per-frame hashes of this rank's shard

    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
        ln = torch.tensor([launches, lines], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches, lines = int(ln[0]), int(ln[1])
        full = torch.zeros(total, device="cuda", dtype=torch.int64)
        full[start:end] = torch.from_numpy(mine.astype(np.int64)).cuda()
        dist.all_reduce(full, op=dist.ReduceOp.SUM)  # shards are disjoint: the sum is the concatenation
        gathered = full.cpu().numpy().astype(np.uint32)
    else:
        gathered = mine

    identical = None
    parity = None
    if rank == 0:
        # the same sequence on ONE GPU (this rank alone, its context, batches cut at other places)
        whole = seq_frames(0, total) if world > 1 else shard
        if world > 1:
            ctx.host_register(whole)
        _, ref = run_shard(0, total, 0, whole, True)
        identical = bool(np.array_equal(ref, gathered))
        if not args.no_parity:
            # and a few frames of it against the CPU oracle chain
            n0 = min(B, total)
            ctx.submit(0, whole[:n0], scale=2, num_octaves=OCT, k=K, chain=False)
            ctx.collect_dense_into(0, counts[0], kl[0], desc[0], mt[0])
            check = sorted({0, 1, n0 // 2, n0 - 1})
            bad = verify_dense_step(whole[:n0], None, counts[0], kl[0], desc[0], mt[0], K, OCT, check)
            parity = {"checked": len(bad) == 0, "frames": len(check), "mismatches": [list(map(str, b)) for b in bad][:8]}

    frames_total = total * args.steps
    e2e_value = frames_total / (e2e_ms * 1e-3)
    line = {"metric": lsd_metric(wl), "value": e2e_value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": e2e_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
            "config": sharded_config(args, wl, world),
            "lines_per_frame": round(lines / max(total, 1), 1),
            "note": "value = e2e: host buffers in and out through the C ABI (this workload has no resident variant)",
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": (total + world - 1) * W * H, "d2h_bytes_per_step": None},
            "gpu_launches": launches, "identical_to_1gpu": identical, "parity_checked": bool(parity and parity["checked"]),
            "parity": parity, "clocks": clocks, "cpu_baseline": None, "roofline": None}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores, nothing else."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload.startswith("V"):
        return run_reference_vp(args)
    if args.workload.startswith("R"):
        wl = WORKLOADS[args.workload]
        unique = make_frames(min(args.unique, 32), args.seed, args.workload)
        cb = None
        for _ in range(max(args.steps, 1)):  # every step is one bounded sample of the four stages
            cb = readimage_cpu_baseline(unique, seconds=5.0, name=wl["name"])
        line = {"impl": "reference", "metric": RI_METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/s16/f32/f64", "data": "synthetic" if wl["frames"] == "C2" else "reference frames",
                "config": {"workload": wl["name"], "note": cb["sample"]}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    if args.workload.startswith("E"):
        return run_reference_edlines(args)
    wl = WORKLOADS[args.workload]
    OCT, K = wl["octaves"], wl["k"]
    unique = make_frames(min(args.unique, 48) if wl.get("sharded") else args.unique, args.seed, args.workload)
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    t = time.time()
    O.frontend_sequence(unique[:4], num_octaves=OCT, max_lines=16384, threads=1)
    per_frame = max((time.time() - t) / 4, 1e-3)
    # bounded sample per step: ~6 s of wall time on the host cores
    n = int(max(threads * 2, min(6.0 / per_frame * threads, args.batch)))
    frames = tile_frames(unique, n)
    for _ in range(args.warmup):
        O.frontend_sequence(frames[:max(threads, n // 4)], num_octaves=OCT, max_lines=16384, threads=threads)
    t0 = time.time()
    total = 0
    for _ in range(args.steps):
        total = O.frontend_sequence(frames, num_octaves=OCT, max_lines=16384, threads=threads)
    dt = time.time() - t0
    fps = n * args.steps / dt
    line = {"impl": "reference", "metric": lsd_metric(wl),
            "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if wl.get("sharded") else "weak", "vs_baseline": None,
            "dtype": "u8/f32/f64", "data": "synthetic",
            "config": sharded_config(args, wl, args.gpus) if wl.get("sharded") else lsd_config(args, wl, args.gpus),
            "sample_frames_per_step": n, "lines_per_frame": round(total / max(n, 1), 1),
            "note": "CPU oracle port (oracle/, -O3 -march=native, all host threads); the reference's OpenCV-3.4 "
                    "line_descriptor binary is not buildable here (no OpenCV C++ / contrib); each step is a bounded "
                    f"sample of {n} frames of the workload",
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{n} frames/step x {args.steps} steps of {wl['name']}",
                             "single_thread_frames_per_s": 1.0 / per_frame},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per step (0: 4736 = 32 frames per SM for C2, so that the engine launches of the two slots "
                         "fill the 64 warp slots of every SM; 4736 / 2368 / 1184 for C1 / C3 / C4; 4096 for the other workloads)")
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--max-lines", type=int, default=0, help="KeyLine capacity per frame (0 = workload default)")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--unique", type=int, default=96, help="distinct synthetic frames generated (tiled to fill a batch)")
    ap.add_argument("--seed", type=int, default=20240601)
    ap.add_argument("--total-frames", type=int, default=8192, help="C5: length of the sharded sequence")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the last timed e2e step")
    ap.add_argument("--parity-frames", type=int, default=6, help="frames of the last e2e step compared with the oracle")
    ap.add_argument("--no-upload-ahead", dest="upload_ahead", action="store_false",
                    help="e2e: upload a slot's next batch only after its collect (round-1 behaviour)")
    ap.add_argument("--stress-upload-passes", type=int, default=0,
                    help="experiments: upload every batch this many extra times (a slow host link on one GPU)")
    ap.add_argument("--e2e-staggered", action="store_true",
                    help="e2e: collect and resubmit one slot at a time (round-1 behaviour) instead of submitting the "
                         "slots' batches back to back")
    ap.add_argument("--e2e-collect-first", action="store_true",
                    help="e2e: collect a group before the next one is submitted (one result generation per slot)")
    ap.add_argument("--timeline", action="store_true",
                    help="print where the stages of one group of resident batches lie in time on the device, and exit")
    ap.add_argument("--e2e-no-chain", action="store_true", help="experiments: no match across batch boundaries in the e2e pass")
    ap.add_argument("--e2e-only", action="store_true",
                    help="experiments: print only the end-to-end figure (host buffers in and out) and exit")
    ap.add_argument("--profile-region", action="store_true",
                    help="cudaProfilerStart/Stop around the resident steps (for ncu --profile-from-start off; V workloads)")
    args = ap.parse_args()
    if args.batch == 0:
        # frames per step for two slots in 180 GB (26 bytes of device state per pixel of every octave): 32 frames per SM
        # for the 752x480 workloads, 16 for 1280x720 with two octaves (= 32 engine warps per SM and slot), 8 for
        # 1920x1080; 4096 frames for the other workloads
        args.batch = {"C2": 4736, "C1": 4736, "C3": 2368, "C5": 2368, "C4": 1184}.get(args.workload, 4096)
    args.e2e_together = args.upload_ahead and not args.e2e_staggered
    if args.impl == "reference":
        return run_reference(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    if args.workload.startswith("V"):
        return run_vp(args, torch, dist, rank, local_rank, world)
    if args.workload.startswith("M"):
        return run_matcher(args, torch, dist, rank, local_rank, world)
    if args.workload.startswith("E") or args.workload.startswith("R"):
        return run_edlines(args, torch, dist, rank, local_rank, world)
    if WORKLOADS[args.workload].get("sharded"):
        return run_sharded_sequence(args, torch, dist, rank, local_rank, world)

    vpl = importlib.import_module("vplines_slam_b200")
    capi = vpl.capi
    wl = WORKLOADS[args.workload]
    W, H, OCT = wl["w"], wl["h"], wl["octaves"]
    B, S, cap, K = args.batch, args.slots, (args.max_lines or wl["max_lines"]), wl["k"]
    wl_name = wl["name"]

    # weak scaling: every rank owns `steps*batch` frames of the job's sequence.  Ranks hold DIFFERENT frames at every
    # position (rank r's part is the tiling of the `unique` frames rotated by r*unique//world) but the same multiset of
    # frames as the 1-GPU run, so that v_N / (N v_1) compares equal loads: frames drawn with another seed per rank carry
    # 209 to 278 lines per frame (oracle count over seeds + 7919 r), and the max over ranks then measures the content.
    # A rank's shard starts one frame early (halo = the last frame of rank r-1's part) so that the pair across the
    # shard boundary is matched exactly once, on rank r.
    def rank_frames(r):
        u = make_frames(args.unique, args.seed, args.workload)
        return np.roll(u, -((r * args.unique) // world), axis=0)

    unique = rank_frames(rank)
    halo = 1 if rank > 0 else 0
    halo_frame = unique[-1:]
    if rank > 0:
        prev = rank_frames(rank - 1)
        halo_frame = prev[(B - 1) % len(prev)][None]
    # one pinned host buffer: [halo frame | B frames]; the shard's first step starts at the halo
    host_buf = np.ascontiguousarray(np.concatenate([halo_frame, tile_frames(unique, B)]))
    batch_frames = host_buf[1:]

    ctx = capi.Context(device=local_rank, max_width=W, max_height=H, max_octaves=OCT, max_lines=cap,
                       max_batch=B + 1, num_slots=S, blur_first=True, profile=True)
    ctx.host_register(host_buf)  # frames live in pinned host memory: uploaded without a staging copy
    # host result buffers (dense: rows of all frames of a step packed back to back), pinned too
    rows = (B + 1) * cap
    kl = [np.zeros(rows, capi.KEYLINE_DTYPE) for _ in range(S)]
    counts = [np.zeros(B + 1, np.int32) for _ in range(S)]
    desc = [np.zeros((rows, 32), np.uint8) for _ in range(S)]
    mt = [np.zeros((rows, K), capi.DMATCH_DTYPE) for _ in range(S)]
    for a in kl + desc + mt:
        ctx.host_register(a)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_steps(n_steps, first_has_halo):
        """n_steps batches through (upload,) submit and collect, pipelined over the slots.  With upload-ahead (the
        default) every batch goes to the device through vpl_frontend_upload -- one copy stream for the context, so the
        batches arrive in order -- up to `slots` batches ahead of the one being submitted: the copy of batch i + slots
        overlaps the kernels of batch i instead of following the slot's collect."""
        pending = []
        lines = 0

        def frames_of(i):
            return host_buf if (i == 0 and first_has_halo) else batch_frames

        def upload(i):
            for _ in range(1 + args.stress_upload_passes):
                ctx.upload(i % S, frames_of(i))

        if args.upload_ahead:
            for i in range(min(S, n_steps)):
                upload(i)
        if args.e2e_together:
            # the slots' batches are submitted as a group (vpl_frontend_submit_group): every kernel of the path then
            # runs next to the SAME kernel of the other slot, and the engine launches start behind a barrier across the
            # group -- two of them side by side fill the SMs' warp slots, which is what the latency-bound engine needs
            # (52.1 k against 47.2 k frames/s for one slot at a time, profiles/r02_engine_wave_runs.txt).  Group g is
            # submitted BEFORE group g-1 is collected (it queues behind it on the slots' streams and writes the slots'
            # other result generation), so the download of g-1 and the host's turn-around overlap the kernels of g.
            groups = [range(g, min(g + S, n_steps)) for g in range(0, n_steps, S)]

            def submit(grp):
                ctx.submit_group([i % S for i in grp], [len(frames_of(i)) for i in grp], W, H, scale=2, num_octaves=OCT,
                                 k=K, chain=[(i > 0 and not args.e2e_no_chain) for i in grp])

            def collect(grp):
                for i in grp:
                    ctx.collect_dense_into(i % S, counts[i % S], kl[i % S], desc[i % S], mt[i % S])

            submit(groups[0])
            for gi in range(1, len(groups)):
                for i in groups[gi]:
                    upload(i)  # (the slot's second input buffer is free: group gi-2 has been collected)
                if args.e2e_collect_first:
                    collect(groups[gi - 1])
                    submit(groups[gi])
                else:
                    submit(groups[gi])
                    collect(groups[gi - 1])
            collect(groups[-1])
            lines = int(counts[(n_steps - 1) % S][:B].sum())
            return lines
        for i in range(n_steps):
            s = i % S
            if len(pending) == S:
                ps = pending.pop(0)
                ctx.collect_dense_into(ps, counts[ps], kl[ps], desc[ps], mt[ps])
            if args.upload_ahead:
                ctx.submit_uploaded(s, len(frames_of(i)), W, H, scale=2, num_octaves=OCT, k=K,
                                    chain=(i > 0 and not args.e2e_no_chain))
                if i + S < n_steps:
                    upload(i + S)
            else:
                ctx.submit(s, frames_of(i), scale=2, num_octaves=OCT, k=K, chain=(i > 0))
            pending.append(s)
        while pending:
            ps = pending.pop(0)
            ctx.collect_dense_into(ps, counts[ps], kl[ps], desc[ps], mt[ps])
            lines = int(counts[ps][:B].sum())
        return lines

    # ---- warm-up (also leaves a batch resident in every slot)
    # (per-stage events off in the timed e2e pass: with them on, a submit waits for the slot's previous batch)
    ctx.set_profile(False)
    e2e_steps(max(args.warmup, S), False)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- e2e: host buffers in, host buffers out
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    lines_last = e2e_steps(args.steps, halo == 1)
    d2h_bytes = ctx.last_d2h_bytes((args.steps - 1) % S)
    ctx.sync()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)

    if args.timeline:
        sampler.stop()
        ctx.set_profile(True)
        for rep in range(2):
            ctx.sync()
            ctx.mark()
            if args.e2e_together:
                ctx.run_resident_group(list(range(S)), k=K)
            else:
                for sl in range(S):
                    ctx.run_resident(sl, k=K)
            tl = {sl: ctx.timeline(sl) for sl in range(S)}
        if rank == 0:
            print(json.dumps({"timeline_ms": {str(sl): {k: [round(a, 2), round(b, 2)] for k, (a, b) in tl[sl].items()}
                                              for sl in tl}, "frames_per_step": B, "slots": S}))
        if dist is not None:
            dist.destroy_process_group()
        return

    if args.e2e_only:
        sampler.stop()
        per_rank = e2e_ms
        if dist is not None:
            t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [round(float(x[0]) / args.steps, 2) for x in allt]
            e2e_ms = max(float(x[0]) for x in allt)
        if rank == 0:
            print(json.dumps({"e2e_only": True, "n_gpus": world, "slots": S, "frames_per_step": B, "steps": args.steps,
                              "e2e_frames_per_s": B * args.steps * world / (e2e_ms * 1e-3),
                              "upload_ahead": bool(args.upload_ahead),
                              "ms_per_step": e2e_ms / args.steps, "ms_per_step_by_rank": per_rank}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- parity of what the timed e2e region just produced (outside the timed regions, rank 0): frames of the last
    # collected step against the CPU oracle chain, bit for bit
    parity = None
    if rank == 0 and not args.no_parity:
        last_slot = (args.steps - 1) % S
        last_frames = host_buf if (args.steps == 1 and halo == 1) else batch_frames
        prev = batch_frames[-1] if args.steps > 1 else None
        nb = len(last_frames)
        check = sorted({0, 1, nb // 2, nb - 1} | set(range(2, 2 + max(0, args.parity_frames - 4))))
        bad = verify_dense_step(last_frames, prev, counts[last_slot], kl[last_slot], desc[last_slot], mt[last_slot],
                                K, OCT, [f for f in check if f < nb])
        parity = {"checked": len(bad) == 0, "frames": len(check), "mismatches": [list(map(str, b)) for b in bad][:8],
                  "what": "KeyLines (17 fields), 32-byte descriptors, match indices + distances of the last timed e2e "
                          "step vs the CPU oracle chain, bit-exact"}

    # ---- stage times: one pass of resident batches with the per-stage events on (not the headline: with the
    # events on, a submit waits for the slot's previous batch)
    def resident_steps(n):
        if args.e2e_together:  # the same grouping as the e2e pass
            for g in range(0, n, S):
                ctx.run_resident_group([i % S for i in range(g, min(g + S, n))], k=K)
        else:
            for i in range(n):
                ctx.run_resident(i % S, k=K)

    ctx.set_profile(True)
    resident_steps(max(args.warmup, S))
    ctx.sync()
    ctx.reset_stage_times()
    resident_steps(args.steps)
    ctx.sync()
    stage = ctx.stage_times()
    ctx.set_profile(False)

    # ---- value: batches resident in HBM, kernels only, stage events off
    for w in range(S):
        ctx.run_resident(w % S, k=K)
    ctx.sync()
    l0 = ctx.kernel_launches()
    barrier()
    ev0.record()
    resident_steps(args.steps)  # grouped like the e2e pass: engine launches of the slots behind a barrier
    ctx.sync()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop()

    # ---- latency of ONE frame through the same C ABI (host buffer in, host results out), outside the timed regions
    latency = None
    if rank == 0:
        ctx.set_profile(True)
        one = np.ascontiguousarray(batch_frames[:1])
        for _ in range(2):
            ctx.frontend_batch(one, scale=2, num_octaves=OCT, k=0)
        ctx.reset_stage_times()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            ctx.frontend_batch(one, scale=2, num_octaves=OCT, k=0)
            ts.append((time.perf_counter() - t0) * 1e3)
        st1 = ctx.stage_times()
        latency = {"frames": 1, "e2e_ms_median": float(np.median(ts)), "region_engine_ms": st1["region"][0] / 5.0,
                   "note": "vpl_frontend_batch on one frame: upload, every kernel, download; median of 5 calls"}
        ctx.set_profile(False)

    # max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln[0])

    frames_total = B * args.steps * world
    value = frames_total / (dev_ms * 1e-3)
    e2e_value = frames_total / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (region engine), rank 0's launches
    peak, peak_kind = measured_peak()
    S_px = 0
    for o in range(OCT):
        S_px += int(round((W >> o) * 0.8)) * int(round((H >> o) * 0.8))
    # DESIGN.md "algorithmic bytes": region engine = 8 B per scaled pixel per frame
    # (grow: 4 B angle read + 1 B used read + 1 B used write per pixel = 6; rectangle moments: 2)
    eng_ms, eng_n = stage["region"]
    eng_bytes = 8.0 * S_px * B
    eng_dur = eng_ms / max(eng_n, 1)
    achieved = eng_bytes / (eng_dur * 1e-3) / 1e9 if eng_dur > 0 else 0.0
    stage_share = {k: round(v[0] / max(sum(x[0] for x in stage.values()), 1e-9), 4) for k, v in stage.items()}
    traffic = None
    issue = None
    tp = os.path.join(ROOT, "profiles", "engine_traffic.json")
    if os.path.exists(tp) and args.workload == "C2":
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per frame
            et = json.load(open(tp))
            traffic = float(et["dram_bytes_per_frame"]) * B
            # supplementary yardstick for this issue-bound kernel: warp instructions per frame (same ncu capture) x
            # frames / the launch duration measured live, against 148 SMs x 4 schedulers x the SM clock under load
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            ach = float(et["warp_inst_per_frame"]) * B / (eng_dur * 1e-3) if eng_dur > 0 else 0.0
            issue = {"achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": 148 * 4 * sm_hz,
                     "frac": ach / (148 * 4 * sm_hz), "warp_inst_per_frame": float(et["warp_inst_per_frame"]),
                     "source": "profiles/engine_traffic.json (ncu smsp__inst_executed.sum per frame) x live launch time"}
        except Exception:
            traffic = None
    conc = S if (args.e2e_together and args.steps >= S) else 1
    if issue:
        issue["frac_all_concurrent_launches"] = conc * issue["frac"]
    roofline = {"bound": "hbm", "kernel": "region_engine_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "concurrent_launches": conc,
                "frac_all_concurrent_launches": conc * achieved / peak,
                "concurrency_note": "achieved / frac are per launch as the contract asks; the launches of a group's slots "
                                    "run side by side behind a barrier (DESIGN.md 5a), so the device moves "
                                    "concurrent_launches x that in the same time",
                "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum per frame of the committed ncu --set full "
                                  "capture (profiles/engine_traffic.json) x frames per launch; not measured in this run",
                "ms_per_launch": eng_dur, "algorithmic_bytes_per_launch": eng_bytes,
                "note": "latency-bound by the sequential seed/FIFO order of LSD region growing; see DESIGN.md",
                "issue_rate": issue,
                "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in stage.items()},
                "stage_share": stage_share}

    line = {"metric": lsd_metric(wl), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
            "config": lsd_config(args, wl, world), "lines_per_frame": round(lines_last / B, 1),
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": B * W * H,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "latency": latency}
    nfa_chk = ctx.nfa_stats()
    if nfa_chk[0]:  # a -DVPL_NFA_CHECK build: the float32 early-exit test of nfa() against the double sequence
        line["nfa_check"] = {"decisions": nfa_chk[0], "answered_in_float32": nfa_chk[1], "disagreements": nfa_chk[2]}
    if rank == 0:
        line["parity_checked"] = bool(parity and parity["checked"])
        line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(unique, octaves=OCT, name=wl_name)
            line["cpu_baseline"]["cv2_crosscheck"] = cv2_crosscheck(unique, octaves=OCT, k=K)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
