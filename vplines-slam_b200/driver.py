"""Batch driver of the line front end: streams a frame sequence through a Context in
pipelined batches (submit on slot s while slot s-1 computes), and shards a sequence
across GPUs by contiguous frame range with a one-frame halo.

It plays the part of LineFeatureTracker::readImage's detect+match calls
(/root/reference/feature_tracker/src/line_feature_tracker.cpp:84-126) for a whole
sequence: per frame it yields the KeyLines, their descriptors and the match of every
line against the previous frame.  No collective is needed: only results are gathered.
"""
import numpy as np

from . import capi


def shard_range(n_frames, rank, world):
    """Frames [start, end) owned by `rank`; `halo` = 1 if frame start-1 must also be
    described here so that the pair (start-1, start) is matched on exactly one GPU."""
    start = n_frames * rank // world
    end = n_frames * (rank + 1) // world
    halo = 1 if (start > 0 and end > start) else 0
    return start, end, halo


class FrontEnd:
    def __init__(self, ctx, scale=2, num_octaves=1, k=1):
        self.ctx = ctx
        self.scale, self.num_octaves, self.k = scale, num_octaves, k

    def run(self, frames, start=0, end=None, halo=0):
        """Process frames[start-halo:end]; returns per-frame lists for frames [start, end):
        (keylines, descriptors, matches vs the previous frame; frame 0 of the sequence
        has all trainIdx = -1)."""
        ctx = self.ctx
        end = len(frames) if end is None else end
        lo = start - halo
        B, S, cap, k = ctx.max_batch, ctx.num_slots, ctx.max_lines, self.k
        bufs = [dict(kl=np.zeros((B, cap), capi.KEYLINE_DTYPE), counts=np.zeros(B, np.int32),
                     desc=np.zeros((B, cap, 32), np.uint8), m=np.zeros((B, cap, max(k, 1)), capi.DMATCH_DTYPE))
                for _ in range(S)]
        out_kl, out_desc, out_m = [], [], []
        pending = []  # (slot, first_frame, n)

        def collect(slot, f0, n):
            b = bufs[slot]
            ctx.collect_into(slot, b["kl"], b["counts"], cap, b["desc"], b["m"])
            for i in range(n):
                if f0 + i < start:
                    continue  # halo frame: only there to be matched against
                c = b["counts"][i]
                out_kl.append(b["kl"][i, :c].copy())
                out_desc.append(b["desc"][i, :c].copy())
                out_m.append(b["m"][i, :c].copy())

        slot = 0
        f = lo
        while f < end:
            n = min(B, end - f)
            if len(pending) == S:
                collect(*pending.pop(0))
            ctx.submit(slot, frames[f:f + n], scale=self.scale, num_octaves=self.num_octaves, k=k, chain=(f > lo))
            pending.append((slot, f, n))
            slot = (slot + 1) % S
            f += n
        while pending:
            collect(*pending.pop(0))
        return out_kl, out_desc, out_m

    def run_grouped(self, frames, start=0, end=None, halo=0):
        """The same through the fast path of the batch API: `frames` is one contiguous
        (n, h, w) uint8 array, pinned once; every batch is copied to the device ahead of its turn
        (vpl_frontend_upload), the slots' batches are submitted as one group (vpl_frontend_submit_group: their
        region-engine launches start together), results come back dense.  Same results and order as run()."""
        ctx = self.ctx
        frames = np.ascontiguousarray(frames)
        end = len(frames) if end is None else end
        lo = start - halo
        B, S, cap, k = ctx.max_batch, ctx.num_slots, ctx.max_lines, self.k
        h, w = frames.shape[1:]
        batches = [(f, min(B, end - f)) for f in range(lo, end, B)]
        bufs = [dict(kl=np.zeros(B * cap, capi.KEYLINE_DTYPE), counts=np.zeros(B, np.int32),
                     desc=np.zeros((B * cap, 32), np.uint8), m=np.zeros((B * cap, max(k, 1)), capi.DMATCH_DTYPE))
                for _ in range(S)]
        out_kl, out_desc, out_m = [], [], []

        def upload(bi):
            f0, n = batches[bi]
            ctx.upload(bi % S, frames[f0:f0 + n])

        def collect(bi):
            f0, n = batches[bi]
            b = bufs[bi % S]
            ctx.collect_dense_into(bi % S, b["counts"], b["kl"], b["desc"], b["m"])
            off = 0
            for i in range(n):
                c = int(b["counts"][i])
                if f0 + i >= start:  # (a halo frame is only there to be matched against)
                    out_kl.append(b["kl"][off:off + c].copy())
                    out_desc.append(b["desc"][off:off + c].copy())
                    out_m.append(b["m"][off:off + c].copy())
                off += c

        ctx.host_register(frames)
        try:
            # group g is submitted BEFORE group g-1 is collected (two result generations per slot): the download of
            # g-1 and the Python work on it overlap the kernels of g
            def submit(g):
                grp = range(g, min(g + S, len(batches)))
                ctx.submit_group([bi % S for bi in grp], [batches[bi][1] for bi in grp], w, h, scale=self.scale,
                                 num_octaves=self.num_octaves, k=k, chain=[bi > 0 for bi in grp])

            for bi in range(min(S, len(batches))):
                upload(bi)
            if batches:
                submit(0)
            for g in range(S, len(batches), S):
                for bi in range(g, min(g + S, len(batches))):
                    upload(bi)  # (group g-2 is collected: the slot's second input buffer is free)
                submit(g)
                for bi in range(g - S, g):
                    collect(bi)
            for bi in range((len(batches) - 1) // S * S if batches else 0, len(batches)):
                collect(bi)
        finally:
            ctx.host_unregister(frames)
        return out_kl, out_desc, out_m


class LineFrontEnd:
    """The reference's real per-frame loop (EDline on every frame + Matching(prev, cur),
    line_feature_tracker.cpp:87, :115) over a frame sequence through Context.linefront_submit /
    linefront_collect_into.  Batches, and shards of different ranks, overlap by one frame (the
    device matches frame f against frame f-1 inside a batch), so that every consecutive pair of the
    sequence is matched exactly once and no state crosses batches."""

    def __init__(self, ctx, smoothed=True):
        self.ctx = ctx
        self.smoothed = smoothed

    def run(self, frames, start=0, end=None, halo=0):
        """Frames [start, end) (plus `halo` frames in front, only there to be matched against).
        Returns (lines per frame, prev_to_cur per frame); the match row of frame 0 of the sequence
        is empty."""
        ctx = self.ctx
        end = len(frames) if end is None else end
        lo = start - halo
        B, S, cap = ctx.max_batch, ctx.num_slots, ctx.max_lines
        assert B >= 2, "a batch must hold the overlap frame and at least one new frame"
        bufs = [dict(lines=np.zeros((B, cap), capi.LINE_DTYPE), counts=np.zeros(B, np.int32),
                     p2c=np.zeros((B, cap), np.int32)) for _ in range(S)]
        out_lines, out_p2c = [], []
        pending = []  # (slot, first_frame, n, first_is_overlap)

        def collect(slot, f0, n, overlap):
            b = bufs[slot]
            ctx.linefront_collect_into(slot, b["lines"], b["counts"], cap, b["p2c"])
            for i in range(1 if overlap else 0, n):
                c = b["counts"][i]
                out_lines.append(b["lines"][i, :c].copy())
                out_p2c.append(b["p2c"][i, :b["counts"][i - 1]].copy() if i else b["p2c"][0, :0].copy())

        slot = 0
        f = start         # next frame whose results are still to be produced
        while f < end:
            overlap = f > lo              # a frame to match against precedes f: start the batch on it
            first = f - 1 if overlap else f
            n = min(B, end - first)
            if len(pending) == S:
                collect(*pending.pop(0))
            ctx.linefront_submit(slot, frames[first:first + n], smoothed=self.smoothed)
            pending.append((slot, first, n, overlap))
            slot = (slot + 1) % S
            f = first + n
        while pending:
            collect(*pending.pop(0))
        return out_lines, out_p2c


class VanishingPoints:
    """The vanishing-point stage readImage runs on every frame's lines after matching
    (vanishing_point_detection::run_vanishing_point_detection, line_feature_tracker.cpp:233-262) over a
    sequence of line sets through Context.vp_submit / vp_collect_into.  Frames are independent given their
    seed and their position in the sequence (the object's frame_count: only frame 0 of the whole sequence
    is a first call), so a shard is a plain contiguous range: no halo, no exchange."""

    def __init__(self, ctx):
        self.ctx = ctx

    def run(self, line_sets, seeds, start=0, end=None):
        """line_sets: one LINE_DTYPE array per frame; seeds: what time(NULL) returned per frame.
        Returns (vps (n,3,3), [labels per frame], status (n,)) for frames [start, end)."""
        ctx = self.ctx
        end = len(line_sets) if end is None else end
        B, S, cap = ctx.max_batch, ctx.num_slots, ctx.max_lines
        seeds = np.ascontiguousarray(seeds, np.uint32)
        bufs = [dict(vps=np.zeros((B, 3, 3), np.float64), idx=np.zeros((B, cap), np.int32), st=np.zeros(B, np.int32))
                for _ in range(S)]
        out_vps, out_idx, out_st = [], [], []
        pending = []  # (slot, first_frame, n)

        def collect(slot, f0, n):
            b = bufs[slot]
            ctx.vp_collect_into(slot, cap, b["vps"], b["idx"], b["st"])
            for i in range(n):
                out_vps.append(b["vps"][i].copy())
                out_idx.append(b["idx"][i, :len(line_sets[f0 + i])].copy())
                out_st.append(int(b["st"][i]))

        slot, f = 0, start
        while f < end:
            n = min(B, end - f)
            if len(pending) == S:
                collect(*pending.pop(0))
            lines = np.zeros((n, cap), capi.LINE_DTYPE)
            counts = np.zeros(n, np.int32)
            for i in range(n):
                counts[i] = len(line_sets[f + i])
                lines[i, :counts[i]] = np.asarray(line_sets[f + i]).view(capi.LINE_DTYPE).reshape(-1)
            ctx.vp_submit(slot, lines, counts, np.ascontiguousarray(seeds[f:f + n]), frame_count0=f)
            pending.append((slot, f, n))
            slot = (slot + 1) % S
            f += n
        while pending:
            collect(*pending.pop(0))
        return (np.array(out_vps).reshape(-1, 3, 3), out_idx, np.array(out_st, np.int32))


class LinePipeline:
    """readImage's whole line pipeline over a frame sequence through Context.readimage_submit /
    readimage_collect_into (remap + CLAHE when configured, EDline, Matching(previous frame, frame), vanishing
    points on the frame's own lines; line_feature_tracker.cpp:56-288).  Batches, and shards of different
    ranks, overlap by one frame so that every consecutive pair is matched exactly once; the overlap frame's
    vanishing points are computed twice and delivered once.  frame_count0 of a batch is the index of its first
    frame: only frame 0 of the whole sequence is the object's first call.

    This is the four pixel stages chained on every frame's RAW line set, not a replay of the tracker object:
    in the reference the first image never reaches the vanishing-point call (it sits inside
    `if (curframe_->vecLine.size() > 0)`, line_feature_tracker.cpp:109), so the stage's frame_count 0 falls on the SECOND
    image and its vps[1] / vps[2] swap rule starts one frame later than here; the reference only calls the stage on
    frames that keep more than 2 lines and does not count the others; and it matches against the previous frame's
    SELECTED, re-ordered line set.  For those semantics use the tracker object (compat/linefeature_tracker_b200.hpp /
    oracle.orc_tracker.Tracker), which restates that bookkeeping and is pinned against the reference's own readImage."""

    def __init__(self, ctx, smoothed=True):
        self.ctx = ctx
        self.smoothed = smoothed

    def run(self, frames, seeds, start=0, end=None, halo=0):
        """Frames [start, end) (plus `halo` frames in front, only there to be matched against).  Returns
        (lines per frame, prev_to_cur per frame, vps (n,3,3), labels per frame, status (n,))."""
        ctx = self.ctx
        end = len(frames) if end is None else end
        lo = start - halo
        B, S, cap = ctx.max_batch, ctx.num_slots, ctx.max_lines
        assert B >= 2, "a batch must hold the overlap frame and at least one new frame"
        seeds = np.ascontiguousarray(seeds, np.uint32)
        bufs = [dict(lines=np.zeros((B, cap), capi.LINE_DTYPE), counts=np.zeros(B, np.int32), p2c=np.zeros((B, cap), np.int32),
                     vps=np.zeros((B, 3, 3), np.float64), idx=np.zeros((B, cap), np.int32), st=np.zeros(B, np.int32))
                for _ in range(S)]
        out_lines, out_p2c, out_vps, out_idx, out_st = [], [], [], [], []
        pending = []  # (slot, first_frame, n, first_is_overlap)

        def collect(slot, f0, n, overlap):
            b = bufs[slot]
            ctx.readimage_collect_into(slot, b["lines"], b["counts"], cap, b["p2c"], b["vps"], b["idx"], b["st"])
            for i in range(1 if overlap else 0, n):
                c = b["counts"][i]
                out_lines.append(b["lines"][i, :c].copy())
                out_p2c.append(b["p2c"][i, :b["counts"][i - 1]].copy() if i else b["p2c"][0, :0].copy())
                out_vps.append(b["vps"][i].copy())
                out_idx.append(b["idx"][i, :c].copy())
                out_st.append(int(b["st"][i]))

        slot = 0
        f = start
        while f < end:
            overlap = f > lo
            first = f - 1 if overlap else f
            n = min(B, end - first)
            if len(pending) == S:
                collect(*pending.pop(0))
            ctx.readimage_submit(slot, frames[first:first + n], seeds[first:first + n], smoothed=self.smoothed, frame_count0=first)
            pending.append((slot, first, n, overlap))
            slot = (slot + 1) % S
            f = first + n
        while pending:
            collect(*pending.pop(0))
        return out_lines, out_p2c, np.array(out_vps).reshape(-1, 3, 3), out_idx, np.array(out_st, np.int32)
