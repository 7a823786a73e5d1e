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
