"""Host-side logic without a GPU: frame sharding with a one-frame halo, the pipelined batch
driver, and the N>1 path over gloo (world_size 2).  The CUDA context is replaced by a
stand-in that answers from the CPU oracle -- test infrastructure only; the product driver
never sees the oracle."""
import importlib
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleContext:
    """Quacks like capi.Context (submit / collect_into) but computes with the oracle."""

    def __init__(self, orc, capi, max_batch=4, num_slots=2, max_lines=1024):
        self.orc, self.capi = orc, capi
        self.max_batch, self.num_slots, self.max_lines = max_batch, num_slots, max_lines
        self.pending = {}
        self.prev_desc = None
        self.submitted = []
        self.staged = {}
        self.queued = {}
        self.log = []

    def submit(self, slot, frames, scale=2, num_octaves=1, k=1, chain=False):
        assert slot not in self.pending and len(frames) <= self.max_batch
        out = []
        prev = self.prev_desc if chain else None
        for img in frames:
            kl = self.orc.lsd_detector_detect(img, scale, num_octaves)
            d = self.orc.lbd_compute(img, kl)
            m = np.zeros((len(kl), k), self.capi.DMATCH_DTYPE)
            m["queryIdx"] = np.arange(len(kl))[:, None]
            m["trainIdx"] = -1
            if prev is not None:
                idx, dist = self.orc.hamming_knn(d, prev, k)
                m["trainIdx"] = idx
                m["distance"] = dist
            out.append((kl, d, m))
            prev = d
        self.prev_desc = prev
        self.pending[slot] = out
        self.submitted.append(len(frames))
        return len(frames)

    def collect_into(self, slot, kl, counts, cap, desc, matches):
        for i, (k_, d, m) in enumerate(self.pending.pop(slot)):
            counts[i] = len(k_)
            kl[i, :len(k_)] = k_
            desc[i, :len(k_)] = d
            matches[i, :len(k_)] = m

    # the upload-ahead / group-submit / dense-collect calls of run_grouped, with the C library's rules
    def host_register(self, a):
        self.registered = a

    def host_unregister(self, a):
        assert a is self.registered
        self.registered = None

    def upload(self, slot, frames):
        assert getattr(self, "registered", None) is not None and np.shares_memory(frames, self.registered)
        assert slot not in self.staged and len(frames) <= self.max_batch  # one uploaded batch per slot
        assert len(self.queued.get(slot, [])) < 2  # the older of two uncollected batches still owns the buffer
        self.staged[slot] = frames
        self.log.append(("upload", slot, len(frames)))

    def submit_group(self, slots, ns, w, h, scale=2, num_octaves=1, k=1, chain=None):
        assert len(set(slots)) == len(slots)
        for s_, n in zip(slots, ns):
            # an uploaded batch may queue behind ONE uncollected batch of its slot
            assert s_ in self.staged and len(self.staged[s_]) == n and len(self.queued.get(s_, [])) < 2
        self.log.append(("group", tuple(slots), tuple(ns)))
        for s_, c in zip(slots, chain):
            held = self.pending.pop(s_, None)
            self.submit(s_, self.staged.pop(s_), scale=scale, num_octaves=num_octaves, k=k, chain=c)
            self.queued.setdefault(s_, []).append(self.pending.pop(s_))
            if held is not None:
                self.pending[s_] = held

    def collect_dense_into(self, slot, counts, kl, desc, matches):
        off = 0
        for i, (k_, d, m) in enumerate(self.queued[slot].pop(0)):  # oldest first
            counts[i] = len(k_)
            kl[off:off + len(k_)] = k_
            desc[off:off + len(k_)] = d
            matches[off:off + len(k_)] = m
            off += len(k_)
        self.log.append(("collect", slot))
        return off


class OracleLineContext:
    """Quacks like capi.Context's linefront_submit / linefront_collect_into, computing with the oracle
    restatement of the reference's EDLines + LineMatching."""

    def __init__(self, orc, capi, max_batch=4, num_slots=2, max_lines=512):
        self.orc, self.capi = orc, capi
        self.max_batch, self.num_slots, self.max_lines = max_batch, num_slots, max_lines
        self.pending = {}
        self.submitted = []
        self.param = orc.EDLineParam(minLineLen=15)

    def linefront_submit(self, slot, frames, smoothed=True):
        assert slot not in self.pending and 1 <= len(frames) <= self.max_batch
        lines = [self.orc.edline_detect(f, self.param, smoothed) for f in frames]
        p2c = [None] + [self.orc.line_matching(frames[i - 1], frames[i], lines[i - 1], lines[i]) for i in range(1, len(frames))]
        self.pending[slot] = (lines, p2c)
        self.submitted.append(len(frames))
        return len(frames)

    def linefront_collect_into(self, slot, lines, counts, cap, p2c):
        ls, ms = self.pending.pop(slot)
        p2c[:] = -1
        for i, l in enumerate(ls):
            counts[i] = len(l)
            lines[i, :len(l)] = l.view(self.capi.LINE_DTYPE)
            if i and ms[i] is not None:
                p2c[i, :len(ms[i])] = ms[i]


class OracleVpContext:
    """Quacks like capi.Context's vp_submit / vp_collect_into, computing with the oracle restatement of the
    reference's vanishing-point stage (device arithmetic: math_mode 1)."""
    CAM = (230.0, 80.0, 60.0)

    def __init__(self, orc, capi, max_batch=4, num_slots=2, max_lines=256):
        self.orc, self.capi = orc, capi
        self.max_batch, self.num_slots, self.max_lines = max_batch, num_slots, max_lines
        self.pending = {}
        self.submitted = []

    def vp_submit(self, slot, lines, n_lines, seeds, frame_count0=0):
        assert slot not in self.pending and 1 <= len(n_lines) <= self.max_batch
        out = []
        for i in range(len(n_lines)):
            ln = lines[i, :n_lines[i]]
            if len(ln) < 2:
                out.append((np.zeros((3, 3)), np.full(len(ln), 3, np.int32), -1))
                continue
            vps, idx, d = self.orc.vp_detect(ln, None, *self.CAM, int(seeds[i]), frame_count0 + i, math_mode=1, details=True)
            out.append((vps, idx, d["flags"] & 1))
        self.pending[slot] = out
        self.submitted.append((len(n_lines), frame_count0))
        return len(n_lines)

    def vp_collect_into(self, slot, cap, vps, vp_idx, status=None, line_vps=None):
        for i, (v, idx, st) in enumerate(self.pending.pop(slot)):
            vps[i] = v
            vp_idx[i, :len(idx)] = idx
            if status is not None:
                status[i] = st


class OraclePipelineContext(OracleLineContext):
    """Quacks like capi.Context's readimage_submit / readimage_collect_into: the oracle's EDLines + matching, then the
    oracle's vanishing-point stage (device arithmetic) on each frame's own lines."""
    CAM = (230.0, 80.0, 60.0)

    def readimage_submit(self, slot, frames, seeds, smoothed=True, frame_count0=0):
        self.linefront_submit(slot, frames, smoothed)
        lines, p2c = self.pending[slot]
        vp = []
        for i, ln in enumerate(lines):
            if len(ln) < 2:
                vp.append((np.zeros((3, 3)), np.full(len(ln), 3, np.int32), -1))
            else:
                v, idx, d = self.orc.vp_detect(ln, None, *self.CAM, int(seeds[i]), frame_count0 + i, math_mode=1, details=True)
                vp.append((v, idx, d["flags"] & 1))
        self.pending[slot] = (lines, p2c, vp)
        self.submitted[-1] = (len(frames), frame_count0)

    def readimage_collect_into(self, slot, lines, counts, cap, p2c, vps, vp_idx, vp_status=None):
        ls, ms, vp = self.pending[slot]
        self.pending[slot] = (ls, ms)
        self.linefront_collect_into(slot, lines, counts, cap, p2c)
        for i, (v, idx, st) in enumerate(vp):
            vps[i] = v
            vp_idx[i, :len(idx)] = idx
            if vp_status is not None:
                vp_status[i] = st


def test_line_pipeline_driver_batches_and_shards(vpl, orc, synth):
    frames = synth.sequence(8, w=192, h=128, seed=6, n_quads=6, n_strokes=10)
    seeds = np.arange(300, 308, dtype=np.uint32)
    ctx = OraclePipelineContext(orc, vpl.capi, max_batch=3, num_slots=2)
    lines, p2c, vps, idx, st = vpl.LinePipeline(ctx).run(frames, seeds)
    # 8 frames in batches of 3 with a one-frame overlap; frame_count0 = index of the batch's first frame
    assert ctx.submitted == [(3, 0), (3, 2), (3, 4), (2, 6)] and len(lines) == 8
    el, em = _line_front_expected(orc, frames, ctx.param)
    for f in range(8):
        assert lines[f].tobytes() == el[f].tobytes() and np.array_equal(p2c[f], em[f])
        if len(el[f]) >= 2:
            ev, ei = orc.vp_detect(el[f], None, *OraclePipelineContext.CAM, int(seeds[f]), f, math_mode=1)
            assert vps[f].tobytes() == ev.tobytes() and np.array_equal(idx[f], ei), f
    for world in (2, 3):
        got = [[], [], [], []]
        for r in range(world):
            s, e, halo = vpl.shard_range(len(frames), r, world)
            a, b, v, i, _ = vpl.LinePipeline(OraclePipelineContext(orc, vpl.capi, max_batch=4)).run(frames, seeds, s, e, halo)
            got[0] += a; got[1] += b; got[2] += list(v); got[3] += i
        assert all(x.tobytes() == y.tobytes() for x, y in zip(got[0], lines))
        assert all(np.array_equal(x, y) for x, y in zip(got[1], p2c))
        assert np.array(got[2]).tobytes() == vps.tobytes() and all(np.array_equal(x, y) for x, y in zip(got[3], idx))


def _vp_inputs(orc, synth, n=7):
    frames = synth.sequence(n, w=160, h=120, seed=33, n_quads=6, n_strokes=10)
    sets = [orc.edline_detect(f, orc.EDLineParam(minLineLen=12), True) for f in frames]
    sets[3] = sets[3][:1]  # a frame the tracker would not run the stage on
    return sets, np.arange(500, 500 + n, dtype=np.uint32)


def test_vanishing_point_driver_batches_and_shards(vpl, orc, synth):
    sets, seeds = _vp_inputs(orc, synth)
    exp = [orc.vp_detect(s, None, *OracleVpContext.CAM, int(sd), f, math_mode=1) if len(s) >= 2 else None
           for f, (s, sd) in enumerate(zip(sets, seeds))]
    ctx = OracleVpContext(orc, vpl.capi, max_batch=3, num_slots=2)
    vps, idx, st = vpl.VanishingPoints(ctx).run(sets, seeds)
    assert ctx.submitted == [(3, 0), (3, 3), (1, 6)] and st[3] == -1 and (vps[3] == 0).all()
    for f, e in enumerate(exp):
        if e is not None:
            assert vps[f].tobytes() == e[0].tobytes() and np.array_equal(idx[f], e[1])
    # frames shard without a halo: only frame 0 of the sequence is a first call, whichever rank holds it
    for world in (2, 3):
        got_v, got_i = [], []
        for r in range(world):
            s, e, _ = vpl.shard_range(len(sets), r, world)
            v, i, _ = vpl.VanishingPoints(OracleVpContext(orc, vpl.capi, max_batch=2)).run(sets, seeds, s, e)
            got_v += list(v); got_i += i
        assert np.array(got_v).tobytes() == vps.tobytes() and all(np.array_equal(a, b) for a, b in zip(got_i, idx))


def _line_front_expected(orc, frames, param):
    lines = [orc.edline_detect(f, param, True) for f in frames]
    p2c = [np.zeros(0, np.int32)]
    for i in range(1, len(frames)):
        m = orc.line_matching(frames[i - 1], frames[i], lines[i - 1], lines[i])
        p2c.append(np.full(len(lines[i - 1]), -1, np.int32) if m is None else m)
    return lines, p2c


def test_line_front_driver_overlaps_batches_by_one_frame(vpl, orc, synth):
    frames = synth.sequence(8, w=192, h=128, seed=6, n_quads=6, n_strokes=10)
    ctx = OracleLineContext(orc, vpl.capi, max_batch=3, num_slots=2)
    lines, p2c = vpl.LineFrontEnd(ctx).run(frames)
    # 8 frames in batches of 3 with a one-frame overlap: [0,1,2] [2,3,4] [4,5,6] [6,7]
    assert ctx.submitted == [3, 3, 3, 2] and len(lines) == 8
    el, em = _line_front_expected(orc, frames, ctx.param)
    for f in range(8):
        assert lines[f].tobytes() == el[f].tobytes()
        assert np.array_equal(p2c[f], em[f]), f


def test_line_front_sharded_run_equals_single_run(vpl, orc, synth):
    frames = synth.sequence(9, w=160, h=120, seed=9, n_quads=5, n_strokes=8)
    el, em = _line_front_expected(orc, frames, orc.EDLineParam(minLineLen=15))
    for world in (1, 2, 3):
        lines, p2c = [], []
        for r in range(world):
            s, e, halo = vpl.shard_range(len(frames), r, world)
            a, b = vpl.LineFrontEnd(OracleLineContext(orc, vpl.capi, max_batch=4)).run(frames, s, e, halo)
            lines += a
            p2c += b
        assert len(lines) == 9
        for f in range(9):
            assert lines[f].tobytes() == el[f].tobytes() and np.array_equal(p2c[f], em[f]), (world, f)


def test_shard_range_partitions_every_pair_once(vpl):
    for n in (1, 2, 7, 100, 20000):
        for world in (1, 2, 3, 4, 8):
            covered, pairs = [], []
            for r in range(world):
                s, e, halo = vpl.shard_range(n, r, world)
                covered += list(range(s, e))
                lo = s - halo
                pairs += [(f - 1, f) for f in range(lo + 1, e)]
                assert halo in (0, 1) and (halo == 0 or s > 0)
            assert covered == list(range(n))
            assert sorted(pairs) == [(f - 1, f) for f in range(1, n)], (n, world)


def test_driver_pipelining_and_chaining(vpl, orc, synth):
    frames = synth.sequence(7, w=192, h=128, seed=5, n_quads=6, n_strokes=10)
    ctx = OracleContext(orc, vpl.capi, max_batch=3, num_slots=2)
    kls, descs, ms = vpl.FrontEnd(ctx, k=2).run(frames)
    assert ctx.submitted == [3, 3, 1] and len(kls) == 7
    prev = None
    for f, img in enumerate(frames):
        ekl = orc.lsd_detector_detect(img, 2, 1)
        ed = orc.lbd_compute(img, ekl)
        assert all(np.array_equal(kls[f][n], ekl[n]) for n in ekl.dtype.names)
        assert np.array_equal(descs[f], ed)
        if prev is None:
            assert (ms[f]["trainIdx"] == -1).all()
        else:
            idx, _ = orc.hamming_knn(ed, prev, 2)
            assert np.array_equal(ms[f]["trainIdx"], idx)
        prev = ed


def test_grouped_driver_equals_plain_driver(vpl, orc, synth):
    """FrontEnd.run_grouped (upload ahead, group submits, dense collects) delivers what run() delivers, whole and
    sharded, and keeps the C library's rules: one uploaded batch per slot, a slot collected before it is resubmitted,
    uploads issued right after the group that frees the slot's second input buffer."""
    frames = np.ascontiguousarray(synth.sequence(9, w=160, h=120, seed=8, n_quads=5, n_strokes=8))
    full = vpl.FrontEnd(OracleContext(orc, vpl.capi, max_batch=4), k=1).run(frames)
    ctx = OracleContext(orc, vpl.capi, max_batch=2, num_slots=2)
    got = vpl.FrontEnd(ctx, k=1).run_grouped(frames)
    assert ctx.log[:3] == [("upload", 0, 2), ("upload", 1, 2), ("group", (0, 1), (2, 2))]
    # group 1 is uploaded and submitted before group 0 is collected
    assert ctx.log[3:8] == [("upload", 0, 2), ("upload", 1, 2), ("group", (0, 1), (2, 2)), ("collect", 0), ("collect", 1)]
    assert ctx.log[-5:] == [("upload", 0, 1), ("group", (0,), (1,)), ("collect", 0), ("collect", 1), ("collect", 0)]
    assert ctx.registered is None
    for world in (1, 2, 3):
        part = ([], [], [])
        for r in range(world):
            s, e, halo = vpl.shard_range(len(frames), r, world)
            out = vpl.FrontEnd(OracleContext(orc, vpl.capi, max_batch=2), k=1).run_grouped(frames, s, e, halo)
            for a, b in zip(part, out):
                a.extend(b)
        for ref, a in ((full, got), (full, part)):
            for x, y in zip(ref, a):
                assert len(x) == len(y)
                assert all(p.tobytes() == q.tobytes() for p, q in zip(x, y)), world


def test_sharded_run_equals_single_run(vpl, orc, synth):
    frames = synth.sequence(9, w=160, h=120, seed=8, n_quads=5, n_strokes=8)
    full = vpl.FrontEnd(OracleContext(orc, vpl.capi, max_batch=4), k=1).run(frames)
    for world in (2, 3):
        got = ([], [], [])
        for r in range(world):
            s, e, halo = vpl.shard_range(len(frames), r, world)
            part = vpl.FrontEnd(OracleContext(orc, vpl.capi, max_batch=2), k=1).run(frames, s, e, halo)
            for a, b in zip(got, part):
                a.extend(b)
        for a, b in zip(got, full):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and np.array_equal(x, y) if x.dtype.names is None else \
                    all(np.array_equal(x[n], y[n]) for n in x.dtype.names)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    vpl = importlib.import_module("vplines_slam_b200")
    synth = importlib.import_module("vplines-slam_b200.synth")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = synth.sequence(6, w=160, h=120, seed=21, n_quads=5, n_strokes=8)
    s, e, halo = vpl.shard_range(len(frames), rank, world)
    kls, descs, ms = vpl.FrontEnd(OracleContext(O, vpl.capi, max_batch=2), k=1).run(frames, s, e, halo)
    # the reference's real front end (EDLines + KLT line matching) through the same sharding
    ll, pp = vpl.LineFrontEnd(OracleLineContext(O, vpl.capi, max_batch=3)).run(frames, s, e, halo)
    # ... and the vanishing-point stage on the shard's own line sets (no halo)
    sets = [O.edline_detect(f, O.EDLineParam(minLineLen=15), True) for f in frames]
    _, vidx, _ = vpl.VanishingPoints(OracleVpContext(O, vpl.capi, max_batch=2)).run(sets, np.arange(70, 70 + len(frames)), s, e)
    # ... and the fused pipeline driver over the same shard: same lines, matches and labels as the separate drivers
    pl = vpl.LinePipeline(OraclePipelineContext(O, vpl.capi, max_batch=3)).run(frames, np.arange(70, 70 + len(frames)), s, e, halo)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(pl[0], ll)) and all(np.array_equal(a, b) for a, b in zip(pl[1], pp))
    assert all(np.array_equal(a, b) for a, b in zip(pl[3], vidx))
    # no data-path collective: the host only gathers results (here: per-frame line counts and match sums)
    mine = torch.tensor([[len(k), int(m["trainIdx"].astype(np.int64).sum()) + 1000 * len(l) + 7 * int((p >= 0).sum())
                          + 100000 * int((v * np.arange(1, len(v) + 1)).sum())]
                         for k, m, l, p, v in zip(kls, ms, ll, pp, vidx)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)], dtype=torch.int64))
    pad = torch.zeros((len(frames), 2), dtype=torch.int64)
    pad[:len(mine)] = mine
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank == 0:
        q.put(torch.cat([p[:int(n)] for p, n in zip(parts, sizes)]).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather(vpl, orc, synth):
    world, port = 2, 29000 + os.getpid() % 2000
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = synth.sequence(6, w=160, h=120, seed=21, n_quads=5, n_strokes=8)
    kls, descs, ms = vpl.FrontEnd(OracleContext(orc, vpl.capi, max_batch=8), k=1).run(frames)
    ll, pp = vpl.LineFrontEnd(OracleLineContext(orc, vpl.capi, max_batch=8)).run(frames)
    _, vidx, _ = vpl.VanishingPoints(OracleVpContext(orc, vpl.capi, max_batch=8)).run(ll, np.arange(70, 70 + len(frames)))
    exp = np.array([[len(k), int(m["trainIdx"].astype(np.int64).sum()) + 1000 * len(l) + 7 * int((p >= 0).sum())
                     + 100000 * int((v * np.arange(1, len(v) + 1)).sum())]
                    for k, m, l, p, v in zip(kls, ms, ll, pp, vidx)])
    assert np.array_equal(got, exp)
