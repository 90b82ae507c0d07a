"""GPU parity tests proper: the CUDA path, called through the C ABI with host buffers,
against (a) the reference's own per-frame MD5s for the golden streams and (b) the CPU oracle
byte for byte over the WHOLE frame allocation (borders included).  Bit-exact or fail."""
import numpy as np
import pytest

from conftest import CASES, load_case
from oracle_lib import OracleDecoder
from vp8b200 import abi, frames
import randrec

pytestmark = pytest.mark.gpu


def _compare(got, want, geo, what):
    if np.array_equal(got, want):
        return
    m = geo.defined_mask()
    bad = np.flatnonzero((got != want) & m)
    if bad.size == 0:
        return                                    # only undefined row padding differs
    off = int(bad[0])
    if off < geo.yplane:
        r, c = divmod(off, geo.y_stride)
        where = "Y row %d col %d" % (r - 32, c - 32)
    else:
        o = (off - geo.yplane) % geo.uvplane
        r, c = divmod(o, geo.uv_stride)
        where = "%s row %d col %d" % ("U" if off < geo.yplane + geo.uvplane else "V", r - 16, c - 16)
    raise AssertionError("%s: %d bytes differ, first at %s (got %d want %d)"
                         % (what, bad.size, where, got[off], want[off]))


@pytest.mark.parametrize("name", CASES)
def test_golden_stream_bit_exact(gpu_lib, name):
    rec, md5s = load_case(name)
    geo = frames.Geometry(rec.coded_width, rec.coded_height)
    ctx = abi.Context(rec.coded_width, rec.coded_height, rec.n_fb)
    ora = OracleDecoder(rec.coded_width, rec.coded_height, rec.n_fb)
    shown = 0
    for i, fr in enumerate(rec.frames):
        ctx.submit(fr)
        ora.frame(fr)
        fb = int(fr.hdr["fb_new"])
        got = ctx.fetch(fb)
        _compare(got, ora.fb(fb), geo, "%s frame %d (type %d)" % (name, i, fr.hdr["frame_type"]))
        if fr.show_frame:
            m = frames.md5_hex(geo.i420(got, rec.display_width, rec.display_height))
            assert m == md5s[shown], "%s frame %d: MD5 differs from the reference decoder" % (name, i)
            shown += 1
    assert shown == len(md5s)
    assert ctx.launch_count() > 0
    ctx.close()


RANDOM_CASES = [
    # (mb_cols, mb_rows, kwargs)
    (1, 1, dict()),                                          # single macroblock
    (2, 1, dict()), (1, 3, dict()),                          # ragged tiny frames
    (5, 4, dict(key=True)),
    (7, 5, dict(filter_type=1)),
    (11, 9, dict(bilinear=True, filter_type=1)),
    (11, 9, dict(bilinear=True, full_pixel=True, filter_level=0)),
    (22, 18, dict(p_intra=0.5, big_coefs=True)),
    (22, 18, dict(p_intra=0.0, p_split=1.0, p_skip=0.0, coef_density=1.0)),   # all split, dense
    (22, 18, dict(p_skip=1.0, filter_level=63, sharpness=0)),
    (40, 3, dict(sharpness=7)),
    (33, 2, dict(key=True, coef_density=1.0, big_coefs=True)),
    (80, 45, dict()),                                        # 720p
    (65, 5, dict(filter_type=1)),                            # record batches of 32 + 32 + 1, simple filter
    (32, 6, dict(sharpness=3)), (9, 70, dict()),             # exactly one batch; more rows than a CTA wave is deep
]


@pytest.mark.parametrize("idx", range(len(RANDOM_CASES)))
def test_random_records_bit_exact(gpu_lib, idx):
    mb_cols, mb_rows, kw = RANDOM_CASES[idx]
    rng = np.random.default_rng(1000 + idx)
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    ctx = abi.Context(w, h, 4)
    ora = OracleDecoder(w, h, 4)
    for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
        ctx.upload(fb, buf)
        ora.fb(fb)[:] = buf
    fbs = [0, 1, 2, 3]
    for rep in range(3):
        kw2 = dict(kw)
        if rep == 1 and "filter_level" not in kw2:
            kw2["filter_level"] = int(rng.integers(1, 64))
        fr = randrec.random_frame(rng, mb_cols, mb_rows, fbs=tuple(fbs), **kw2)
        ctx.submit(fr)
        ora.frame(fr)
        _compare(ctx.fetch(fbs[0]), ora.fb(fbs[0]), geo, "random case %d rep %d" % (idx, rep))
        fbs = fbs[1:] + fbs[:1]                   # next frame predicts from the one just made
    ctx.close()


def test_batched_streams_match_oracle(gpu_lib):
    """vp8b200_batch_run: one launch per kernel over several independent streams."""
    n_streams, mb_cols, mb_rows = 5, 22, 18
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(77)
    ctxs = [abi.Context(w, h, 4) for _ in range(n_streams)]
    oras = [OracleDecoder(w, h, 4) for _ in range(n_streams)]
    for c, o in zip(ctxs, oras):
        for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
            c.upload(fb, buf)
            o.fb(fb)[:] = buf
    fbs = [0, 1, 2, 3]
    for step in range(4):
        frs = [randrec.random_frame(rng, mb_cols, mb_rows, key=(step == 0 and s == 1),
                                    filter_level=(0 if s == 2 else None),
                                    p_intra=(0.0 if s == 3 else 0.15), fbs=tuple(fbs))
               for s in range(n_streams)]
        staged = [c.stage(fr) for c, fr in zip(ctxs, frs)]
        abi.batch_run(ctxs, staged)
        ctxs[0].sync()
        for s in range(n_streams):
            oras[s].frame(frs[s])
            _compare(ctxs[s].fetch(fbs[0]), oras[s].fb(fbs[0]), geo, "batch step %d stream %d" % (step, s))
        fbs = fbs[1:] + fbs[:1]
    for c in ctxs:
        c.close()


def test_1080p_round_trip_properties(gpu_lib):
    """BASELINE-size frame (1920x1088 coded): full compare against the oracle on one P frame
    plus size-independent properties: a zero-MV, no-residual, no-filter frame reproduces its
    reference exactly, and reconstruction is deterministic across repeats."""
    mb_cols, mb_rows = 120, 68
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(5)
    ctx = abi.Context(w, h, 4)
    ora = OracleDecoder(w, h, 4)
    bufs = randrec.random_buffers(rng, geo.frame_size, 4)
    for fb, buf in enumerate(bufs):
        ctx.upload(fb, buf)
        ora.fb(fb)[:] = buf
    fr = randrec.random_frame(rng, mb_cols, mb_rows, fbs=(0, 1, 2, 3))
    ctx.submit(fr)
    first = ctx.fetch(0).copy()
    ora.frame(fr)
    _compare(first, ora.fb(0), geo, "1080p random P frame")
    ctx.submit(fr)                                         # determinism
    assert np.array_equal(ctx.fetch(0), first)
    # identity frame: ZEROMV from `last`, everything skipped, loop filter off
    ident = randrec.random_frame(rng, mb_cols, mb_rows, p_intra=0.0, p_split=0.0, p_skip=1.0,
                                 filter_level=0, fbs=(2, 1, 1, 1))
    ident.mb["y_mode"] = 7
    ident.mb["ref_frame"] = 1
    ident.mb["mv_row"] = 0
    ident.mb["mv_col"] = 0
    ident.mb["flags"] = 4
    ctx.submit(ident)
    y0, u0, v0 = geo.planes(bufs[1])
    y1, u1, v1 = geo.planes(ctx.fetch(2))
    assert np.array_equal(y0, y1) and np.array_equal(u0, u1) and np.array_equal(v0, v1)
    ctx.close()


def test_batch_then_individual_use_is_ordered(gpu_lib):
    """A member of a batch is written on the leader's stream; fetching it through its own
    context right away (no explicit sync) must still see the finished frame, and a frame
    submitted individually must be visible to a following batch."""
    n_streams, mb_cols, mb_rows = 6, 40, 30
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(123)
    ctxs = [abi.Context(w, h, 4) for _ in range(n_streams)]
    oras = [OracleDecoder(w, h, 4) for _ in range(n_streams)]
    for c, o in zip(ctxs, oras):
        for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
            c.upload(fb, buf)
            o.fb(fb)[:] = buf
    # 1) individual submit on every context, immediately followed by a batch that predicts from it
    f0 = [randrec.random_frame(rng, mb_cols, mb_rows, fbs=(0, 1, 2, 3)) for _ in range(n_streams)]
    for c, o, fr in zip(ctxs, oras, f0):
        c.submit(fr)
        o.frame(fr)
    f1 = [randrec.random_frame(rng, mb_cols, mb_rows, p_intra=0.05, fbs=(1, 0, 0, 0)) for _ in range(n_streams)]
    staged = [c.stage(fr) for c, fr in zip(ctxs, f1)]
    abi.batch_run(ctxs, staged)
    # 2) no sync: fetch members through their own contexts
    for s in reversed(range(n_streams)):
        oras[s].frame(f1[s])
        _compare(ctxs[s].fetch(1), oras[s].fb(1), geo, "ordered batch stream %d" % s)
    for c in ctxs:
        c.close()


def test_leader_destroyed_before_its_batch_members(gpu_lib):
    """The members of a batch borrow an event of the leader (first context of the batch).
    Destroying the leader right after the batch - no sync, members untouched so far - must
    leave the members usable: their frames are complete and later calls do not wait on the
    dead leader."""
    n_streams, mb_cols, mb_rows = 4, 22, 18
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(321)
    ctxs = [abi.Context(w, h, 4) for _ in range(n_streams)]
    oras = [OracleDecoder(w, h, 4) for _ in range(n_streams)]
    for c, o in zip(ctxs, oras):
        for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
            c.upload(fb, buf)
            o.fb(fb)[:] = buf
    f0 = [randrec.random_frame(rng, mb_cols, mb_rows, fbs=(0, 1, 2, 3)) for _ in range(n_streams)]
    staged = [ctxs[0].stage(fr) for fr in f0]            # the leader owns the staged frames too
    abi.batch_run(ctxs, staged)
    ctxs[0].close()                                       # waits for its own batch, retires the borrowed events
    f1 = [randrec.random_frame(rng, mb_cols, mb_rows, p_intra=0.1, fbs=(1, 0, 0, 0)) for _ in range(n_streams)]
    for s in range(1, n_streams):
        oras[s].frame(f0[s])
        _compare(ctxs[s].fetch(0), oras[s].fb(0), geo, "member %d after the leader is gone" % s)
        ctxs[s].submit(f1[s])
        oras[s].frame(f1[s])
        _compare(ctxs[s].fetch(1), oras[s].fb(1), geo, "member %d next frame" % s)
    # the surviving contexts can form a new batch with a new leader
    f2 = [randrec.random_frame(rng, mb_cols, mb_rows, fbs=(2, 1, 0, 0)) for _ in range(n_streams)]
    staged = [ctxs[1].stage(f2[s]) for s in range(1, n_streams)]
    abi.batch_run(ctxs[1:], staged)
    for s in range(1, n_streams):
        oras[s].frame(f2[s])
        _compare(ctxs[s].fetch(2), oras[s].fb(2), geo, "member %d in the second batch" % s)
    for c in ctxs[1:]:
        c.close()


def test_two_pictures_in_flight_are_collected_oldest_first(gpu_lib):
    """Frame-delay order at the C ABI: picture N is queued (frame_submit_show)
    BEFORE picture N-1 is collected; fetch_wait must hand over N-1 - complete and correct - while
    N may still be running."""
    mb_cols, mb_rows, n_frames = 9, 6, 7
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(123)
    ctx = abi.Context(w, h, 4)
    ora = OracleDecoder(w, h, 4)
    for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
        ctx.upload(fb, buf)
        ora.fb(fb)[:] = buf
    outs = [np.zeros(geo.frame_size, np.uint8) for _ in range(3)]
    want = {}
    # two buffers alternate as new / last: frame f writes the buffer picture f-2 was copied from
    fbs = [0, 1, 2, 3]
    for f in range(n_frames):
        fr = randrec.random_frame(rng, mb_cols, mb_rows, key=(f == 0), p_intra=0.1, fbs=tuple(fbs))
        ctx.submit_show(fr, show_fb=fbs[0], out=outs[f % 3], display=(w, h))
        ora.frame(fr)
        want[f] = geo.i420(ora.fb(fbs[0]), w, h)
        if f >= 1:
            ctx.fetch_wait()                                      # collects picture f-1
            assert geo.i420(outs[(f - 1) % 3], w, h) == want[f - 1], f - 1
        fbs = [fbs[1], fbs[0]] + fbs[2:]
    ctx.fetch_wait()
    assert geo.i420(outs[(n_frames - 1) % 3], w, h) == want[n_frames - 1]
    ctx.close()


def test_coalesced_submit_batches_frames_of_many_contexts(gpu_lib):
    """SURVEY 8b "shared batch scheduler across ctxs": vp8b200_frame_submit_show hands frames to
    the per-device engine, whose thread issues ONE launch of each kernel over the frames of all
    contexts that queued one, plus their uploads and the copies of the shown pictures.  Results
    must equal the oracle's, frame after frame (every frame predicts from the previous one), and
    the engine must really have batched."""
    import ctypes as C
    n_streams, mb_cols, mb_rows, n_frames = 8, 11, 7, 5
    w, h = mb_cols * 16, mb_rows * 16
    geo = frames.Geometry(w, h)
    rng = np.random.default_rng(77)
    ctxs = [abi.Context(w, h, 4) for _ in range(n_streams)]
    oras = [OracleDecoder(w, h, 4) for _ in range(n_streams)]
    for c, o in zip(ctxs, oras):
        for fb, buf in enumerate(randrec.random_buffers(rng, geo.frame_size, 4)):
            c.upload(fb, buf)
            o.fb(fb)[:] = buf
    st0 = (C.c_uint64 * 2)()
    gpu_lib.vp8b200_engine_stats(0, st0)
    fbs = [0, 1, 2, 3]
    outs = [np.zeros(geo.frame_size, np.uint8) for _ in range(n_streams)]
    for f in range(n_frames):
        frs = [randrec.random_frame(rng, mb_cols, mb_rows, key=(f == 0 and s % 2 == 0), p_intra=0.1, fbs=tuple(fbs))
               for s in range(n_streams)]
        for s in range(n_streams):                           # every decoder queues its frame ...
            ctxs[s].submit_show(frs[s], show_fb=fbs[0], out=outs[s], display=(w - 5, h - 3))
        for s in range(n_streams):                           # ... and only then waits for its picture
            ctxs[s].fetch_wait()
            oras[s].frame(frs[s])
            want = oras[s].fb(fbs[0])
            assert geo.i420(outs[s], w - 5, h - 3) == geo.i420(want, w - 5, h - 3), (f, s)
        # the whole buffer (borders included) through the synchronous fetch of the same context
        _compare(ctxs[3].fetch(fbs[0]), oras[3].fb(fbs[0]), geo, "coalesced frame %d stream 3" % f)
        fbs = fbs[1:] + fbs[:1]
    st1 = (C.c_uint64 * 2)()
    gpu_lib.vp8b200_engine_stats(0, st1)
    assert st1[1] - st0[1] == n_streams * n_frames
    assert st1[0] - st0[0] < n_streams * n_frames, "no two frames ever shared a launch"
    # mixing with the direct path on the same context keeps the order
    fr = randrec.random_frame(rng, mb_cols, mb_rows, fbs=tuple(fbs))
    ctxs[0].submit(fr)
    oras[0].frame(fr)
    _compare(ctxs[0].fetch(fbs[0]), oras[0].fb(fbs[0]), geo, "direct submit after coalesced ones")
    for c in ctxs:
        c.close()
