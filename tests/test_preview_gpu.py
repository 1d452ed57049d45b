"""The RGB565X preview image with overlays: every byte equal to what the
host-built reference draws, for all five sensors, 1:1 / down-scaled / up-scaled / non-square previews.
The checker here is oracle/_ref itself (the reference's own drawing code); skipped where it is absent."""
import ctypes as C

import numpy as np
import pytest

from conftest import requires_ref
from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import open_sensor, sensors, synth, xdm

pytestmark = [pytest.mark.gpu, requires_ref]

GEOMS = [  # (w, h, out_w, out_h)
    (320, 240, 320, 240),
    (320, 240, 160, 120),
    (320, 240, 240, 320),      # the line sensors' default portrait preview
    (640, 480, 320, 240),
    (160, 120, 320, 240),      # up-scaling: unwritten pixels stay 0
    (320, 240, 200, 100),
]


def args_for(kind, variant):
    if kind == "oo":
        return [(1, 200, 45, 55, 40, 50, 45, 0), (1, 120, 25, 60, 35, 55, 40, 0)][variant]
    if kind == "om":
        return [(3, 3), (5, 7)][variant]
    return [(0, 359, 0, 100, 0, 40, 0), (300, 40, 20, 100, 30, 100, 0)][variant]


@pytest.mark.parametrize("kind", xdm.KIND_NAMES)
@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: "%dx%d_to_%dx%d" % g)
def test_preview_matches_reference(kind, geom):
    w, h, ow, oh = geom
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    rs = oracle.RefSensor(kind)
    assert rs.setup(w, h, out_w=ow, out_h=oh)[0] == 0
    codec = open_sensor(kind, w, h, out_w=ow, out_h=oh)
    orc = oracle.OracleSensor(kind, w, h)                 # only to learn which frames hit undefined reference behaviour
    fams = [("scene", 1), ("scene", 2), ("noise", 0), ("halves", 0), ("bluewrap", 0)]
    fams += [("noise", s) for s in range(1, 6)] if kind == "oo" else []
    fams += [("grid", 1)] if kind == "om" else []
    InAlg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]]
    checked = 0
    for variant in range(2):
        a = args_for(kind, variant)
        for fam, seed in fams:
            kw = {"m": a[0], "n": a[1]} if fam == "grid" else {}
            f0 = synth.make_frame(fam, seed, w, h, layout, **kw)
            fr = oracle.aligned_bytes(f0.size)
            fr[:] = f0
            orc.process(fr, oracle.IN_ARGS[kind](*a))
            ret, oa = codec.process(fr, InAlg(*a))
            assert ret == 0
            if orc.last_flags():
                # OO with fewer than 8 labels: the reference reads past its cluster vector, draws garbage targets
                # and can even divide by zero there -- do not run it on such frames
                continue
            rret, rout, _ = rs.process(fr, oracle.IN_ARGS[kind](*a))
            assert rret == 0
            want = rs.preview[:oh * ow * 2].copy()
            got = codec.preview[:oh * ow * 2]
            if not np.array_equal(got, want):
                bad = np.nonzero(got != want)[0]
                px = bad[0] // 2
                raise AssertionError("%s %s %s seed %d args %s: %d differing bytes, first at dst row %d col %d: got %02x%02x want %02x%02x"
                                     % (kind, geom, fam, seed, a, bad.size, px // ow, px % ow, got[2 * px + 1], got[2 * px],
                                        want[2 * px + 1], want[2 * px]))
            checked += 1
    assert checked >= (3 if kind == "oo" else 10)
    codec.close()


def test_batch_previews_equal_single_calls():
    w, h = 320, 240
    codec = open_sensor("wo", w, h, out_w=160, out_h=120)
    frames = np.stack([synth.make_frame("scene", s, w, h, "yuyv") for s in range(5)])
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 60, 0)
    prev = np.zeros((5, 160 * 120 * 2), dtype=np.uint8)
    ret, outs = codec.process_batch(frames, ia, previews=prev)
    assert ret == 0
    for i in range(5):
        r, oa = codec.process(frames[i], ia)
        assert r == 0
        assert np.array_equal(codec.preview[:160 * 120 * 2], prev[i])
    codec.close()


@pytest.mark.parametrize("kind", xdm.KIND_NAMES)
def test_one_to_one_batch_previews_in_sub_batches(kind):
    """1:1 previews of a batch go through the streaming kernel in sub-batches (so that the overlay stores hit L2): the
    images must not depend on the sub-batch size, and must equal single process() calls (which the test above pins to
    the reference's drawing)."""
    w, h, n = 320, 240, 21
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    fam = "grid" if kind == "om" else "blobs" if kind == "oo" else "scene"
    frames = np.stack([synth.make_frame(fam, s, w, h, layout) for s in range(n)])
    ia = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]](*args_for(kind, 0))
    L = sensors.lib()
    images = []
    try:
        for mb in (0, 1, 40):                      # whole batch at once (the default) / 8-frame sub-batches / 40 MiB
            L.trikb200_setPreviewChunkMB(mb)
            codec = open_sensor(kind, w, h, out_w=w, out_h=h)
            prev = np.full((n, w * h * 2), 0xAB, dtype=np.uint8)
            ret, outs = codec.process_batch(frames, ia, previews=prev)
            assert ret == 0, sensors.last_error()
            images.append(prev)
            codec.close()
    finally:
        L.trikb200_setPreviewChunkMB(0)
    assert np.array_equal(images[0], images[1]) and np.array_equal(images[0], images[2])
    codec = open_sensor(kind, w, h, out_w=w, out_h=h)
    for i in range(n):
        r, oa = codec.process(frames[i], ia)
        assert r == 0 and np.array_equal(codec.preview[:w * h * 2], images[0][i]), (kind, i)
    codec.close()


@pytest.mark.parametrize("kind", ["wl", "ol"])
@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120)])
def test_line_overlay_modes_give_the_same_previews(kind, size):
    """line sensors, 1:1 preview: the overlays come from the generic overlay kernel (0), from the full-sector kernel (1), from
    the last CTA of each frame inside the streaming kernel (2, 4: CTAs count themselves in at the frame's DrawInfo record) or
    from every CTA for its own rows (9, 12, and -1 = the default): every byte the same, on a second batch too (the counters
    must be back at zero), and equal to single process() calls, which test_preview_matches_reference pins to the reference's
    own drawing"""
    w, h = size
    n = 37
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    fams = ["scene", "halves", "noise", "bluewrap"]
    frames = np.stack([synth.make_frame(fams[s % 4], s, w, h, layout) for s in range(n)])
    ia = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]](*args_for(kind, 0))
    L = sensors.lib()
    images = {}
    try:
        for mode in (0, 1, 2, 4, 9, 12, -1):
            L.trikb200_setPreviewSectorOverlay(mode)
            codec = open_sensor(kind, w, h, out_w=w, out_h=h)
            for rep in range(2):
                prev = np.full((n, w * h * 2), 0xAB, dtype=np.uint8)
                ret, outs = codec.process_batch(frames, ia, previews=prev)
                assert ret == 0, sensors.last_error()
                images[(mode, rep)] = (prev, [bytes(memoryview(o)) for o in outs])
            codec.close()
    finally:
        L.trikb200_setPreviewSectorOverlay(-1)
    # OL carries state from frame to frame, so compare rep 0 with rep 0 and rep 1 with rep 1 across the modes
    for rep in range(2):
        for mode in (1, 2, 4, 9, 12, -1):
            assert images[(mode, rep)][1] == images[(0, rep)][1], (kind, size, mode, rep)
            assert np.array_equal(images[(mode, rep)][0], images[(0, rep)][0]), (kind, size, mode, rep)
    try:
        L.trikb200_setPreviewSectorOverlay(2)
        codec = open_sensor(kind, w, h, out_w=w, out_h=h)
        for i in range(n):
            r, oa = codec.process(frames[i], ia)
            assert r == 0 and np.array_equal(codec.preview[:w * h * 2], images[(0, 0)][0][i]), (kind, size, i)
        codec.close()
    finally:
        L.trikb200_setPreviewSectorOverlay(-1)


@pytest.mark.parametrize("kind", ["wl", "ol"])
def test_long_preview_pass_default_route(kind):
    """a long pass of 1:1 previews (700 device-resident 640x480 frames, eight blocks of 256 items per CTA by the default rule):
    every preview byte and every record equal to the route through the separate sector kernel, twice; one launch fewer"""
    import torch
    from trik_media_sensors_dsp_b200 import launch_count
    w, h, n, uniq = 640, 480, 700, 20
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    fams = ["scene", "halves", "noise", "bluewrap"]
    base = np.stack([synth.make_frame(fams[s % 4], s, w, h, layout) for s in range(uniq)])
    dev = torch.device("cuda", 0)
    d_frames = torch.from_numpy(base).to(dev).repeat((n + uniq - 1) // uniq, 1)[:n].contiguous()
    fbytes = base.shape[1]
    ia = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]](*args_for(kind, 0))
    L = sensors.lib()
    got = {}
    try:
        for mode in (1, -1):
            L.trikb200_setPreviewSectorOverlay(mode)
            codec = open_sensor(kind, w, h, out_w=w, out_h=h)
            for rep in range(2):
                d_prev = torch.full((n, w * h * 2), 0xAB, dtype=torch.uint8, device=dev)
                torch.cuda.synchronize()
                before = launch_count()
                ret, outs = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                                previews_device_ptr=d_prev.data_ptr(), preview_stride=w * h * 2)
                assert ret == 0, sensors.last_error()
                torch.cuda.synchronize()
                got[(mode, rep)] = (d_prev, [bytes(memoryview(o)) for o in outs], launch_count() - before)
            codec.close()
    finally:
        L.trikb200_setPreviewSectorOverlay(-1)
    for rep in range(2):
        assert got[(-1, rep)][1] == got[(1, rep)][1], (kind, rep)
        assert torch.equal(got[(-1, rep)][0], got[(1, rep)][0]), (kind, rep)
        assert got[(-1, rep)][2] == got[(1, rep)][2] - 1, (kind, rep, got[(-1, rep)][2], got[(1, rep)][2])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("size", [(320, 240), (640, 480)])
def test_wo_preview_through_the_chroma_table(size, variant):
    """webcam object sensor, 1:1 preview of a batch that goes through the chroma table of its one threshold set: the preview
    pass detects through the same table (two byte look-ups per pixel pair, the luma masks for ragged entries) instead of the
    HSV arithmetic -- every preview byte and every record equal to the arithmetic route (trikb200_setPreviewTable(0) and the
    arithmetic main pass) and to single process() calls, which test_preview_matches_reference pins to the reference"""
    w, h = size
    n = 41
    fams = ["scene", "noise", "halves", "bluewrap", "blobs"]
    frames = np.stack([synth.make_frame(fams[s % 5], s, w, h, "yuyv") for s in range(n)])
    ia = xdm.RangeInArgsAlg(*args_for("wo", variant))
    L = sensors.lib()
    got = {}
    try:
        for name, lut, pvt in (("table", 1, 1), ("table_main_only", 1, 0), ("arith", -1, 1)):
            L.trikb200_setLutMode(lut)
            L.trikb200_setPreviewTable(pvt)
            codec = open_sensor("wo", w, h, out_w=w, out_h=h)
            prev = np.full((n, w * h * 2), 0xAB, dtype=np.uint8)
            ret, outs = codec.process_batch(frames, ia, previews=prev)
            assert ret == 0, sensors.last_error()
            got[name] = (prev, [bytes(memoryview(o)) for o in outs])
            codec.close()
    finally:
        L.trikb200_setLutMode(0)
        L.trikb200_setPreviewTable(1)
    for name in ("table", "table_main_only"):
        assert got[name][1] == got["arith"][1], name
        assert np.array_equal(got[name][0], got["arith"][0]), (name, size, variant)
    assert (got["table"][0].reshape(n, -1, 2) == np.array([0xE0, 0xFF], dtype=np.uint8)).all(axis=2).any(), "nothing detected at all"
    codec = open_sensor("wo", w, h, out_w=w, out_h=h)
    for i in range(0, n, 5):
        r, oa = codec.process(frames[i], ia)
        assert r == 0 and np.array_equal(codec.preview[:w * h * 2], got["table"][0][i]), (size, variant, i)
    codec.close()
