"""The chroma-indexed detection table of the webcam object sensor (csrc/trik_kernels_lut.cu): the table must
reproduce the arithmetic threshold on ALL 2^24 (Y,U,V) inputs for every threshold set tried, and the sensor
run through it must give the oracle's bytes."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, synth, xdm

pytestmark = pytest.mark.gpu

THRESHOLDS = [
    (300, 40, 20, 100, 30, 100, 0),    # hue wraps through 0
    (90, 150, 35, 100, 35, 100, 0),    # green-ish
    (0, 359, 0, 100, 0, 100, 0),       # everything
    (0, 359, 0, 100, 0, 40, 0),        # dark
    (200, 260, 40, 100, 20, 90, 0),
    (10, 20, 50, 60, 50, 60, 0),       # narrow in all three
    (0, 0, 0, 0, 60, 30, 0),           # empty value range
    (350, 10, 0, 10, 90, 100, 0),      # near-white reds
    (120, 120, 50, 50, 50, 50, 0),     # single values
]


@pytest.fixture(autouse=True)
def _defaults_after():
    yield
    lib().trikb200_setLutMode(0)


@pytest.mark.parametrize("args", THRESHOLDS)
def test_table_equals_arithmetic_on_all_inputs(args):
    ia = xdm.RangeInArgsAlg(*args)
    stats = (C.c_uint64 * 5)()
    assert lib().trikb200_probeLut(C.addressof(ia), C.addressof(stats)) == 0, lib().trikb200_lastError()
    mismatches, never, interval, ragged, passing = [int(v) for v in stats]
    assert never + interval + ragged == 65536
    assert mismatches == 0, (args, mismatches, never, interval, ragged, passing)


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 4), (96, 8), (1280, 720), (640, 44)])
def test_wo_through_table_matches_oracle(size):
    w, h = size
    fams = [("noise", s) for s in range(3)] + [("scene", s) for s in range(4)] + [(e, 0) for e in synth.EDGE_CASES]
    if w < 64 or h < 16:
        fams = [("noise", s) for s in range(4)] + [("zero", 0), ("full", 0), ("bluewrap", 0), ("greyramp", 0)]
    frames = np.stack([synth.make_frame(f, s, w, h, "yuyv") for f, s in fams])
    frames = np.concatenate([frames] * 3)                       # more frames than one CTA's groups
    codec = open_sensor("wo", w, h)
    for args in THRESHOLDS:
        orc = oracle.OracleSensor("wo", w, h)
        want = [bytes(memoryview(orc.process(frames[i], oracle.RangeInArgs(*args))[1]))[:3] for i in range(len(fams))] * 3
        for mode in (1, -1):
            lib().trikb200_setLutMode(mode)
            before = lib().trikb200_launchCount()
            ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*args))
            assert ret == 0, lib().trikb200_lastError()
            got = [bytes(memoryview(o))[:3] for o in outs]
            assert got == want, (size, args, mode, [i for i in range(len(got)) if got[i] != want[i]][:8])
    codec.close()


def test_table_is_reused_and_rebuilt():
    """Same thresholds -> one table build; new thresholds -> a new build; large batches pick the table on their own."""
    w, h = 320, 240
    frames = np.stack([synth.make_frame("scene", s % 7, w, h, "yuyv") for s in range(256)])
    codec = open_sensor("wo", w, h)
    a1, a2 = xdm.RangeInArgsAlg(300, 40, 20, 100, 30, 100, 0), xdm.RangeInArgsAlg(90, 150, 35, 100, 35, 100, 0)
    lib().trikb200_setLutMode(0)
    n0 = lib().trikb200_launchCount()
    assert codec.process_batch(frames, a1)[0] == 0
    n1 = lib().trikb200_launchCount()
    assert codec.process_batch(frames, a1)[0] == 0
    n2 = lib().trikb200_launchCount()
    assert codec.process_batch(frames, a2)[0] == 0
    n3 = lib().trikb200_launchCount()
    assert (n1 - n0, n2 - n1, n3 - n2) == (2, 1, 2)            # build + pass, pass, build + pass
    codec.close()


OBJ_ARGS = [(1, 0, 20, 80, 20, 50, 30, 0), (1, 120, 25, 60, 35, 55, 40, 0), (1, 0, 40, 60, 40, 60, 40, 0), (1, 230, 30, 70, 30, 50, 40, 0)]


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 8), (1280, 720)])
def test_oo_bitmap_through_table_matches_oracle(size):
    """OO with step 1 (threshold -> metapixel bitmap) through the table, one threshold set per batch."""
    w, h = size
    fams = [("noise", s) for s in range(3)] + [("scene", s) for s in range(3)] + [("blobs", s) for s in range(4)] + [(e, 0) for e in synth.EDGE_CASES]
    if w < 64:
        fams = [("noise", s) for s in range(4)] + [("zero", 0), ("red", 0)]
    frames = np.stack([synth.make_frame(f, s, w, h, "yuv422p") for f, s in fams])
    codec = open_sensor("oo", w, h)
    for args in OBJ_ARGS:
        orc = oracle.OracleSensor("oo", w, h)
        want = [bytes(memoryview(orc.process(frames[i], oracle.ObjInArgs(*args))[1]))[:24] for i in range(len(fams))]
        for mode in (1, -1):
            lib().trikb200_setLutMode(mode)
            ret, outs = codec.process_batch(frames, xdm.ObjInArgsAlg(*args))
            assert ret == 0, lib().trikb200_lastError()
            got = [bytes(memoryview(o))[:24] for o in outs]
            assert got == want, (size, args, mode, [i for i in range(len(got)) if got[i] != want[i]][:8])
    codec.close()


@pytest.mark.parametrize("size", [(320, 240), (96, 8), (640, 480)])
def test_skewed_table_layout_gives_the_same_bytes(size):
    """trikb200_setLutSkew: rows of the shared-memory table 260 (default) or 256 bytes apart -- same results."""
    w, h = size
    fams = [("camera", s) for s in range(4)] + [("noise", s) for s in range(3)] + [("scene", 1), ("bluewrap", 0), ("full", 0)]
    frames = np.concatenate([np.stack([synth.make_frame(f, s, w, h, "yuyv") for f, s in fams])] * 3)
    codec = open_sensor("wo", w, h)
    try:
        for args in THRESHOLDS[:6]:
            orc = oracle.OracleSensor("wo", w, h)
            want = [bytes(memoryview(orc.process(frames[i], oracle.RangeInArgs(*args))[1]))[:3] for i in range(len(fams))] * 3
            lib().trikb200_setLutMode(1)
            for skew in (1, 0):
                lib().trikb200_setLutSkew(skew)
                ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*args))
                assert ret == 0, lib().trikb200_lastError()
                got = [bytes(memoryview(o))[:3] for o in outs]
                assert got == want, (size, args, skew, [i for i in range(len(got)) if got[i] != want[i]][:8])
    finally:
        lib().trikb200_setLutSkew(1)
        codec.close()


def test_batch_under_several_threshold_sets_takes_one_table_per_set():
    """frames of one batch under DIFFERENT thresholds (per-stream thresholds gathered into one batch) used
    to fall back to the arithmetic kernel.  Now the batch is partitioned by threshold set, one cached table per set and one
    launch whose persistent CTAs are dealt out to the sets; results must be the per-frame results of the arithmetic path and of the oracle."""
    from trik_media_sensors_dsp_b200 import launch_count
    w, h, n = 320, 240, 400
    sets = [THRESHOLDS[0], THRESHOLDS[1], THRESHOLDS[4], THRESHOLDS[7], THRESHOLDS[3]]
    frames = synth.make_batch("scene", range(n), w, h, "yuyv")
    arr = (xdm.RangeInArgsAlg * n)(*[xdm.RangeInArgsAlg(*sets[(i * 7) % len(sets)]) for i in range(n)])
    codec = open_sensor("wo", w, h)
    lib().trikb200_setLutMode(-1)                                   # the arithmetic kernel as the in-library comparison
    ret, want = codec.process_batch(frames, arr)
    assert ret == 0
    lib().trikb200_setLutMode(0)
    ret, got = codec.process_batch(frames, arr)                     # builds the five tables
    assert ret == 0, lib().trikb200_lastError()
    l0 = launch_count()
    ret, got2 = codec.process_batch(frames, arr)                    # tables cached: ONE launch, its CTAs dealt out to the sets
    assert ret == 0 and launch_count() - l0 == 1
    orc = oracle.OracleSensor("wo", w, h)
    for i in range(n):
        assert bytes(memoryview(got[i]))[:3] == bytes(memoryview(want[i]))[:3], i
        assert bytes(memoryview(got2[i]))[:3] == bytes(memoryview(want[i]))[:3], i
        if i % 9 == 0:
            ok, exp = orc.process(frames[i], oracle.RangeInArgs(*sets[(i * 7) % len(sets)]))
            assert ok == 1 and bytes(memoryview(got[i]))[:3] == bytes(memoryview(exp))[:3], i
    # a set with fewer than 32 frames in the batch: the whole batch takes the arithmetic kernel, same bytes
    arr3 = (xdm.RangeInArgsAlg * n)(*[xdm.RangeInArgsAlg(*(THRESHOLDS[5] if i == 17 else sets[i % 2])) for i in range(n)])
    lib().trikb200_setLutMode(-1)
    ret, want3 = codec.process_batch(frames, arr3)
    lib().trikb200_setLutMode(0)
    ret, got3 = codec.process_batch(frames, arr3)
    assert ret == 0
    for i in range(n):
        assert bytes(memoryview(got3[i]))[:3] == bytes(memoryview(want3[i]))[:3], i
    codec.close()


def test_oo_batch_under_several_ranges_takes_the_tables():
    """object sensor frames whose ranges differ inside one batch (instances gathered by trikb200_processMixed, or a batch that
    re-sets the range now and then): step 1 through one cached table per range, in one launch; same bytes as the arithmetic
    step 1 and as the oracle"""
    from trik_media_sensors_dsp_b200 import launch_count
    w, h, n = 320, 240, 320
    ranges = [(1, 0, 20, 80, 20, 50, 30, 0), (1, 200, 45, 55, 40, 50, 45, 0), (1, 120, 60, 50, 50, 50, 50, 0)]
    frames = np.stack([synth.make_frame("blobs" if i % 2 else "scene", i, w, h, "yuv422p") for i in range(n)])
    # every 40 frames the range is set anew (cycling through the three), in between it is the carried one
    arr = (xdm.ObjInArgsAlg * n)(*[xdm.ObjInArgsAlg(*ranges[(i // 40) % 3]) if i % 40 == 0 else xdm.ObjInArgsAlg(0, 0, 0, 0, 0, 0, 0, 0)
                                   for i in range(n)])
    codec = open_sensor("oo", w, h)
    lib().trikb200_setLutMode(-1)
    ret, want = codec.process_batch(frames, arr)
    assert ret == 0
    lib().trikb200_setLutMode(0)
    codec2 = open_sensor("oo", w, h)
    ret, got = codec2.process_batch(frames, arr)
    assert ret == 0, lib().trikb200_lastError()
    codec3 = open_sensor("oo", w, h)
    ret, _ = codec3.process_batch(frames, arr)
    l0 = launch_count()
    codec3.set_params(w, h)
    ret, got3 = codec3.process_batch(frames, arr)              # tables cached: step 1 + labelling, two launches
    assert ret == 0 and launch_count() - l0 == 2
    orc = oracle.OracleSensor("oo", w, h)
    for i in range(n):
        assert bytes(memoryview(got[i]))[:24] == bytes(memoryview(want[i]))[:24], i
        assert bytes(memoryview(got3[i]))[:24] == bytes(memoryview(want[i]))[:24], i
        a = ranges[(i // 40) % 3] if i % 40 == 0 else (0, 0, 0, 0, 0, 0, 0, 0)
        ok, exp = orc.process(frames[i], oracle.ObjInArgs(*a))
        if not orc.last_flags():
            assert ok == 1 and bytes(memoryview(got[i]))[:24] == bytes(memoryview(exp))[:24], i
    for c in (codec, codec2, codec3):
        c.close()


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120)])
def test_bands_of_rows_give_the_same_records(size):
    """the table kernel's work items are whole frames or 2 / 4 / 8 bands of rows meeting in the frame's accumulator record
    (chosen per launch; forced here): the records must not depend on it -- one threshold set and several"""
    w, h = size
    n = 330
    frames = synth.make_batch("scene", range(n), w, h, "yuyv")
    one = xdm.RangeInArgsAlg(*THRESHOLDS[0])
    many = (xdm.RangeInArgsAlg * n)(*[xdm.RangeInArgsAlg(*THRESHOLDS[(0, 1, 4)[i % 3]]) for i in range(n)])
    L = lib()
    L.trikb200_setLutMode(1)
    results = {}
    try:
        for parts in (1, 2, 4, 8, 0):
            L.trikb200_setLutParts(parts)
            codec = open_sensor("wo", w, h)
            for name, ia in (("one", one), ("many", many)):
                for rep in range(2):                              # twice: the accumulator records must be back at zero
                    ret, outs = codec.process_batch(frames, ia)
                    assert ret == 0, L.trikb200_lastError()
                    results[(parts, name, rep)] = [bytes(memoryview(o))[:3] for o in outs]
            codec.close()
    finally:
        L.trikb200_setLutParts(0)
    for name in ("one", "many"):
        want = results[(1, name, 0)]
        for parts in (1, 2, 4, 8, 0):
            for rep in range(2):
                assert results[(parts, name, rep)] == want, (size, parts, name, rep)
    orc = oracle.OracleSensor("wo", w, h)
    for i in range(0, n, 11):
        ok, exp = orc.process(frames[i], oracle.RangeInArgs(*THRESHOLDS[0]))
        assert ok == 1 and results[(1, "one", 0)][i] == bytes(memoryview(exp))[:3], i
