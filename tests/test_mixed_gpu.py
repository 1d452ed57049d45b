"""trikb200_processMixed with many handles per sensor class (BASELINE config 4: one codec instance per stream): handles of
one (kind, geometry) share launches, and every frame must still be judged with the carried state of ITS OWN handle --
the one-frame lag of the ov7670 line sensor (ov7670/line_sensor/.../cv_line_detector_seqpass.hpp:449-450) and the object
sensor's persisting range (ov7670/object_sensor/.../cv_bitmap_builder_reference.hpp:110-130).  Checked against sequential
process() calls on fresh handles and against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import launch_count, open_sensor, process_mixed, sensors, synth, xdm

pytestmark = pytest.mark.gpu

KINDS = ["wl", "oo", "om", "wo", "ol"]
NBYTES = {"om": 36, "oo": 24}


def _args(kind, s, t):
    """per-handle arguments: every handle of a kind has its own thresholds; the object sensor sets its range at its own
    first frame only and then lives on the carried one"""
    if kind == "oo":
        if t == 0:
            return (1, 0, 10 + 3 * (s % 7), 70 + (s % 5), 20, 50 + (s % 3), 30, 0)
        return (0, 0, 0, 0, 0, 0, 0, 0)
    if kind == "om":
        return (3, 3) if s % 2 else (2, 4)
    if kind == "wo":
        return (300 - 5 * (s % 6), 40 + (s % 4), 20, 100, 30, 100, 0)
    return (0, 359, 0, 100, 0, 35 + (s % 9), 0)


def _in_alg(kind, a):
    return xdm.ObjInArgsAlg(*a) if kind == "oo" else xdm.MxnInArgsAlg(*a) if kind == "om" else xdm.RangeInArgsAlg(*a)


def _oracle_in(kind, a):
    return oracle.ObjInArgs(*a) if kind == "oo" else oracle.MxnInArgs(*a) if kind == "om" else oracle.RangeInArgs(*a)


def _frame(kind, s, t, w, h):
    fam = "blobs" if kind == "oo" else "grid" if kind == "om" else "scene"
    return synth.make_frame(fam, 31 * s + t, w, h, sensors.layout_of(xdm.KIND_OF[kind]))


@pytest.mark.parametrize("pinned,gather", [(True, 0), (True, 1), (True, -1), (False, 0)])
def test_many_handles_per_class_time_major(pinned, gather):
    import torch
    sizes = [(320, 240), (160, 120)]
    streams, T = 40, 4
    L = sensors.lib()
    L.trikb200_setGatherMode(gather)
    try:
        codecs, meta = [], []
        for s in range(streams):
            kind = KINDS[s % 5]
            w, h = sizes[(s // 5) % 2]
            codecs.append(open_sensor(kind, w, h))
            meta.append((kind, w, h))
        keep = []                                              # frames must stay alive (and pinned) until the call returns
        outs = {}
        l0 = launch_count()
        for t in range(T):                                     # time-major: one frame of every stream per call
            items = []
            for s, (kind, w, h) in enumerate(meta):
                f = _frame(kind, s, t, w, h)
                if pinned:
                    pf = torch.empty(f.shape, dtype=torch.uint8, pin_memory=True)
                    pf.numpy()[...] = f
                    keep.append(pf)
                    f = pf.numpy()
                oa = xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]]()
                outs[(s, t)] = oa
                items.append((codecs[s], f, _in_alg(kind, _args(kind, s, t)), oa, 7))
            assert process_mixed(items) == 0, sensors.last_error()
        launches = launch_count() - l0
        # 5 kinds x 2 geometries = 10 classes; a class is a handful of launches whatever the number of handles
        assert launches <= T * 10 * 5, launches
        for s, (kind, w, h) in enumerate(meta):
            fresh = open_sensor(kind, w, h)
            orc = oracle.OracleSensor(kind, w, h)
            nb = NBYTES.get(kind, 3)
            for t in range(T):
                f = _frame(kind, s, t, w, h)
                a = _args(kind, s, t)
                ret, oa = fresh.process(f, _in_alg(kind, a), seed=7)
                assert ret == 0
                got = bytes(memoryview(outs[(s, t)]))[:nb]
                assert got == bytes(memoryview(oa.alg))[:nb], (kind, s, t, "vs sequential process()")
                ok, exp = orc.process(f, _oracle_in(kind, a))
                undefined = (kind == "oo" and orc.last_flags()) or (kind == "ol" and t == 0)
                if not undefined:
                    assert ok == 1 and got == bytes(memoryview(exp))[:nb], (kind, s, t, "vs oracle")
            fresh.close()
        for c in codecs:
            c.close()
    finally:
        L.trikb200_setGatherMode(0)


def test_identical_per_frame_arguments_take_the_table_path():
    """a per-frame ARRAY of identical thresholds is a broadcast: same results, chroma-table kernel"""
    w, h, n = 320, 240, 300
    frames = synth.make_batch("scene", range(n), w, h, "yuyv")
    a = (300, 40, 20, 100, 30, 100, 0)
    codec = open_sensor("wo", w, h)
    ret, want = codec.process_batch(frames, xdm.RangeInArgsAlg(*a))
    assert ret == 0
    arr = (xdm.RangeInArgsAlg * n)(*[xdm.RangeInArgsAlg(*a) for _ in range(n)])
    l0 = launch_count()
    ret, got = codec.process_batch(frames, arr)
    assert ret == 0, sensors.last_error()
    assert launch_count() - l0 == 1                          # the table kernel alone (table already built)
    for i in range(n):
        assert bytes(memoryview(got[i]))[:3] == bytes(memoryview(want[i]))[:3]
    codec.close()
