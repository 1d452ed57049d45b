"""One explicit parity case per BASELINE.json config, named after it (the broader sweeps live in the other files).

config 1  ov7670 object sensor, ONE synthetic 320x240 frame, reference C built for the host: the reference-style
          process() call against the reference build itself (result record, every xDM bookkeeping field, the preview)
          -- and, since BASELINE words the input as "RGB565", the same picture arriving as packed RGB565 through the
          ingest front end (the reference has no such input; its sensors read YUV422P).
config 2  webcam line sensor, 4096 x 320x240 YUYV: the batch's distinct frames against the reference build.
config 3  mxn grid sensor 3x3 / 5x5 on 640x480 + auto-detect HSV: against the reference build.
config 4  mixed line + object + mxn instances over concurrent streams: trikb200_processMixed against the oracle.
config 5  frame sizes 160x120 .. 1920x1080: WL / WO against the oracle (the reference build stops at 640x480)."""
import ctypes as C

import numpy as np
import pytest

from conftest import requires_ref
from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import open_sensor, process_mixed, sensors, synth, xdm

pytestmark = pytest.mark.gpu


def _frame_with_8_labels(w, h):
    """BASELINE config 1 names seed 1 of the scene family; the reference's tail reads past its cluster vector when a
    frame has fewer than 8 labels, so the first scene seed from 1 on that is defined is taken."""
    orc = oracle.OracleSensor("oo", w, h)
    for seed in range(1, 200):
        for a in ((1, 0, 20, 80, 20, 50, 30, 0), (1, 200, 45, 55, 40, 50, 45, 0), (1, 120, 60, 50, 50, 50, 50, 0)):
            f = synth.make_frame("scene", seed, w, h, "yuv422p")
            orc.process(f, oracle.ObjInArgs(*a))
            if not orc.last_flags():
                return seed, a, f
    raise AssertionError("no scene frame with 8 labels")


@requires_ref
def test_config1_single_object_sensor_frame_against_the_reference_build():
    w, h = 320, 240
    seed, a, f0 = _frame_with_8_labels(w, h)
    fr = oracle.aligned_bytes(f0.size)
    fr[:] = f0
    rs = oracle.RefSensor("oo")
    assert rs.setup(w, h, out_w=w, out_h=h)[0] == 0
    codec = open_sensor("oo", w, h, out_w=w, out_h=h)
    for call in range(2):                                    # twice: carried state (the packed range) included
        rret, rout, _ = rs.process(fr, oracle.ObjInArgs(*a))
        ret, out = codec.process(fr, xdm.ObjInArgsAlg(*a))
        assert (ret, rret) == (0, 0)
        assert bytes(memoryview(out.alg))[:36] == bytes(memoryview(rout))[:36], (seed, a, call)
        assert np.array_equal(codec.preview[:w * h * 2], rs.preview[:w * h * 2])
    codec.close()


def test_config1_same_picture_arriving_as_rgb565():
    """RGB565 words -> ingest front end -> ov7670 object sensor == the oracle on the converted bytes."""
    w, h = 320, 240
    rng = np.random.default_rng(1)
    img = np.full((h, w), 0x2104, np.uint16)                                  # dark grey
    for k in range(12):                                                       # twelve red-ish blocks: more than 8 labels
        x0, y0 = 4 + 24 * k, 8 + 16 * k
        img[y0:y0 + 24, x0:x0 + 20] = np.uint16(0xF800 | (k << 5))
    img ^= (rng.integers(0, 2, (h, w), dtype=np.uint16))                      # one LSB of blue noise
    ret, frames = sensors.ingest_rgb565(img.view(np.uint8).reshape(1, -1), w, h)
    assert ret == 0, sensors.last_error()
    a = (1, 0, 25, 75, 25, 60, 40, 0)
    orc = oracle.OracleSensor("oo", w, h)
    ok, exp = orc.process(frames[0], oracle.ObjInArgs(*a))
    assert ok == 1 and not orc.last_flags()
    codec = open_sensor("oo", w, h)
    ret, out = codec.process(frames[0], xdm.ObjInArgsAlg(*a))
    assert ret == 0
    assert bytes(memoryview(out.alg))[:24] == bytes(memoryview(exp))[:24]
    assert any(out.alg.target[i].size for i in range(8))                          # something was found
    codec.close()


@requires_ref
def test_config2_line_sensor_batch_against_the_reference_build():
    w, h, n, uniq = 320, 240, 4096, 64
    a = (0, 359, 0, 100, 0, 40, 0)
    hu = synth.make_batch("scene", range(uniq), w, h, "yuyv")
    frames = np.concatenate([hu] * (n // uniq))
    codec = open_sensor("wl", w, h)
    ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*a))
    assert ret == 0, sensors.last_error()
    rs = oracle.RefSensor("wl")
    assert rs.setup(w, h)[0] == 0
    fr = oracle.aligned_bytes(hu.shape[1])
    for i in range(uniq):
        fr[:] = hu[i]
        rret, rout, _ = rs.process(fr, oracle.RangeInArgs(*a))
        assert rret == 0
        for j in (i, i + uniq, n - uniq + i):
            assert bytes(memoryview(outs[j]))[:3] == bytes(memoryview(rout))[:3], (i, j)
    codec.close()


@requires_ref
@pytest.mark.parametrize("grid", [(3, 3), (5, 5)])
def test_config3_mxn_grid_against_the_reference_build(grid):
    w, h, n = 640, 480, 96
    m, g = grid
    frames = synth.make_batch("grid", range(n), w, h, "yuv422p", m=m, n=g)
    codec = open_sensor("om", w, h)
    ret, outs = codec.process_batch(frames, xdm.MxnInArgsAlg(m, g))
    assert ret == 0, sensors.last_error()
    rs = oracle.RefSensor("om")
    assert rs.setup(w, h)[0] == 0
    fr = oracle.aligned_bytes(frames.shape[1])
    for i in range(0, n, 3):
        fr[:] = frames[i]
        rret, rout, _ = rs.process(fr, oracle.MxnInArgs(m, g))
        assert rret == 0
        assert list(outs[i].outColor[:m * g]) == list(rout.outColor[:m * g]), (grid, i)
    codec.close()


@requires_ref
def test_config3_auto_detect_hsv_against_the_reference_build():
    """The deterministic family (webcam object sensor) straight against the reference build; the annealed families are
    pinned through the seeded oracle in test_sensors_gpu / test_anneal_gpu (the reference seeds with time(NULL))."""
    w, h = 640, 480
    a = (0, 359, 0, 100, 0, 100, 1)
    frames = synth.make_batch("scene", range(24), w, h, "yuyv")
    codec = open_sensor("wo", w, h)
    ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*a))
    assert ret == 0, sensors.last_error()
    rs = oracle.RefSensor("wo")
    assert rs.setup(w, h)[0] == 0
    fr = oracle.aligned_bytes(frames.shape[1])
    for i in range(frames.shape[0]):
        fr[:] = frames[i]
        rret, rout, _ = rs.process(fr, oracle.RangeInArgs(*a))
        assert rret == 0
        assert bytes(memoryview(outs[i])) == bytes(memoryview(rout)), i
    codec.close()


def test_config4_mixed_instances_over_streams():
    w, h = 320, 240
    kinds = ["wl", "oo", "om", "wo", "ol"]
    codecs, items, wants = [], [], []
    for s in range(20):                                      # 20 streams, 3 frames each, kinds cycling
        kind = kinds[s % len(kinds)]
        layout = sensors.layout_of(xdm.KIND_OF[kind])
        codec = open_sensor(kind, w, h)
        orc = oracle.OracleSensor(kind, w, h)
        codecs.append(codec)
        for t in range(3):
            if kind == "oo":
                a, fam = (1, 0, 20, 80, 20, 50, 30, 0), "blobs"
                ia, oia = xdm.ObjInArgsAlg(*a), oracle.ObjInArgs(*a)
            elif kind == "om":
                a, fam = (3, 3), "grid"
                ia, oia = xdm.MxnInArgsAlg(*a), oracle.MxnInArgs(*a)
            else:
                a, fam = (0, 359, 0, 100, 0, 45, 0), "scene"
                ia, oia = xdm.RangeInArgsAlg(*a), oracle.RangeInArgs(*a)
            f = synth.make_frame(fam, 100 * s + t, w, h, layout)
            oa = xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]]()
            items.append((codec, f, ia, oa, None))
            ok, exp = orc.process(f, oia)
            nbytes = 36 if kind == "om" else (24 if kind == "oo" else 3)
            skip = (kind == "oo" and orc.last_flags()) or (kind == "ol" and t == 0)
            wants.append(None if skip else bytes(memoryview(exp))[:nbytes])
    assert process_mixed(items) == 0, sensors.last_error()
    for (codec, f, ia, oa, _), want in zip(items, wants):
        if want is not None:
            assert bytes(memoryview(oa))[:len(want)] == want
    for c in codecs:
        c.close()


@pytest.mark.parametrize("size", [(160, 120), (320, 240), (640, 480), (1280, 720), (1920, 1080)])
def test_config5_frame_size_sweep(size):
    w, h = size
    for kind, a in (("wl", (0, 359, 0, 100, 0, 40, 0)), ("wo", (300, 40, 20, 100, 30, 100, 0))):
        frames = synth.make_batch("scene", range(6), w, h, "yuyv")
        codec = open_sensor(kind, w, h)
        ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*a))
        assert ret == 0, sensors.last_error()
        orc = oracle.OracleSensor(kind, w, h)
        for i in range(frames.shape[0]):
            ok, exp = orc.process(frames[i], oracle.RangeInArgs(*a))
            assert ok == 1 and bytes(memoryview(outs[i]))[:3] == bytes(memoryview(exp))[:3], (kind, size, i)
        codec.close()
