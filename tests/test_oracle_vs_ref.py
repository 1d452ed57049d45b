"""Pin the oracle (oracle/trik_oracle.c) against the reference's own sources built for the host
(oracle/_ref/libtrikref_*.so).  CPU only; skipped where the reference build is absent."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import requires_ref
from oracle import ref
from trik_media_sensors_dsp_b200 import synth

pytestmark = requires_ref
N24 = 1 << 24


@pytest.fixture(scope="module")
def oracle_tables():
    lib = ref.port_lib()
    rgb = np.empty(N24, np.uint32)
    hsv = np.empty(N24, np.uint32)
    lib.trik_oracle_yuv_to_rgb888_range(0, N24, rgb.ctypes.data)
    lib.trik_oracle_rgb888_to_hsv_range(0, N24, hsv.ctypes.data)
    return rgb, hsv


@pytest.mark.parametrize("kind", ref.KINDS)
def test_pixel_functions_exhaustive(kind, oracle_tables):
    """All 2^24 (Y,U,V) and all 2^24 RGB inputs: closed forms == the intrinsic code of every sensor."""
    rgb, hsv = oracle_tables
    rs = ref.RefSensor(kind)
    assert rs.setup(320, 240)[0] == 0
    r1 = np.empty(N24, np.uint32)
    r1b = np.empty(N24, np.uint32)
    r2 = np.empty(N24, np.uint32)
    rs.lib.trikref_probe_yuv2rgb(0, N24, r1.ctypes.data, r1b.ctypes.data)
    assert rs.lib.trikref_probe_rgb2hsv(0, N24, r2.ctypes.data) == 0
    assert np.array_equal(r1, rgb)
    idx = np.arange(N24, dtype=np.uint32)
    assert np.array_equal(r1b, rgb[(255 - (idx & 0xFF)) | (idx & 0xFFFF00)])     # second pixel of the word
    assert np.array_equal(r2, hsv)
    # the blue lane really does wrap negative (129*U - 17672 + 74*Y > 32767) for 106 (Y,U) pairs, and
    # those pixels come out with B = 0 in the reference: the closed form must keep that quirk
    y = (idx[:65536] & 0xFF).astype(np.int64)
    u = (idx[:65536] >> 8).astype(np.int64)
    wraps = (129 * u - 17672 + 74 * y) > 32767
    assert int(wraps.sum()) == 106
    assert np.all((r1[:65536][wraps] & 0xFF) == 0)


def _args_for(kind):
    if kind in ("wo", "wl", "ol"):
        return [(0, 359, 0, 100, 0, 40, 0), (300, 40, 20, 100, 30, 100, 0), (10, 200, 0, 60, 20, 90, 1), (0, 359, 0, 100, 30, 100, 1)]
    if kind == "oo":
        return [(1, 0, 20, 80, 20, 50, 30, 0), (0, 0, 0, 0, 0, 0, 0, 0), (1, 120, 25, 60, 35, 55, 40, 0), (1, 0, 40, 60, 40, 60, 40, 1)]
    return [(3, 3), (5, 5), (2, 7), (10, 10), (1, 1)]


@pytest.mark.parametrize("kind", ref.KINDS)
@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 4)])
def test_sensor_frames(kind, size):
    """Seeded frames through both, one codec instance per (kind, size) so carried state is exercised."""
    w, h = size
    layout = "yuyv" if kind in ("wo", "wl") else "yuv422p"
    rs = ref.RefSensor(kind)
    assert rs.setup(w, h)[0] == 0
    orc = ref.OracleSensor(kind, w, h)
    fams = [("noise", 0), ("noise", 3), ("scene", 0), ("scene", 4), ("blobs", 1), ("blobs", 2)] + [(e, 0) for e in synth.EDGE_CASES]
    compared = 0
    for args in _args_for(kind):
        if kind == "om" and (args[0] > h or args[1] > w):
            continue
        for call, (fam, fseed) in enumerate(fams):
            f0 = synth.make_frame(fam, fseed, w, h, layout)
            fr = ref.aligned_bytes(f0.size)
            fr[:] = f0
            seed = 77 + call
            ok, oout = orc.process(fr, ref.IN_ARGS[kind](*args), seed=seed)
            assert ok == 1
            if orc.last_flags() and kind == "oo":
                continue                                   # fewer than 8 labels: the reference reads past its vector
                                                           # (garbage sizes, sometimes SIGFPE) -- never run it there
            ret, out, _ = rs.process(fr, ref.IN_ARGS[kind](*args), seed=seed)
            assert ret == 0
            if orc.last_flags():
                continue                                   # undefined behaviour of the reference, see trik_oracle.h
            a, b = ref.struct_bytes(out), ref.struct_bytes(oout)
            if kind == "om":
                a, b = a[:args[0] * args[1] * 4], b[:args[0] * args[1] * 4]
            assert a == b, (kind, size, args, fam, fseed, a.hex(), b.hex())
            compared += 1
    assert compared > 10 or (kind == "oo" and w < 160)      # a 32x4 frame cannot hold 8 labels


def test_object_sensor_sort_ties():
    """Noise + sparse range gives hundreds of labels with many equal sizes: the restated std::sort
    must reproduce libstdc++'s tie order (the eight reported targets depend on it)."""
    w, h = 640, 480
    rs = ref.RefSensor("oo")
    assert rs.setup(w, h)[0] == 0
    orc = ref.OracleSensor("oo", w, h)
    nontrivial = 0
    for seed in range(12):
        f0 = synth.make_frame("noise", 100 + seed, w, h, "yuv422p")
        fr = ref.aligned_bytes(f0.size)
        fr[:] = f0
        for args in [(1, 120, 25 + seed, 60, 35, 55, 40, 0), (1, 30 * seed, 20, 50, 30, 50, 45, 0)]:
            ok, oout = orc.process(fr, ref.ObjInArgs(*args))
            assert ok == 1
            if orc.last_flags():
                continue                                   # see test_sensor_frames: the reference is not run here
            ret, out, _ = rs.process(fr, ref.ObjInArgs(*args))
            assert ret == 0
            assert ref.struct_bytes(out) == ref.struct_bytes(oout), (seed, args)
            nontrivial += 1
    assert nontrivial >= 12


def test_line_sensor_first_frame_lag():
    """OL judges a frame with the PREVIOUS frame's cross band; a fresh object starts at 0,0."""
    w, h = 320, 240
    rs = ref.RefSensor("ol")
    assert rs.setup(w, h)[0] == 0
    orc = ref.OracleSensor("ol", w, h)
    f0 = synth.make_frame("halves", 0, w, h, "yuv422p")
    fr = ref.aligned_bytes(f0.size)
    fr[:] = f0
    ia = ref.RangeInArgs(0, 359, 0, 100, 0, 40, 0)
    ys = []
    for _ in range(3):
        ret, out, _ = rs.process(fr, ia)
        ok, oout = orc.process(fr, ia)
        assert ref.struct_bytes(out)[:3] == ref.struct_bytes(oout)[:3]
        ys.append(out.targetY)
    assert ys[0] != ys[1] and ys[1] == ys[2]


def test_glibc_rand_restatement():
    """trik_oracle_rand == libc srand()/rand() (glibc TYPE_3), for several seeds."""
    libc = C.CDLL(None)
    lib = ref.port_lib()
    st = (C.c_uint32 * 34)()
    for seed in (0, 1, 2, 4242, 0x7FFFFFFF, 0xFFFFFFFF, 123456789):
        libc.srand(C.c_uint(seed))
        lib.trik_oracle_srand(st, C.c_uint(seed))
        for _ in range(2000):
            assert libc.rand() == lib.trik_oracle_rand(st)


# ---------------------------------------------------------------------------------------------
# ov7670/edge_line_sensor: the restatement of the whole sensor against the reference's own sensor code
# built with the open IMGLIB restatement (oracle/imglib_open.c) -- the IMGLIB arithmetic itself stays unpinned
# ---------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(ref.__file__), "_ref", "libtrikref_oe.so")),
                    reason="oracle/_ref/libtrikref_oe.so not built")
@pytest.mark.parametrize("size", [(320, 240), (160, 120), (64, 8), (320, 100)])
def test_edge_line_restatement_equals_reference_sensor_code(size):
    w, h = size
    rs = ref.RefSensor("oe")
    # the reference's work buffer is a file-scope static that is never cleared: rows the Sobel does not write keep what an
    # earlier geometry left there.  An all-zero frame at the largest geometry wipes it, as in a fresh process.
    assert rs.setup(320, 240)[0] == 0
    zero = ref.aligned_bytes(3 * 320 * 240)
    zero[:] = 0
    assert rs.process(zero, ref.RangeInArgs(0, 359, 0, 100, 0, 100, 0), num_bytes=2 * 320 * 240)[0] == 0
    assert rs.setup(w, h)[0] == 0
    fams = [("scene", s) for s in range(8)] + [("noise", 0), ("blobs", 1), ("camera", 2)] + [(e, 0) for e in synth.EDGE_CASES]
    buf = ref.aligned_bytes(3 * w * h)          # the reference reads 2*W*H bytes from the chroma offset (:160-168)
    seen = set()
    for fam, seed in fams:
        f = synth.make_frame(fam, seed, w, h, "yuv422p")
        buf[:] = 0
        buf[:f.size] = f
        ret, out, _ = rs.process(buf, ref.RangeInArgs(0, 359, 0, 100, 0, 100, 0), num_bytes=2 * w * h)
        e = ref.edge_line(f, w, h)
        assert ret == 0 and (out.targetX, out.targetY, out.targetSize) == (e.targetX, e.targetY, e.targetSize), (size, fam, seed)
        seen.add((e.targetX, e.targetY, e.targetSize))
    assert len(seen) > 3
