"""The mxn grid sensor through its colour-bin table (csrc/trik_kernels_omtab.cu): the table must hold the oracle's
bin (H>>3, S>>6, V>>6 of hsv(rgb(yuv))) for ALL 2^24 (Y,U,V), and the sensor run through it must give the oracle's
bytes -- including on frames built so that several bins tie at the maximum of a cell, where the reference's "first bin
to reach the final maximum" (ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp:434-440) decides."""
import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu
N24 = 1 << 24


@pytest.fixture(autouse=True)
def _defaults_after():
    yield
    lib().trikb200_setMxnTableMode(0)
    lib().trikb200_setMxnTableThreads(0)


def test_table_holds_the_oracle_bin_for_all_inputs():
    got = np.empty(N24, np.uint32)
    assert lib().trikb200_probePixels(3, 0, N24, got.ctypes.data) == 0, sensors.last_error()
    olib = oracle.port_lib()
    rgb = np.empty(N24, np.uint32)
    olib.trik_oracle_yuv_to_rgb888_range(0, N24, rgb.ctypes.data)          # index = Y | U << 8 | V << 16
    hsv_of_rgb = np.empty(N24, np.uint32)
    olib.trik_oracle_rgb888_to_hsv_range(0, N24, hsv_of_rgb.ctypes.data)
    hsv = hsv_of_rgb[rgb]
    # GetImgColor2 (:425-432): bin = (H >> 3, S >> 6, V >> 6) of the 0x00VVSSHH word
    want = ((hsv & 0xFF) >> 3 << 4) | (((hsv >> 8) & 0xFF) >> 6 << 2) | ((hsv >> 16) >> 6)
    assert np.array_equal(got, want)


GRIDS = [(3, 3), (5, 5), (1, 1), (2, 7), (10, 10), (1, 100), (100, 1), (4, 25), (7, 3), (1, 13), (2, 30)]


def _check(codec, orc, frames, m, n, tag):
    want = []
    for i in range(frames.shape[0]):
        ok, exp = orc.process(frames[i], oracle.MxnInArgs(m, n))
        assert ok == 1
        want.append(list(exp.outColor[:m * n]))
    for mode in (0, 1, -1):            # majority pass + histogram rest (default) / histogram table only / arithmetic
        lib().trikb200_setMxnTableMode(mode)
        ret, outs = codec.process_batch(frames, xdm.MxnInArgsAlg(m, n))
        assert ret == 0, sensors.last_error()
        for i in range(frames.shape[0]):
            got = list(outs[i].outColor[:m * n])
            assert got == want[i], (tag, mode, m, n, i, [j for j in range(m * n) if got[j] != want[i][j]][:8])
            assert all(v == 0 for v in outs[i].outColor[m * n:])


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 4), (96, 8), (1280, 720)])
def test_mxn_through_table_matches_oracle(size):
    w, h = size
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    for (m, n) in GRIDS:
        if m > h or n > w:
            continue
        fams = [("noise", 1), ("noise", 2), ("scene", 2), ("zero", 0), ("greyramp", 0), ("bluewrap", 0), ("checker", 0),
                ("halves", 0), ("full", 0)]
        if w * h > 640 * 480:
            fams = fams[:3] + fams[4:6]
        frames = [synth.make_frame(f, s, w, h, "yuv422p") for f, s in fams]
        frames += [synth.make_frame("grid", s, w, h, "yuv422p", m=m, n=n) for s in range(3)]
        _check(codec, orc, np.stack(frames), m, n, size)
    codec.close()


def _two_colour_frames(w, h):
    """Frames in which two colours own exactly the same number of pixels of every cell of several grids, laid out so
    that now one, now the other reaches the final count first."""
    a, b = (60, 90, 200), (200, 180, 60)                    # (Y, U, V): different colour bins
    out = []
    for first, second in ((a, b), (b, a)):
        for layout in ("cols", "rows", "rowpairs", "colpairs", "quarters"):
            y = np.empty((h, w), np.uint8); u = np.empty((h, w // 2), np.uint8); v = np.empty((h, w // 2), np.uint8)
            rows = np.arange(h)[:, None]
            cols2 = np.arange(w // 2)[None, :]
            if layout == "cols":            # left half / right half of every 32 columns
                sel2 = ((cols2 * 2) % 32 >= 16) & (rows >= 0)
            elif layout == "rows":          # alternating rows
                sel2 = (rows % 2 == 1) & (cols2 >= 0)
            elif layout == "rowpairs":      # two rows on, two rows off
                sel2 = ((rows // 2) % 2 == 1) & (cols2 >= 0)
            elif layout == "colpairs":      # alternating pixel pairs
                sel2 = (cols2 % 2 == 1) & (rows >= 0)
            else:                           # the second colour leads in the upper half, trails in the lower
                sel2 = ((cols2 % 2 == 1) ^ (rows >= h // 2))
            sel = np.repeat(sel2, 2, axis=1)
            y[:] = np.where(sel, second[0], first[0])
            u[:] = np.where(sel2, second[1], first[1])
            v[:] = np.where(sel2, second[2], first[2])
            out.append(synth.pack(y, u, v, "yuv422p"))
    return np.stack(out)


@pytest.mark.parametrize("size", [(320, 240), (64, 32), (640, 480)])
def test_ties_at_the_maximum_follow_the_reference(size):
    w, h = size
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    frames = _two_colour_frames(w, h)
    for (m, n) in [(1, 1), (2, 2), (4, 4), (2, 5), (3, 3), (1, 10), (8, 2)]:
        _check(codec, orc, frames, m, n, ("ties", size))
    codec.close()


def test_large_batches_take_the_table_on_their_own_and_agree():
    w, h = 320, 240
    codec = open_sensor("om", w, h)
    frames = np.stack([synth.make_frame("grid", s % 11, w, h, "yuv422p", m=3, n=3) for s in range(96)])
    res = []
    for mode, threads in ((0, 0), (-1, 0), (1, 96), (1, 512)):
        lib().trikb200_setMxnTableMode(mode)
        lib().trikb200_setMxnTableThreads(threads)
        ret, outs = codec.process_batch(frames, xdm.MxnInArgsAlg(3, 3))
        assert ret == 0, sensors.last_error()
        res.append([list(o.outColor[:9]) for o in outs])
    assert res[0] == res[1] == res[2] == res[3]
    codec.close()


def test_per_frame_grids_in_one_batch_through_the_table():
    w, h = 320, 240
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    grids = [(3, 3), (5, 5), (2, 7), (1, 1), (10, 10), (3, 3), (1, 20), (20, 1)]
    frames = np.stack([synth.make_frame("grid", s, w, h, "yuv422p", m=g[0], n=g[1]) for s, g in enumerate(grids)])
    ias = (xdm.MxnInArgsAlg * len(grids))(*[xdm.MxnInArgsAlg(*g) for g in grids])
    for mode in (0, 1):
        lib().trikb200_setMxnTableMode(mode)
        ret, outs = codec.process_batch(frames, ias)
        assert ret == 0
        for i, (m, n) in enumerate(grids):
            ok, exp = orc.process(frames[i], oracle.MxnInArgs(m, n))
            assert list(outs[i].outColor[:m * n]) == list(exp.outColor[:m * n]), (mode, i, m, n)
    codec.close()


def _majority_edge_frames(w, h, m, n):
    """Cells in which the leading colour holds exactly half, one pixel more than half, and one pixel less than half of
    the pixels (the majority pass may only decide the second), the rest one other colour or scattered colours."""
    a, b = (60, 90, 200), (200, 180, 60)
    hs, ws = h // m, w // n
    out = []
    for delta in (0, 1, -1, 2):
        for scatter in (False, True):
            y = np.full((h, w), b[0], np.uint8); u = np.full((h, w // 2), b[1], np.uint8); v = np.full((h, w // 2), b[2], np.uint8)
            if scatter:
                nz = synth.planes_noise(7 + delta, w, h)
                y, u, v = nz[0].copy(), nz[1].copy(), nz[2].copy()
            for i in range(m):
                for j in range(n):
                    r0, c0 = i * hs, (j * ws + 1) // 2 * 2                  # chroma pairs wholly inside the cell
                    c1 = ((j + 1) * ws) // 2 * 2
                    want = (hs * ws) // 2 + delta                           # pixels of colour a, whole pairs only
                    pairs = max(0, min(want // 2, hs * ((c1 - c0) // 2)))
                    k = 0
                    for r in range(r0, r0 + hs):
                        for c in range(c0, c1, 2):
                            if k >= pairs:
                                break
                            y[r, c] = a[0]; y[r, c + 1] = a[0]; u[r, c // 2] = a[1]; v[r, c // 2] = a[2]
                            k += 1
            out.append(synth.pack(y, u, v, "yuv422p"))
    return np.stack(out)


@pytest.mark.parametrize("size,grid", [((320, 240), (3, 3)), ((640, 480), (3, 3)), ((160, 120), (2, 5)), ((320, 240), (4, 7))])
def test_majority_pass_only_decides_what_it_can_prove(size, grid):
    w, h = size
    m, n = grid
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    _check(codec, orc, _majority_edge_frames(w, h, m, n), m, n, ("majority-edge", size))
    cam = np.stack([synth.make_frame("camera", s, w, h, "yuv422p", m=m, n=n) for s in range(3)])
    _check(codec, orc, cam, m, n, ("camera", size))
    codec.close()


def test_majority_pass_one_cta_per_frame_and_back_to_back():
    """Batches of >= 592 frames run the majority pass with one CTA per frame (all cell rows); the list of undecided cell
    rows must be handed back empty, so a second and third batch on the same handle see no leftovers."""
    w, h = 160, 120
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    fams = [("grid", s) for s in range(10)] + [("noise", 1), ("scene", 3), ("halves", 0), ("camera", 2)]
    for (m, n) in [(3, 3), (5, 5), (2, 7), (7, 1)]:
        distinct = [synth.make_frame(f, s, w, h, "yuv422p", **({"m": m, "n": n} if f in ("grid", "camera") else {})) for f, s in fams]
        want = []
        for fr in distinct:
            ok, exp = orc.process(fr, oracle.MxnInArgs(m, n))
            want.append(list(exp.outColor[:m * n]))
        frames = np.stack([distinct[i % len(distinct)] for i in range(640)])
        lib().trikb200_setMxnTableMode(0)
        for rep in range(3):
            ret, outs = codec.process_batch(frames, xdm.MxnInArgsAlg(m, n))
            assert ret == 0, sensors.last_error()
            for i in range(640):
                assert list(outs[i].outColor[:m * n]) == want[i % len(distinct)], (m, n, rep, i)
    codec.close()
