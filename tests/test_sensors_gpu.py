"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, bit for bit.

Bar: every OutArgsAlg byte the sensor produces is identical to the oracle's (integer / byte work).
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm

pytestmark = pytest.mark.gpu

SIZES = [(320, 240), (640, 480), (160, 120), (32, 4), (96, 8), (1280, 720)]
RANGE_ARGS = [
    (0, 359, 0, 100, 0, 40, 0),       # dark line, the BASELINE config-2 setting
    (300, 40, 20, 100, 30, 100, 0),   # hue wraps through 0
    (10, 200, 0, 60, 20, 90, 0),
    (0, 359, 0, 100, 0, 100, 0),      # everything
    (0, 0, 0, 0, 60, 30, 0),          # empty value range
    (90, 150, 35, 100, 35, 100, 0),   # green-ish
]


def layout_of(kind):
    return "yuyv" if kind in ("wo", "wl") else "yuv422p"


def frames_for(kind, w, h):
    fams = [("noise", s) for s in range(3)] + [("scene", s) for s in range(4)] + [(e, 0) for e in synth.EDGE_CASES]
    if w < 64 or h < 16:
        fams = [("noise", s) for s in range(4)] + [("zero", 0), ("full", 0), ("bluewrap", 0), ("greyramp", 0)]
    return fams, np.stack([synth.make_frame(f, s, w, h, layout_of(kind)) for f, s in fams])


def out_bytes(o, n=3):
    return bytes(memoryview(o))[:n]


@pytest.mark.parametrize("kind", ["wl", "wo", "ol"])
@pytest.mark.parametrize("size", SIZES)
def test_sum_sensors_match_oracle(kind, size):
    w, h = size
    fams, frames = frames_for(kind, w, h)
    codec = open_sensor(kind, w, h)
    for args in RANGE_ARGS:
        ia = xdm.RangeInArgsAlg(*args)
        orc = oracle.OracleSensor(kind, w, h)      # fresh object: carried state starts at zero on both sides
        assert codec.set_params(w, h) == 0          # SETPARAMS re-creates the algorithm object
        ret, outs = codec.process_batch(frames, ia)
        assert ret == 0
        for i in range(frames.shape[0]):
            ok, exp = orc.process(frames[i], oracle.RangeInArgs(*args))
            assert ok == 1
            assert out_bytes(outs[i]) == out_bytes(exp), (kind, size, args, fams[i], out_bytes(outs[i]).hex(), out_bytes(exp).hex())
    codec.close()


@pytest.mark.parametrize("kind", ["wl", "wo", "ol"])
def test_single_process_equals_batch(kind):
    """n sequential process() calls == one processBatch() of n frames (layer 4)."""
    w, h = 320, 240
    fams, frames = frames_for(kind, w, h)
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 45, 0)
    codec = open_sensor(kind, w, h)
    ret, outs = codec.process_batch(frames, ia)
    assert ret == 0
    assert codec.set_params(w, h) == 0
    for i in range(frames.shape[0]):
        r, oa = codec.process(frames[i], ia)
        assert r == 0
        assert out_bytes(oa.alg) == out_bytes(outs[i]), (kind, fams[i])
        assert oa.base.bitsConsumed == frames[i].nbytes * 8
        assert oa.base.outputID[0] == 1
    codec.close()


def test_slabs_do_not_change_results():
    """Splitting a frame over several CTAs (atomics + last-CTA finalisation) is bit-exact."""
    from trik_media_sensors_dsp_b200 import lib
    w, h = 640, 480
    frames = np.stack([synth.make_frame("scene", s, w, h, "yuyv") for s in range(6)])
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 50, 0)
    base = None
    for kind in ("wl", "wo"):
        codec = open_sensor(kind, w, h)
        res = []
        for slabs in (1, 2, 3, 5, 8):
            lib().trikb200_setSlabsPerFrame(slabs)
            for _ in range(2):                       # twice: the accumulators must have been reset
                ret, outs = codec.process_batch(frames, ia)
                assert ret == 0
                res.append(b"".join(out_bytes(o) for o in outs))
        lib().trikb200_setSlabsPerFrame(0)
        assert all(r == res[0] for r in res), kind
        codec.close()


# ---------------------------------------------------------------------------------------------
# mxn grid colour sensor
# ---------------------------------------------------------------------------------------------
GRIDS = [(3, 3), (5, 5), (1, 1), (2, 7), (10, 10), (1, 100), (100, 1), (4, 25), (7, 3)]


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 4)])
def test_mxn_matches_oracle(size):
    w, h = size
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    for (m, n) in GRIDS:
        if m > h or n > w:
            continue
        fams = [("noise", 1), ("scene", 2), ("zero", 0), ("greyramp", 0), ("bluewrap", 0)]
        frames = [synth.make_frame(f, s, w, h, "yuv422p") for f, s in fams]
        frames += [synth.make_frame("grid", s, w, h, "yuv422p", m=m, n=n) for s in range(3)]
        frames = np.stack(frames)
        ia = xdm.MxnInArgsAlg(m, n)
        ret, outs = codec.process_batch(frames, ia)
        assert ret == 0
        for i in range(frames.shape[0]):
            ok, exp = orc.process(frames[i], oracle.MxnInArgs(m, n))
            assert ok == 1
            got = list(outs[i].outColor[:m * n])
            want = list(exp.outColor[:m * n])
            assert got == want, (size, m, n, i)
            assert all(v == 0 for v in outs[i].outColor[m * n:])      # untouched entries keep the caller's bytes
    codec.close()


def test_mxn_rejects_undefined_grids():
    codec = open_sensor("om", 320, 240)
    frames = np.stack([synth.make_frame("noise", 0, 320, 240, "yuv422p")])
    for (m, n) in [(0, 3), (3, 0), (11, 10), (101, 1)]:
        ret, _ = codec.process_batch(frames, xdm.MxnInArgsAlg(m, n))
        assert ret == xdm.XDM_EFAIL
    codec.close()


def test_mxn_per_frame_grids_in_one_batch():
    w, h = 320, 240
    codec = open_sensor("om", w, h)
    orc = oracle.OracleSensor("om", w, h)
    grids = [(3, 3), (5, 5), (2, 7), (1, 1), (10, 10), (3, 3)]
    frames = np.stack([synth.make_frame("grid", s, w, h, "yuv422p", m=g[0], n=g[1]) for s, g in enumerate(grids)])
    ias = (xdm.MxnInArgsAlg * len(grids))(*[xdm.MxnInArgsAlg(*g) for g in grids])
    ret, outs = codec.process_batch(frames, ias)
    assert ret == 0
    for i, (m, n) in enumerate(grids):
        ok, exp = orc.process(frames[i], oracle.MxnInArgs(m, n))
        assert list(outs[i].outColor[:m * n]) == list(exp.outColor[:m * n]), (i, m, n)
    codec.close()


# ---------------------------------------------------------------------------------------------
# ov7670 object sensor
# ---------------------------------------------------------------------------------------------
OBJ_ARGS = [
    # setHsvRange, hue, hueTol, sat, satTol, val, valTol, auto
    (1, 0, 20, 80, 20, 50, 30, 0),      # red-ish, the specks / red frames
    (1, 0, 40, 60, 40, 60, 40, 0),
    (0, 0, 0, 0, 0, 0, 0, 0),           # keeps the previous range
    (1, 200, 100, 50, 50, 50, 50, 0),   # broad
    (1, 120, 25, 60, 35, 55, 40, 0),    # sparse on noise: many small labels, exercises the sort
    (1, 350, 30, 50, 50, 50, 50, 0),    # hue wraps through 0
]


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120), (32, 4), (1280, 720)])
def test_object_sensor_matches_oracle(size):
    w, h = size
    fams = [("noise", s) for s in range(4)] + [("scene", s) for s in range(4)] + [(e, 0) for e in synth.EDGE_CASES]
    if w < 64:
        fams = [("noise", s) for s in range(4)] + [("zero", 0), ("red", 0)]
    frames = np.stack([synth.make_frame(f, s, w, h, "yuv422p") for f, s in fams])
    codec = open_sensor("oo", w, h)
    orc = oracle.OracleSensor("oo", w, h)
    # one long batch: the range set by a frame persists for the following setHsvRange == 0 frames
    seq = [(fi, ai) for ai in range(len(OBJ_ARGS)) for fi in range(frames.shape[0])]
    batch = np.stack([frames[fi] for fi, _ in seq])
    ias = (xdm.ObjInArgsAlg * len(seq))(*[xdm.ObjInArgsAlg(*OBJ_ARGS[ai]) for _, ai in seq])
    ret, outs = codec.process_batch(batch, ias)
    assert ret == 0
    multi = 0
    for i, (fi, ai) in enumerate(seq):
        ok, exp = orc.process(frames[fi], oracle.ObjInArgs(*OBJ_ARGS[ai]))
        assert ok == 1
        assert out_bytes(outs[i], 24) == out_bytes(exp, 24), (size, fams[fi], OBJ_ARGS[ai], out_bytes(outs[i], 24).hex(), out_bytes(exp, 24).hex())
        multi += sum(1 for t in exp.target if t.size > 0) > 1
    if w >= 160:
        assert multi > 0            # the vectors do exercise more than one ranked target
    codec.close()


# ---------------------------------------------------------------------------------------------
# auto-calibration
# ---------------------------------------------------------------------------------------------
def detect_fields(o):
    return (o.detectHue, o.detectHueTolerance, o.detectSat, o.detectSatTolerance, o.detectVal, o.detectValTolerance)


@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120)])
def test_wo_autodetect_matches_oracle(size):
    w, h = size
    fams, frames = frames_for("wo", w, h)
    codec = open_sensor("wo", w, h)
    orc = oracle.OracleSensor("wo", w, h)
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 1)
    prefill = (xdm.TargetOutArgsAlg * frames.shape[0])()
    ret, outs = codec.process_batch(frames, ia, out_algs=prefill)
    assert ret == 0
    for i in range(frames.shape[0]):
        ok, exp = orc.process(frames[i], oracle.RangeInArgs(0, 359, 0, 100, 0, 100, 1))
        assert detect_fields(outs[i]) == detect_fields(exp), (size, fams[i])
        assert out_bytes(outs[i]) == out_bytes(exp)
    # without autoDetectHsv the six fields keep the caller's bytes, as with the reference
    marked = (xdm.TargetOutArgsAlg * frames.shape[0])()
    for o in marked:
        o.detectHue, o.detectVal = 1234, 4321
    ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 0), out_algs=marked)
    assert ret == 0 and all(o.detectHue == 1234 and o.detectVal == 4321 for o in outs)
    codec.close()


@pytest.mark.parametrize("kind", ["wl", "ol", "oo"])
@pytest.mark.parametrize("size", [(320, 240), (640, 480), (160, 120)])
def test_annealed_autodetect_matches_oracle(kind, size):
    """Histogram + ordered seed on the GPU, annealing tail on the host with the library's own
    glibc-compatible generator: identical to the oracle (which calls srand/rand/pow) for every seed."""
    w, h = size
    fams = [("noise", 0), ("noise", 1), ("scene", 0), ("scene", 1), ("scene", 2), ("scene", 3), ("halves", 0),
            ("specks", 0), ("checker", 0), ("bluewrap", 0), ("greyramp", 0), ("red", 0)]
    frames = np.stack([synth.make_frame(f, s, w, h, layout_of(kind)) for f, s in fams])
    codec = open_sensor(kind, w, h)
    orc = oracle.OracleSensor(kind, w, h)
    seeds = [1000 + 17 * i for i in range(frames.shape[0])]
    if kind == "oo":
        ia, oia = xdm.ObjInArgsAlg(1, 0, 40, 60, 40, 60, 40, 1), oracle.ObjInArgs(1, 0, 40, 60, 40, 60, 40, 1)
    else:
        ia, oia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 1), oracle.RangeInArgs(0, 359, 0, 100, 0, 40, 1)
    ret, outs = codec.process_batch(frames, ia, seeds=seeds)
    assert ret == 0
    for i in range(frames.shape[0]):
        ok, exp = orc.process(frames[i], oia, seed=seeds[i])
        assert ok == 1
        assert detect_fields(outs[i]) == detect_fields(exp), (kind, size, fams[i], detect_fields(outs[i]), detect_fields(exp))
    # the single-frame process() path with an explicit seed
    r, oa = codec.process(frames[2], ia, seed=seeds[2])
    assert r == 0 and detect_fields(oa.alg) == detect_fields(outs[2])
    codec.close()


# ---------------------------------------------------------------------------------------------
# mixed sensor instances in one call (BASELINE config 4)
# ---------------------------------------------------------------------------------------------
def test_mixed_streams_equal_sequential_process():
    from trik_media_sensors_dsp_b200 import process_mixed
    w, h = 320, 240
    kinds = ["wo", "wl", "ol", "oo", "om"]
    streams = []
    for si in range(15):                                  # 15 streams, kinds cycling, 4 frames each
        kind = kinds[si % 5]
        streams.append((kind, open_sensor(kind, w, h), oracle.OracleSensor(kind, w, h)))

    def args_for(kind, t):
        if kind == "oo":
            a = (1, 0, 20, 80, 20, 50, 30, 0) if t == 0 else (0, 0, 0, 0, 0, 0, 0, 0)    # range persists
            return xdm.ObjInArgsAlg(*a), oracle.ObjInArgs(*a)
        if kind == "om":
            return xdm.MxnInArgsAlg(3 + t % 3, 4), oracle.MxnInArgs(3 + t % 3, 4)
        a = (0, 359, 0, 100, 0, 40 + 10 * t, 1 if (kind == "wo" and t == 1) else 0)
        return xdm.RangeInArgsAlg(*a), oracle.RangeInArgs(*a)

    items, expect, keep = [], [], []
    for t in range(4):                                    # time-major: one frame of every stream per step
        for si, (kind, codec, orc) in enumerate(streams):
            fam = "blobs" if kind == "oo" else ("grid" if kind == "om" else "scene")
            fr = synth.make_frame(fam, 100 * si + t, w, h, layout_of(kind))
            ia, oia = args_for(kind, t)
            oa = xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]]()
            keep.append((fr, ia))
            items.append((codec, fr, ia, oa, 5))
            ok, exp = orc.process(fr, oia, seed=5)
            assert ok == 1
            expect.append((kind, ia, exp))
    assert process_mixed(items) == 0
    for (codec, fr, ia, oa, _), (kind, _, exp) in zip(items, expect):
        if kind == "om":
            n = ia.widthM * ia.heightN
            assert list(oa.outColor[:n]) == list(exp.outColor[:n])
        elif kind == "oo":
            assert out_bytes(oa, 24) == out_bytes(exp, 24)
        else:
            assert out_bytes(oa) == out_bytes(exp), kind
            if ia.autoDetectHsv:
                assert detect_fields(oa) == detect_fields(exp)
    for _, codec, _ in streams:
        codec.close()


@pytest.mark.parametrize("kind", ["ol", "oo"])
def test_logical_streams_in_one_batch(kind):
    """TRIKB200_Batch.streamIds: many logical streams through ONE handle and ONE launch; every stream carries
    its own state (OL cross-band lag, OO persisting range) exactly like its own codec instance."""
    w, h = 320, 240
    nstreams, steps = 7, 4
    codec = open_sensor(kind, w, h)
    oracles = [oracle.OracleSensor(kind, w, h) for _ in range(nstreams)]
    InAlg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]]
    for t in range(steps):                                 # several calls: the state persists across calls too
        frames, ias, ids, exps = [], [], [], []
        order = list(range(nstreams)) if t % 2 == 0 else list(reversed(range(nstreams)))
        for sidx in order + order[:3]:                     # some streams twice within one batch
            fam = "blobs" if kind == "oo" else "halves"
            fr = synth.make_frame(fam, 10 * sidx + t, w, h, "yuv422p")
            if kind == "oo":
                a = (1, 0, 20 + sidx, 80, 20, 50, 30, 0) if (t == 0 or (sidx + t) % 3 == 0) else (0, 0, 0, 0, 0, 0, 0, 0)
            else:
                a = (0, 359, 0, 100, 0, 40 + sidx, 0)
            frames.append(fr)
            ias.append(InAlg(*a))
            ids.append(sidx)
            ok, exp = oracles[sidx].process(fr, oracle.IN_ARGS[kind](*a))
            exps.append(exp)
        arr = (InAlg * len(ias))(*ias)
        ret, outs = codec.process_batch(np.stack(frames), arr, stream_ids=ids, num_streams=nstreams)
        assert ret == 0
        nb = 24 if kind == "oo" else 3
        for o, e in zip(outs, exps):
            assert out_bytes(o, nb) == out_bytes(e, nb), (kind, t)
    codec.close()
