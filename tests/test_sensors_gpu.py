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
