"""Every selectable implementation of the line-sensor pass (trikb200_setLoadStages) must give the oracle's
bytes: first-version kernel, tuned 8-pixel kernel, wide YUV422P kernel, bulk-copy (mbarrier) kernel, at
every ring depth the launcher accepts, with one CTA per frame and with frames split into slabs, on frame
sizes that leave partial iterations and on a padded line length (the bulk kernel's per-row copies)."""
import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, synth, xdm

pytestmark = pytest.mark.gpu

VARIANTS = {
    "wl": [100, 102, 104, 200, 202, 204, 206, 208, 402, 403, 404, 406],
    "ol": [100, 102, 104, 200, 202, 204, 302, 304, 402, 403, 404],
}
ARGS = [(0, 359, 0, 100, 0, 40, 0), (0, 359, 0, 100, 35, 100, 0), (0, 359, 0, 100, 60, 30, 0)]


@pytest.fixture(autouse=True)
def _defaults_after():
    yield
    lib().trikb200_setLoadStages(-1)
    lib().trikb200_setSlabsPerFrame(0)
    lib().trikb200_setBlockThreads(0)


@pytest.mark.parametrize("kind", ["wl", "ol"])
@pytest.mark.parametrize("size", [(320, 240), (640, 44), (160, 120), (32, 4), (1280, 36)])
def test_variants_match_oracle(kind, size):
    w, h = size
    layout = "yuyv" if kind == "wl" else "yuv422p"
    fams = [("noise", 1), ("scene", 2), ("halves", 0), ("bluewrap", 0), ("full", 0), ("noise", 5), ("scene", 6)]
    frames = np.stack([synth.make_frame(f, s, w, h, layout) for f, s in fams])
    want = {}
    for args in ARGS:
        orc = oracle.OracleSensor(kind, w, h)
        want[args] = [bytes(memoryview(orc.process(frames[i], oracle.RangeInArgs(*args))[1]))[:3] for i in range(len(fams))]
    codec = open_sensor(kind, w, h)
    for variant in VARIANTS[kind]:
        for slabs in (1, 3):
            lib().trikb200_setLoadStages(variant)
            lib().trikb200_setSlabsPerFrame(slabs)
            for args in ARGS:
                assert codec.set_params(w, h) == 0       # carried state (OL cross band) restarts, as in the oracle object
                ret, outs = codec.process_batch(frames, xdm.RangeInArgsAlg(*args))
                assert ret == 0, (variant, slabs, lib().trikb200_lastError())
                got = [bytes(memoryview(o))[:3] for o in outs]
                assert got == want[args], (kind, size, variant, slabs, args)
    codec.close()


@pytest.mark.parametrize("kind", ["wl", "ol"])
def test_variants_padded_lines(kind):
    """inputLineLength larger than the row: the bulk kernel copies row by row, the others stride."""
    w, h = 320, 60
    layout = "yuyv" if kind == "wl" else "yuv422p"
    row = w * 2 if kind == "wl" else w
    line = row + 64
    planes = 1 if kind == "wl" else 2
    fams = [("noise", 2), ("scene", 1), ("halves", 0)]
    tight = np.stack([synth.make_frame(f, s, w, h, layout) for f, s in fams])
    padded = np.full((len(fams), planes * h * line), 0xA5, dtype=np.uint8)
    for i in range(len(fams)):
        src = tight[i].reshape(planes * h, row)
        padded[i].reshape(planes * h, line)[:, :row] = src
    args = (0, 359, 0, 100, 0, 40, 0)
    orc = oracle.OracleSensor(kind, w, h)
    want = [bytes(memoryview(orc.process(tight[i], oracle.RangeInArgs(*args))[1]))[:3] for i in range(len(fams))]
    codec = open_sensor(kind, w, h, line_length=line)
    for variant in VARIANTS[kind]:
        lib().trikb200_setLoadStages(variant)
        assert codec.set_params(w, h, line_length=line) == 0
        ret, outs = codec.process_batch(padded, xdm.RangeInArgsAlg(*args))
        assert ret == 0, (variant, lib().trikb200_lastError())
        assert [bytes(memoryview(o))[:3] for o in outs] == want, (kind, variant)
    codec.close()
