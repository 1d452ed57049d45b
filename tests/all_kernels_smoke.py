#!/usr/bin/env python
"""TEST HELPER: a small run of EVERY kernel (all sensors, auto-detect, preview, several CTAs per frame)
checked against the oracle; usable stand-alone (`python tests/all_kernels_smoke.py`, e.g. under a
sanitizer where one is available) and from tests/test_all_kernels_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oracle  # noqa: E402
from trik_media_sensors_dsp_b200 import open_sensor, sensors, synth, xdm, lib  # noqa: E402


def out_bytes(o, n):
    return bytes(memoryview(o))[:n]


def main():
    checked = 0
    for (w, h, ow, oh) in [(96, 8, 96, 8), (160, 120, 80, 60), (320, 240, 320, 240)]:
        for kind in xdm.KIND_NAMES:
            layout = sensors.layout_of(xdm.KIND_OF[kind])
            codec = open_sensor(kind, w, h, out_w=ow, out_h=oh)
            orc = oracle.OracleSensor(kind, w, h)
            fams = [("noise", 1), ("scene", 2), ("blobs", 3)]
            frames = np.stack([synth.make_frame(f, s, w, h, layout) for f, s in fams])
            auto = 1 if w >= 160 else 0
            if kind == "oo":
                a = (1, 120, 25, 60, 35, 55, 40, auto)
            elif kind == "om":
                a = (3, 4)
            else:
                a = (0, 359, 0, 100, 0, 45, auto)
            InAlg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]]
            prev = np.zeros((len(fams), oh * ow * 2), dtype=np.uint8)
            for slabs in (0, 3):
                lib().trikb200_setSlabsPerFrame(slabs)
                assert codec.set_params(w, h, out_w=ow, out_h=oh) == 0
                ret, outs = codec.process_batch(frames, InAlg(*a), seeds=[5] * len(fams), previews=prev)
                assert ret == 0, sensors.last_error()
                orc = oracle.OracleSensor(kind, w, h)
                for i in range(len(fams)):
                    ok, exp = orc.process(frames[i], oracle.IN_ARGS[kind](*a), seed=5)
                    n = {"om": 48, "oo": 24}.get(kind, 3)
                    assert out_bytes(outs[i], n) == out_bytes(exp, n), (kind, w, h, fams[i])
                    checked += 1
            lib().trikb200_setSlabsPerFrame(0)
            r, oa = codec.process(frames[0], InAlg(*a), seed=5)
            assert r == 0
            codec.close()
    print("sanitize smoke ok:", checked, "frames checked")


if __name__ == "__main__":
    main()
