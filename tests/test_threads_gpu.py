"""Host threads driving different handles at the same time (ctypes releases the GIL during the calls):
per-launch state of the library must not leak from one thread's launch into another's."""
import threading

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm

pytestmark = pytest.mark.gpu

JOBS = [("wl", 320, 240, (0, 359, 0, 100, 0, 40, 0)), ("ol", 320, 240, (0, 359, 0, 100, 0, 40, 0)),
        ("wo", 160, 120, (300, 40, 20, 100, 30, 100, 0)), ("ol", 160, 120, (0, 359, 0, 100, 35, 100, 0)),
        ("wl", 640, 44, (0, 359, 0, 100, 35, 100, 0)), ("om", 320, 240, (3, 3))]


def test_concurrent_handles():
    prepared = []
    for kind, w, h, args in JOBS:
        layout = "yuyv" if kind in ("wl", "wo") else "yuv422p"
        frames = np.stack([synth.make_frame("scene" if kind != "om" else "grid", s, w, h, layout) for s in range(6)])
        orc = oracle.OracleSensor(kind, w, h)
        nbytes = 3 if kind != "om" else args[0] * args[1] * 4
        want = [bytes(memoryview(orc.process(frames[i], oracle.IN_ARGS[kind](*args))[1]))[:nbytes] for i in range(6)]
        prepared.append((kind, w, h, args, frames, want, nbytes))

    errors = []
    start = threading.Barrier(len(prepared))

    def worker(kind, w, h, args, frames, want, nbytes):
        try:
            codec = open_sensor(kind, w, h)
            in_alg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]](*args)
            start.wait()
            for rep in range(60):
                assert codec.set_params(w, h) == 0
                ret, outs = codec.process_batch(frames, in_alg)
                assert ret == 0
                got = [bytes(memoryview(o))[:nbytes] for o in outs]
                assert got == want, (kind, w, h, rep)
            codec.close()
        except BaseException as e:                      # noqa: BLE001 -- reported in the main thread
            errors.append((kind, w, h, repr(e)))

    threads = [threading.Thread(target=worker, args=p) for p in prepared]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
