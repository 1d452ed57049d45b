#!/usr/bin/env python
"""Agreement and cost of the device annealing tail (TRIKB200_BATCH_DEVICE_TAIL) against the host tail: the same
frames and seeds through both, per sensor; prints the number of frames whose result records differ."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w, h = 320, 240
for kind in ("wl", "ol", "oo"):
    layout = "yuyv" if kind == "wl" else "yuv422p"
    uniq = [synth.make_frame(f, s, w, h, layout) for f in ("scene", "noise", "blobs") for s in range(32)]
    frames = np.stack([uniq[i % len(uniq)] for i in range(n)])
    seeds = [(1 + i * 2654435761) % 2147483647 for i in range(n)]      # every frame its own generator state
    ia = xdm.ObjInArgsAlg(1, 0, 40, 60, 40, 60, 40, 1) if kind == "oo" else xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 1)
    codec = open_sensor(kind, w, h)
    res = {}
    for name, flags in (("host", 0), ("device", xdm.BATCH_DEVICE_TAIL)):
        codec.set_params(w, h)
        codec.process_batch(frames[:64], ia, seeds=seeds[:64], flags=flags)          # warm-up (allocations)
        codec.set_params(w, h)
        t0 = time.perf_counter()
        ret, outs = codec.process_batch(frames, ia, seeds=seeds, flags=flags)
        dt = time.perf_counter() - t0
        assert ret == 0, lib().trikb200_lastError()
        res[name] = ([bytes(memoryview(o)) for o in outs], dt)
    differing = sum(1 for a, b in zip(res["host"][0], res["device"][0]) if a != b)
    distinct = len(set(res["host"][0]))
    print(json.dumps({"sensor": kind, "frames": n, "size": "%dx%d" % (w, h), "records_differing": differing,
                      "distinct_results": distinct, "host_tail_s": res["host"][1], "device_tail_s": res["device"][1],
                      "host_threads": os.cpu_count()}), flush=True)
    codec.close()
