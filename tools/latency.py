#!/usr/bin/env python
"""Latency of ONE reference-style process() call (host frame in, result + RGB565X preview out, everything
synchronous) per sensor: the number a caller that switches over frame by frame sees.  Structs are built once,
the loop only makes the C call."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, sensors  # noqa: E402

ARGS = {"wo": (300, 40, 20, 100, 30, 100, 0), "wl": (0, 359, 0, 100, 0, 40, 0), "ol": (0, 359, 0, 100, 0, 40, 0),
        "oo": (1, 0, 40, 60, 40, 60, 40, 0), "om": (3, 3)}
sizes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1] if len(sys.argv) > 1 else "320x240,640x480").split(",")]
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 400
pinned = len(sys.argv) > 3 and sys.argv[3] == "pinned"
sensors.lib().trikb200_setZeroCopyBytes(int(os.environ.get("ZEROCOPY", str(1 << 20))))
sensors.lib().trikb200_setMxnTableMode(int(os.environ.get("OMTAB", "0")))
only = os.environ.get("KINDS", "").split(",") if os.environ.get("KINDS") else None
if pinned:
    import torch  # noqa: E402
for w, h in sizes:
    for kind in xdm.KIND_NAMES:
        if only and kind not in only:
            continue
        layout = sensors.layout_of(xdm.KIND_OF[kind])
        frame = synth.make_frame("blobs" if kind == "oo" else ("grid" if kind == "om" else "scene"), 1, w, h, layout)
        codec = open_sensor(kind, w, h)
        preview = codec.preview
        if pinned:                                     # caller-side pinned buffers (cudaHostAlloc)
            keep = (torch.empty(frame.nbytes, dtype=torch.uint8, pin_memory=True), torch.empty(preview.nbytes, dtype=torch.uint8, pin_memory=True))
            keep[0].numpy()[:] = frame
            frame, preview = keep[0].numpy(), keep[1].numpy()
        in_bufs = xdm.XDM1_BufDesc()
        in_bufs.numBufs = 1
        in_bufs.descs[0].buf = frame.ctypes.data
        in_bufs.descs[0].bufSize = frame.nbytes
        ptrs = (C.c_void_p * 1)(preview.ctypes.data)
        szs = (C.c_int32 * 1)(preview.nbytes)
        out_bufs = xdm.XDM_BufDesc(ptrs, 1, szs)
        ia = codec.InArgs()
        ia.base.size = C.sizeof(ia)
        ia.base.numBytes = frame.nbytes
        ia.base.inputID = 1
        ia.alg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]](*ARGS[kind])
        oa = codec.OutArgs()
        oa.base.size = C.sizeof(oa)
        for _ in range(20):
            assert codec.process_raw(in_bufs, out_bufs, ia, oa) == 0
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            codec.process_raw(in_bufs, out_bufs, ia, oa)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        print(json.dumps({"sensor": kind, "size": "%dx%d" % (w, h), "calls": calls, "zero_copy_bytes": int(os.environ.get("ZEROCOPY", str(1 << 20))), "buffers": "pinned" if pinned else "pageable", "us_median": float(np.median(ts)),
                          "us_p10": float(np.percentile(ts, 10)), "us_p90": float(np.percentile(ts, 90)),
                          "calls_per_sec": 1e6 / float(np.median(ts)),
                          "preview_bytes": int(preview.nbytes), "frame_bytes": int(frame.nbytes)}), flush=True)
        codec.close()
