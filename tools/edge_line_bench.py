#!/usr/bin/env python
"""Throughput of trikb200_edgeLineBatch (ov7670/edge_line_sensor as a batch operation) with device-resident YUV422P frames,
CUDA events.  Only the luma plane is read: algorithmic bytes = width * height per frame."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from trik_media_sensors_dsp_b200 import lib, synth, xdm  # noqa: E402

peak = 6541.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib().trikb200_setEdgeLineVariant(int(os.environ.get("EDGEVARIANT", "0")))
for (w, h) in ((320, 240), (640, 480)):
    hu = synth.make_batch("scene", range(64), w, h, "yuv422p")
    host = np.concatenate([hu] * (n // 64))
    d_frames = torch.from_numpy(host).cuda()
    d_out = torch.zeros((n, 16), dtype=torch.uint8, device="cuda")
    d = xdm.EdgeLineBatch()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.lineLength = n, w, h, w
    d.framesMem, d.outArgsMem, d.outArgsStride = xdm.MEM_DEVICE, xdm.MEM_DEVICE, 16
    d.frames, d.frameStride, d.outArgsAlg = d_frames.data_ptr(), host.shape[1], d_out.data_ptr()
    for _ in range(3):
        assert lib().trikb200_edgeLineBatch(C.byref(d)) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib().trikb200_edgeLineBatch(C.byref(d))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = n * w * h / (ms / 1e3) / 1e9
    print(json.dumps({"kernel": "edge_line4_kernel" if os.environ.get("EDGEVARIANT", "0") == "0" else "edge_line_kernel", "frames": n, "size": "%dx%d" % (w, h), "ms_per_batch": ms,
                      "frames_per_sec": n / (ms / 1e3), "luma_GBps": gbs, "frac_of_measured_hbm_luma_only": gbs / peak}), flush=True)
