#!/usr/bin/env python
"""Throughput of the RGB565 -> YUV422P ingest front end (trikb200_ingestRgb565) with device-resident frames, CUDA events."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from trik_media_sensors_dsp_b200 import lib, xdm  # noqa: E402

peak = 6541.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for (w, h) in ((320, 240), (640, 480)):
    nn = n if w == 320 else n // 4
    src = torch.randint(0, 256, (nn, w * h * 2), dtype=torch.uint8, device="cuda")
    dst = torch.empty((nn, w * h * 2), dtype=torch.uint8, device="cuda")
    d = xdm.Ingest()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.pixelFormat = nn, w, h, xdm.PIXEL_RGB565
    d.srcMem = d.dstMem = xdm.MEM_DEVICE
    d.srcLineLength, d.dstLineLength = 2 * w, w
    d.src, d.srcStride, d.dst, d.dstStride = src.data_ptr(), w * h * 2, dst.data_ptr(), w * h * 2
    for _ in range(3):
        assert lib().trikb200_ingestRgb565(C.byref(d)) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib().trikb200_ingestRgb565(C.byref(d))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = nn * w * h * 4 / (ms / 1e3) / 1e9
    print(json.dumps({"kernel": "ingest_rgb565_kernel", "frames": nn, "size": "%dx%d" % (w, h), "ms_per_batch": ms,
                      "frames_per_sec": nn / (ms / 1e3), "read_plus_write_GBps": gbs,
                      "frac_of_measured_copy_peak": gbs / peak}), flush=True)
