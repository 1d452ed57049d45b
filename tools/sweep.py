#!/usr/bin/env python
"""Sensor / frame-size / batch sweep (BASELINE.json configs 2-5), device-resident, CUDA events.

    python tools/sweep.py [--kinds wl,wo,ol,om,oo] [--sizes 320x240,640x480] [--batch 1024] [--steps 20]

One JSON line per (sensor, size): frames/s, ms per batch, achieved algorithmic GB/s (W*H*2 bytes per
frame) and its fraction of the measured HBM peak.  Not the headline bench (that is bench.py)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", default="wl,ol,wo,om,oo")
    ap.add_argument("--sizes", default="320x240,640x480")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--family", default="scene")
    ap.add_argument("--unique", type=int, default=64)
    ap.add_argument("--grid", default="3x3")
    args = ap.parse_args()

    import torch
    from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, sensors, launch_count
    sensors.lib().trikb200_setLutMode(int(os.environ.get("LUT", "0")))      # -1: arithmetic WO kernel only
    sensors.lib().trikb200_setMxnTableMode(int(os.environ.get("OMTAB", "0")))   # -1: arithmetic mxn kernel only
    sensors.lib().trikb200_setMxnTableThreads(int(os.environ.get("OMTHREADS", "0")))
    sensors.lib().trikb200_setLutSkew(int(os.environ.get("SKEW", "1")))
    sensors.lib().trikb200_setPreviewChunkMB(int(os.environ.get("PVCHUNK", "0")))   # MiB of preview images per sub-batch, 0 = all at once
    sensors.lib().trikb200_setPreviewSectorOverlay(int(os.environ.get("PVSECTOR", "-1")))   # -1: fused into the streaming kernel (default), 0: generic overlay kernel, 1: sector kernel
    sensors.lib().trikb200_setPreviewTable(int(os.environ.get("PVLUT", "1")))       # 0: WO preview detects by arithmetic even when the batch has its table
    sensors.lib().trikb200_setLutParts(int(os.environ.get("LUTPARTS", "0")))        # bands per frame of the WO table kernel
    sensors.lib().trikb200_setOverlapLaunch(int(os.environ.get("OVERLAP", "1")))   # 0: no programmatic dependent launches
    peak = 6541.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    gm, gn = [int(x) for x in args.grid.split("x")]
    for kind in args.kinds.split(","):
        for size in args.sizes.split(","):
            w, h = [int(x) for x in size.split("x")]
            k = xdm.KIND_OF[kind]
            layout = sensors.layout_of(k)
            fbytes = synth.frame_bytes(w, h, layout)
            n = args.batch
            while n * fbytes > 8e9:
                n //= 2
            uniq = min(args.unique, n)
            fam = "grid" if (kind == "om" and args.family == "scene") else args.family
            if fam == "grid" or (fam == "camera" and kind == "om"):
                hu = synth.make_batch(fam, range(uniq), w, h, layout, m=gm, n=gn)
            else:
                hu = synth.make_batch(fam, range(uniq), w, h, layout)
            host = torch.empty((n, fbytes), dtype=torch.uint8)
            hv = host.numpy()
            for i in range(0, n, uniq):
                hv[i:i + uniq] = hu[:min(uniq, n - i)]
            d_frames = host.to(dev)
            rec = C.sizeof(xdm.OUT_ARGS_ALG[k])
            d_out = torch.zeros((n, rec), dtype=torch.uint8, device=dev)
            codec = open_sensor(kind, w, h)
            if kind == "oo":
                # OOARGS=wide: a range that the blob / speck frames actually fall into (the default one finds nothing in
                # them, i.e. measures the walk over empty bitmaps)
                ia = (xdm.ObjInArgsAlg(1, 0, 25, 75, 25, 60, 40, 0) if os.environ.get("OOARGS") == "wide"
                      else xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 30, 0))
            elif kind == "om":
                ia = xdm.MxnInArgsAlg(gm, gn)
            else:
                ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)
            # WOSETS=k: the webcam object sensor's frames under k different threshold sets, interleaved (per-frame arguments)
            nsets = int(os.environ.get("WOSETS", "1"))
            if kind == "wo" and nsets > 1:
                ia = (xdm.RangeInArgsAlg * n)(*[xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 30 + 5 * (i % nsets), 0) for i in range(n)])

            # PREVIEW=1: with the RGB565X preview (1:1, device memory) as process() always produces it
            pv = torch.empty((n, fbytes), dtype=torch.uint8, device=dev) if os.environ.get("PREVIEW") == "1" else None
            pkw = {"previews_device_ptr": pv.data_ptr(), "preview_stride": fbytes} if pv is not None else {}

            def step():
                ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                             out_device_ptr=d_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC, **pkw)
                assert ret == 0, sensors.last_error()

            for _ in range(3):
                step()
            torch.cuda.synchronize(dev)
            l0 = launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / args.steps
            gbs = n * w * h * 2 / (ms / 1e3) / 1e9
            print(json.dumps({"sensor": kind, "width": w, "height": h, "batch": n, "family": fam, "ms_per_batch": ms,
                              "frames_per_sec": n / (ms / 1e3), "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak,
                              "launches_per_batch": (launch_count() - l0) / args.steps, "preview": pv is not None, "wo_threshold_sets": nsets if kind == "wo" else None}), flush=True)
            codec.close()
            del d_frames, d_out


if __name__ == "__main__":
    main()
