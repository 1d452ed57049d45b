#!/usr/bin/env python
"""Effect of CTAs-per-frame (slabs) on the sum kernels: device-resident, CUDA events."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, sensors, lib  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "wl"
w, h = (int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "320x240").split("x"))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)
layout = sensors.layout_of(xdm.KIND_OF[kind])
fbytes = synth.frame_bytes(w, h, layout)
hu = synth.make_batch("scene", range(64), w, h, layout)
host = torch.empty((n, fbytes), dtype=torch.uint8)
for i in range(0, n, 64):
    host.numpy()[i:i + 64] = hu[:min(64, n - i)]
d_frames = host.to(dev)
d_out = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
codec = open_sensor(kind, w, h)
ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)
SLABS = [int(x) for x in os.environ.get("SLABS", "0").split(",")]
STAGES = [int(x) for x in os.environ.get("STAGES", "-1,100,102,104,200,202,204").split(",")]
THREADS = [int(x) for x in os.environ.get("THREADS", "0").split(",")]
lib().trikb200_setOverlapLaunch(int(os.environ.get("OVERLAP", "1")))
lib().trikb200_setFramesPerCta(int(os.environ.get("FPC", "0")))
ref_out = None
for stages, slabs, threads in [(a, b, c) for a in STAGES for b in SLABS for c in THREADS]:
    lib().trikb200_setBlockThreads(threads)
    lib().trikb200_setSlabsPerFrame(slabs)
    lib().trikb200_setLoadStages(stages)

    def step():
        ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                     out_device_ptr=d_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC)
        assert ret == 0
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    cur = d_out.cpu().numpy()[:, :3].tobytes()
    if ref_out is None:
        ref_out = cur
    assert cur == ref_out, "results changed with the tuning knobs"
    print(json.dumps({"sensor": kind, "size": "%dx%d" % (w, h), "batch": n, "slabs": slabs, "stages": stages, "threads": threads, "overlap": int(os.environ.get("OVERLAP", "1")), "fpc": int(os.environ.get("FPC", "0")), "ms": best,
                      "GBps": n * w * h * 2 / best / 1e6}), flush=True)
