#!/usr/bin/env python
"""BASELINE.json config 3: the mxn grid colour sensor (3x3 and 5x5 cells) on a batch of 640x480 YUV422P frames, with the
auto-detect HSV calibration exercised beside it (the mxn sensor itself has none: SURVEY 8(d)) through the webcam object
sensor (deterministic) and the ov7670 object / line sensors (seeded annealing).

    python tools/config3_mxn.py [frames=1024] [check=64]

Per line: frames/s with the frames resident in HBM (CUDA events), frames/s through the host-buffer batch call
(H2D + kernel + D2H inside the timed region), and the number of frames among the first `check` whose records differ
from the oracle (must be 0).  The oracle is only the checker here."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    check = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    import torch
    from oracle import ref as oracle
    from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, sensors
    w, h = 640, 480
    peak = 6541.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)

    # ---- the grid sensor ---------------------------------------------------------------------------------------
    for (m, g) in ((3, 3), (5, 5)):
        uniq = 64
        hu = synth.make_batch("grid", range(uniq), w, h, "yuv422p", m=m, n=g)
        host = torch.empty((n, hu.shape[1]), dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for i in range(0, n, uniq):
            hv[i:i + uniq] = hu[:min(uniq, n - i)]
        d_frames = host.to(dev)
        rec = C.sizeof(xdm.OUT_ARGS_ALG[xdm.KIND_OF["om"]])
        d_out = torch.zeros((n, rec), dtype=torch.uint8, device=dev)
        codec = open_sensor("om", w, h)
        ia = xdm.MxnInArgsAlg(m, g)

        def resident():
            ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=hv.shape[1], num_frames=n,
                                         out_device_ptr=d_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC)
            assert ret == 0, sensors.last_error()

        for _ in range(3):
            resident()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 20
        e0.record(stream)
        for _ in range(steps):
            resident()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        # host buffers: H2D + kernel + D2H inside the call
        outs = None
        codec.process_batch(hv, ia)
        t0 = time.perf_counter()
        for _ in range(3):
            ret, outs = codec.process_batch(hv, ia)
            assert ret == 0, sensors.last_error()
        e2e = 3 * n / (time.perf_counter() - t0)
        orc = oracle.OracleSensor("om", w, h)
        bad = 0
        for i in range(min(check, n)):
            ok, exp = orc.process(hv[i], oracle.MxnInArgs(m, g))
            bad += int(ok != 1 or list(outs[i].outColor[:m * g]) != list(exp.outColor[:m * g]))
        print(json.dumps({"config": 3, "sensor": "om", "grid": "%dx%d" % (m, g), "size": "%dx%d" % (w, h), "frames": n,
                          "resident_frames_per_sec": n / (ms / 1e3), "ms_per_batch": ms,
                          "frac_of_measured_hbm": n * w * h * 2 / (ms / 1e3) / 1e9 / peak,
                          "host_buffers_frames_per_sec": e2e, "checked_against_oracle": min(check, n), "mismatches": bad}),
              flush=True)
        codec.close()
        del d_frames, d_out

    # ---- auto-detect HSV beside it -----------------------------------------------------------------------------
    sub = min(n, 256)
    for kind in ("wo", "oo", "ol", "wl"):
        layout = sensors.layout_of(xdm.KIND_OF[kind])
        frames = synth.make_batch("scene", range(sub), w, h, layout)
        seeds = [(7 + i * 2654435761) % 2147483647 for i in range(sub)]
        ia = xdm.ObjInArgsAlg(1, 0, 40, 60, 40, 60, 40, 1) if kind == "oo" else xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 1)
        codec = open_sensor(kind, w, h)
        codec.process_batch(frames, ia, seeds=seeds)                     # warm-up at full size (allocations, tables, threads)
        codec.set_params(w, h)                                           # fresh carried state, as the oracle below starts from
        t0 = time.perf_counter()
        ret, outs = codec.process_batch(frames, ia, seeds=seeds)
        dt = time.perf_counter() - t0
        assert ret == 0, sensors.last_error()
        orc = oracle.OracleSensor(kind, w, h)
        oia = oracle.ObjInArgs(1, 0, 40, 60, 40, 60, 40, 1) if kind == "oo" else oracle.RangeInArgs(0, 359, 0, 100, 0, 100, 1)
        bad = 0
        nchk = min(check, sub) if kind == "wo" else min(check // 4, sub)      # the annealed ones cost ~10 ms each on the CPU
        for i in range(nchk):
            ok, exp = orc.process(frames[i], oia, seed=seeds[i])
            got = bytes(memoryview(outs[i]))
            want = bytes(memoryview(exp))
            # the calibrated range: the six uint16 at the end of the record
            bad += int(ok != 1 or got[-12:] != want[-12:])
        print(json.dumps({"config": 3, "sensor": kind, "auto_detect_hsv": True, "size": "%dx%d" % (w, h), "frames": sub,
                          "host_buffers_frames_per_sec": sub / dt, "checked_against_oracle": nchk, "mismatches": bad}),
              flush=True)
        codec.close()


if __name__ == "__main__":
    main()
