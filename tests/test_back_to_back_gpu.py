"""Asynchronous batches issued back to back on one handle and one stream, no synchronisation in between.

The line kernels are launched with programmatic stream serialisation (the next batch's CTAs become resident while the
previous batch drains, csrc/trik_kernels.cu) and every sensor keeps per-handle scratch on the device (slab accumulators,
OO metapixel bitmaps and label tables, the detection table).  Batches of DIFFERENT frames -- slow ones followed by quick
ones, so that a batch really is still running when the next one is issued -- must each give the oracle's bytes, with the
overlapped launch on and off."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu

SETS = [[("noise", 100 + i) for i in range(8)] + [("scene", i) for i in range(8)],
        [("blobs", i) for i in range(12)] + [("zero", 0), ("full", 0), ("scene", 40), ("noise", 7)],
        [("scene", 20 + i) for i in range(10)] + [("blobs", 30 + i) for i in range(6)]]
ORDER = [0, 1, 2, 1, 0, 2, 2, 0, 1]
CASES = {
    "oo": (oracle.ObjInArgs, xdm.ObjInArgsAlg, (1, 0, 40, 60, 40, 60, 40, 0), 24),
    "wo": (oracle.RangeInArgs, xdm.RangeInArgsAlg, (300, 40, 20, 100, 30, 100, 0), 3),
    "wl": (oracle.RangeInArgs, xdm.RangeInArgsAlg, (0, 359, 0, 100, 0, 40, 0), 3),
    "ol": (oracle.RangeInArgs, xdm.RangeInArgsAlg, (0, 359, 0, 100, 0, 40, 0), 3),
}


@pytest.mark.parametrize("kind,size,n", [("oo", (320, 240), 288), ("oo", (160, 120), 288), ("wo", (320, 240), 288),
                                         ("wl", (320, 240), 48), ("wl", (320, 240), 2048), ("ol", (320, 240), 48),
                                         ("ol", (640, 480), 512)])
def test_back_to_back_async_batches_match_oracle(kind, size, n):
    import torch
    w, h = size
    OIn, XIn, args, nbytes = CASES[kind]
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    orc = oracle.OracleSensor(kind, w, h)
    batches, wants = [], []
    for fams in SETS:
        uniq = np.stack([synth.make_frame(f, s, w, h, layout) for f, s in fams])
        if kind == "ol":
            orc.process(uniq[0], OIn(*args))            # OL: the first call of a fresh object reads unset members
        one = [bytes(memoryview(orc.process(uniq[i], OIn(*args))[1]))[:nbytes] for i in range(len(fams))]
        reps = n // len(fams)
        batches.append(np.concatenate([uniq] * reps))
        wants.append(one * reps)
    rec = C.sizeof(xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]])
    stream = torch.cuda.Stream()
    sptr = C.c_void_p(stream.cuda_stream)
    d_frames = [torch.from_numpy(b).cuda() for b in batches]
    codec = open_sensor(kind, w, h)
    try:
        if kind == "ol":
            ret, _ = codec.process_batch(batches[0][:1], XIn(*args))
            assert ret == 0, lib().trikb200_lastError()
        for overlap in (1, 0):
            lib().trikb200_setOverlapLaunch(overlap)
            d_outs = [torch.zeros((n, rec), dtype=torch.uint8, device="cuda") for _ in ORDER]
            torch.cuda.synchronize()
            for d_out, k in zip(d_outs, ORDER):
                ret, _ = codec.process_batch(d_frames[k].data_ptr(), XIn(*args), frames_device=True,
                                             frame_stride=batches[k].shape[1], num_frames=n, out_device_ptr=d_out.data_ptr(),
                                             stream=sptr, flags=xdm.BATCH_ASYNC)
                assert ret == 0, lib().trikb200_lastError()
            stream.synchronize()
            for j, (d_out, k) in enumerate(zip(d_outs, ORDER)):
                got = [bytes(r[:nbytes]) for r in d_out.cpu().numpy()]
                assert got == wants[k], (kind, size, overlap, j, k, [i for i in range(n) if got[i] != wants[k][i]][:8])
    finally:
        lib().trikb200_setOverlapLaunch(1)
        codec.close()
