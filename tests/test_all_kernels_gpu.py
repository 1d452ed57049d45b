"""Every kernel of the library in one small run (all sensors, auto-calibration, previews, several CTAs per
frame) checked against the oracle, and guard regions around caller-provided DEVICE output buffers.
compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are looked for with canaries."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_kernel_small():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import all_kernels_smoke
    all_kernels_smoke.main()


@pytest.mark.parametrize("kind", ["wl", "wo", "ol", "oo", "om"])
def test_device_outputs_stay_inside_their_buffers(kind):
    import torch
    from trik_media_sensors_dsp_b200 import open_sensor, sensors, synth, xdm
    w, h, ow, oh, n = 320, 240, 160, 120, 5
    k = xdm.KIND_OF[kind]
    layout = sensors.layout_of(k)
    fbytes = synth.frame_bytes(w, h, layout)
    rec = C.sizeof(xdm.OUT_ARGS_ALG[k])
    pbytes = ow * oh * 2
    guard = 4096
    dev = torch.device("cuda", 0)
    frames = torch.from_numpy(synth.make_batch("scene", range(n), w, h, layout)).to(dev)
    outbuf = torch.full((guard + n * rec + guard,), 0xCD, dtype=torch.uint8, device=dev)
    prevbuf = torch.full((guard + n * pbytes + guard,), 0xCD, dtype=torch.uint8, device=dev)
    codec = open_sensor(kind, w, h, out_w=ow, out_h=oh)
    ia = {"oo": xdm.ObjInArgsAlg(1, 200, 45, 55, 40, 50, 45, 0), "om": xdm.MxnInArgsAlg(3, 3)}.get(kind, xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 45, 0))
    ret, _ = codec.process_batch(frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                 out_device_ptr=outbuf.data_ptr() + guard,
                                 previews_device_ptr=prevbuf.data_ptr() + guard, preview_stride=pbytes)
    assert ret == 0, sensors.last_error()
    torch.cuda.synchronize()
    for buf, body in ((outbuf, n * rec), (prevbuf, n * pbytes)):
        b = buf.cpu().numpy()
        assert (b[:guard] == 0xCD).all() and (b[guard + body:] == 0xCD).all(), kind
    assert (prevbuf.cpu().numpy()[guard:guard + n * pbytes] != 0xCD).any()
    # device results equal the host path (SETPARAMS first: a fresh algorithm object, same carried state as above)
    assert codec.set_params(w, h, out_w=ow, out_h=oh) == 0
    ret, outs = codec.process_batch(frames.cpu().numpy(), ia)
    nb = {"om": 36, "oo": 24}.get(kind, 3)
    dev_recs = outbuf.cpu().numpy()[guard:guard + n * rec].reshape(n, rec)
    for i in range(n):
        assert bytes(dev_recs[i][:nb]) == bytes(memoryview(outs[i]))[:nb]
    codec.close()
