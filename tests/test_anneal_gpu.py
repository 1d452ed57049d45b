"""The annealing tail of the WL / OL / OO auto-calibration on the device (TRIKB200_BATCH_DEVICE_TAIL) against the
host tail (the default, pinned to the reference by the other tests): same histograms, same seeds.  The two differ
only in pow() (device <= 2 ulp, libm <= 1 ulp); a last-bit difference matters only if it moves a value across an
integer, so the outputs are expected to be identical on every frame tried here -- the assert documents that."""
import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, synth, xdm

pytestmark = pytest.mark.gpu

N = 96


def detect_fields(o):
    return (o.detectHue, o.detectHueTolerance, o.detectSat, o.detectSatTolerance, o.detectVal, o.detectValTolerance)


@pytest.mark.parametrize("kind", ["wl", "ol", "oo"])
def test_device_tail_equals_host_tail(kind):
    w, h = 320, 240
    layout = "yuyv" if kind == "wl" else "yuv422p"
    fams = [("scene", s) for s in range(N // 3)] + [("noise", s) for s in range(N // 3)] + [("blobs", s) for s in range(N // 3)]
    frames = np.stack([synth.make_frame(f, s, w, h, layout) for f, s in fams])
    seeds = [1000 + 7 * i for i in range(len(fams))]
    ia = xdm.ObjInArgsAlg(1, 0, 40, 60, 40, 60, 40, 1) if kind == "oo" else xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 1)
    codec = open_sensor(kind, w, h)
    assert codec.set_params(w, h) == 0
    ret, host_out = codec.process_batch(frames, ia, seeds=seeds)
    assert ret == 0, lib().trikb200_lastError()
    assert codec.set_params(w, h) == 0
    ret, dev_out = codec.process_batch(frames, ia, seeds=seeds, flags=xdm.BATCH_DEVICE_TAIL)
    assert ret == 0, lib().trikb200_lastError()
    differing = [i for i in range(len(fams)) if bytes(memoryview(host_out[i])) != bytes(memoryview(dev_out[i]))]
    assert differing == [], (kind, [(fams[i], detect_fields(host_out[i]), detect_fields(dev_out[i])) for i in differing[:5]])
    # and the host tail is the oracle's (spot check, so that "equal" above means "equal to the reference")
    orc = oracle.OracleSensor(kind, w, h)
    for i in range(0, len(fams), 16):
        ok, exp = orc.process(frames[i], oracle.IN_ARGS[kind](*[getattr(ia, f[0]) for f in ia._fields_]), seed=seeds[i])
        if orc.last_flags():
            continue
        assert detect_fields(exp) == detect_fields(host_out[i]), (kind, fams[i])
    codec.close()


def test_device_tail_allows_async_device_results():
    """What the flag is for: calibration without any host step -- frames and results on the device, enqueue only."""
    import torch
    w, h = 320, 240
    frames = np.stack([synth.make_frame("scene", s, w, h, "yuyv") for s in range(8)])
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 100, 1)
    codec = open_sensor("wl", w, h)
    ret, want = codec.process_batch(frames, ia, seeds=[5] * 8)
    assert ret == 0
    assert codec.set_params(w, h) == 0
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((8, 16), dtype=torch.uint8, device="cuda")
    ret, _ = codec.process_batch(d_frames.data_ptr(), ia, seeds=[5] * 8, frames_device=True, frame_stride=frames.shape[1],
                                 num_frames=8, out_device_ptr=d_out.data_ptr(), flags=xdm.BATCH_ASYNC)
    assert ret != 0                                    # the host tail cannot run behind an enqueue-only call
    ret, _ = codec.process_batch(d_frames.data_ptr(), ia, seeds=[5] * 8, frames_device=True, frame_stride=frames.shape[1],
                                 num_frames=8, out_device_ptr=d_out.data_ptr(), flags=xdm.BATCH_ASYNC | xdm.BATCH_DEVICE_TAIL)
    assert ret == 0, lib().trikb200_lastError()
    assert codec.synchronize() == 0
    got = d_out.cpu().numpy()
    for i in range(8):
        rec = bytes(memoryview(want[i]))
        assert got[i].tobytes()[:3] == rec[:3] and got[i].tobytes()[4:16] == rec[4:16], i
    codec.close()
