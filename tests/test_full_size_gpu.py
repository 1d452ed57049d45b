"""BASELINE.json's full batch sizes, checked through properties that do not need the oracle on every frame:

  * a batch is n sequential process() calls, so identical frames under identical carried state give identical
    records wherever they sit in the batch (no cross-frame contamination at 4096 frames per launch);
  * the distinct frames of the batch equal the oracle (run on those few frames only);
  * cutting the batch into uneven pieces (carried state crossing the cuts) or leaving the frames in device memory
    changes nothing;
  * a checksum of the records equals the checksum predicted from the distinct frames.

Config 2: webcam line sensor, 4096 x 320x240 YUYV.  Config 3: mxn grid sensor 3x3 / 5x5, 1024 x 640x480.  The other
sensors at 4096 x 320x240 as well."""
import ctypes as C
import zlib

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu
UNIQUE = 64

CASES = [
    # kind, w, h, n, family, in args (product), in args (oracle), bytes of the record that are defined
    ("wl", 320, 240, 4096, "scene", (0, 359, 0, 100, 0, 40, 0), 3),
    ("ol", 320, 240, 4096, "scene", (0, 359, 0, 100, 0, 40, 0), 3),
    ("wo", 320, 240, 4096, "scene", (300, 40, 20, 100, 30, 100, 0), 3),
    ("oo", 320, 240, 4096, "blobs", (1, 0, 20, 80, 20, 50, 30, 0), 24),
    ("om", 640, 480, 1024, "grid3", (3, 3), 36),
    ("om", 640, 480, 1024, "grid5", (5, 5), 100),
]


def _args(kind, a):
    if kind == "oo":
        return xdm.ObjInArgsAlg(*a), oracle.ObjInArgs(*a)
    if kind == "om":
        return xdm.MxnInArgsAlg(*a), oracle.MxnInArgs(*a)
    return xdm.RangeInArgsAlg(*a), oracle.RangeInArgs(*a)


def _records(outs, nbytes):
    rec = C.sizeof(outs._type_)
    raw = np.frombuffer(memoryview(outs), dtype=np.uint8).reshape(len(outs), rec)
    return np.ascontiguousarray(raw[:, :nbytes])


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%dx%d-%d-%s" % (c[0], c[1], c[2], c[3], c[4]))
def test_full_batch_properties(case):
    kind, w, h, n, fam, a, nbytes = case
    layout = sensors.layout_of(xdm.KIND_OF[kind])
    if fam.startswith("grid"):
        g = int(fam[4:])
        uniq = synth.make_batch("grid", range(UNIQUE), w, h, layout, m=g, n=g)
    else:
        uniq = synth.make_batch(fam, range(UNIQUE), w, h, layout)
    frames = np.empty((n, uniq.shape[1]), np.uint8)
    for i in range(0, n, UNIQUE):
        frames[i:i + UNIQUE] = uniq[:min(UNIQUE, n - i)]
    ia, oia = _args(kind, a)

    codec = open_sensor(kind, w, h)
    ret, outs = codec.process_batch(frames, ia)
    assert ret == 0, sensors.last_error()
    got = _records(outs, nbytes)

    # 1. the distinct frames against the oracle.  The line sensor's cross band lags one frame (its first call on a fresh
    #    object is indeterminate in the reference), so the oracle walks the first UNIQUE + 1 frames and frame 0 is skipped.
    orc = oracle.OracleSensor(kind, w, h)
    want = []
    for i in range(UNIQUE + 1):
        ok, exp = orc.process(frames[i], oia)
        assert ok == 1
        undefined = kind == "oo" and (orc.last_flags() & 2)          # fewer than 8 labels: the reference reads past its vector
        want.append(None if undefined else bytes(memoryview(exp))[:nbytes])
    first = 1 if kind == "ol" else 0
    for i in range(first, UNIQUE + 1):
        if want[i] is not None:
            assert got[i].tobytes() == want[i], (kind, i)

    # 2. position independence: every later copy of a frame gives the record of its first copy (for the line sensor
    #    from the second period on, when the carried band has settled)
    base = got[UNIQUE:2 * UNIQUE] if kind == "ol" else got[:UNIQUE]
    start = UNIQUE if kind == "ol" else 0
    for i in range(start, n, UNIQUE):
        blk = got[i:i + UNIQUE]
        assert np.array_equal(blk, base[:len(blk)]), (kind, "period at", i)

    # 3. checksum of checksums predicted from one period
    periods, rest = divmod(n - start, UNIQUE)
    crc_period = zlib.crc32(base.tobytes())
    crcs = [zlib.crc32(got[start + p * UNIQUE: start + (p + 1) * UNIQUE].tobytes()) for p in range(periods)]
    assert crcs == [crc_period] * periods
    assert int(got[start:, :3].astype(np.int64).sum()) == periods * int(base[:, :3].astype(np.int64).sum()) \
        + int(base[:rest, :3].astype(np.int64).sum())

    # 4. uneven pieces through a fresh handle state (carried state crosses the cuts) == one batch
    codec.set_params(w, h)
    pieces = []
    cuts = [0, 1, n // 4 - 24, n // 4 - 23, n - 37, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ret, o = codec.process_batch(frames[lo:hi], ia)
        assert ret == 0, sensors.last_error()
        pieces.append(_records(o, nbytes))
    assert np.array_equal(np.concatenate(pieces), got), kind

    # 5. frames resident in device memory, results to device memory == host path
    import torch
    codec.set_params(w, h)
    d_frames = torch.from_numpy(frames).cuda()
    rec = C.sizeof(xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]])
    d_out = torch.zeros((n, rec), dtype=torch.uint8, device="cuda")
    ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=frames.shape[1], num_frames=n,
                                 out_device_ptr=d_out.data_ptr())
    assert ret == 0, sensors.last_error()
    codec.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy()[:, :nbytes], got), kind
    codec.close()
