"""ov7670/edge_line_sensor as a batch operation (csrc/trik_kernels_edge.cu).

PARITY UNPINNED for the sensor's two IMGLIB kernels (closed TI library, absent from the reference tree): they are restated
from TI's published natural-C models in oracle/imglib_open.c.  Checkers: (1) the reference's OWN edge_line_sensor code built
against that restatement (oracle/_ref/libtrikref_oe.so; 320x240 at most, its work buffers are that size) and (2) the C
restatement of the whole sensor (trik_oracle_edge_line), which is size generic -- the two are compared with each other in the
CPU suite (tests/test_oracle_vs_ref.py)."""
import ctypes as C

import numpy as np
import pytest

from conftest import requires_ref
from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, sensors, synth, xdm

pytestmark = pytest.mark.gpu

FAMS = [("scene", s) for s in range(6)] + [("noise", 0), ("noise", 1), ("blobs", 1), ("blobs", 2), ("camera", 3)] \
       + [(e, 0) for e in synth.EDGE_CASES]


def rec(o):
    return (o.targetX, o.targetY, o.targetSize)


@pytest.mark.parametrize("size", [(320, 240), (160, 120), (32, 4), (64, 8), (640, 480), (384, 200), (1024, 64)])
def test_edge_line_matches_the_restatement(size):
    w, h = size
    frames = np.stack([synth.make_frame(f, s, w, h, "yuv422p") for f, s in FAMS])
    want = [rec(oracle.edge_line(frames[i], w, h)) for i in range(frames.shape[0])]
    try:
        for variant in (0, 1):                                 # packed four-pixel kernel, one-thread-per-column kernel
            lib().trikb200_setEdgeLineVariant(variant)
            ret, outs = sensors.edge_line_batch(frames, w, h)
            assert ret == 0, sensors.last_error()
            for i in range(frames.shape[0]):
                assert rec(outs[i]) == want[i], (size, variant, FAMS[i])
    finally:
        lib().trikb200_setEdgeLineVariant(0)
    if w >= 160:
        assert len(set(want)) > 4                              # the vectors do tell frames apart


@requires_ref
@pytest.mark.parametrize("size", [(320, 240), (160, 120), (320, 100)])
def test_edge_line_matches_the_reference_sensor_code(size):
    w, h = size
    rs = oracle.RefSensor("oe")
    # the reference's work buffer is a file-scope static that is never cleared: rows the Sobel does not write keep what an
    # earlier geometry left there.  An all-zero frame at the largest geometry wipes it, as in a fresh process.
    assert rs.setup(320, 240)[0] == 0
    zero = oracle.aligned_bytes(3 * 320 * 240)
    zero[:] = 0
    assert rs.process(zero, oracle.RangeInArgs(0, 359, 0, 100, 0, 100, 0), num_bytes=2 * 320 * 240)[0] == 0
    assert rs.setup(w, h)[0] == 0
    frames = np.stack([synth.make_frame(f, s, w, h, "yuv422p") for f, s in FAMS])
    ret, outs = sensors.edge_line_batch(frames, w, h)
    assert ret == 0, sensors.last_error()
    buf = oracle.aligned_bytes(3 * w * h)                      # the reference reads 2*W*H bytes of "chroma" (:160-168)
    for i in range(frames.shape[0]):
        buf[:] = 0
        buf[:frames.shape[1]] = frames[i]
        rret, rout, _ = rs.process(buf, oracle.RangeInArgs(0, 359, 0, 100, 0, 100, 0), num_bytes=2 * w * h)
        assert rret == 0 and rec(outs[i]) == rec(rout), (size, FAMS[i])


def test_row_sums_wrap_like_the_reference_uint16():
    """From 384 columns on a row's column sum can pass 65535 and the reference's uint16_t wraps: an all-edges frame."""
    w, h = 640, 16
    y = np.repeat(((np.arange(h) % 4 < 2) * 255).astype(np.uint8)[:, None], w, axis=1)          # stripes two rows thick: every pixel an edge
    frame = synth.pack(y, np.full((h, w // 2), 128, np.uint8), np.full((h, w // 2), 128, np.uint8), "yuv422p")[None, :]
    want = oracle.edge_line(frame[0], w, h)
    try:
        for variant in (0, 1):
            lib().trikb200_setEdgeLineVariant(variant)
            ret, outs = sensors.edge_line_batch(frame, w, h)
            assert ret == 0 and rec(outs[0]) == rec(want), variant
    finally:
        lib().trikb200_setEdgeLineVariant(0)
    assert rec(want)[0] != 0                                   # with exact sums the centroid would be the middle (0)


def test_strides_device_memory_and_rejections():
    import torch
    w, h, n = 320, 240, 5
    line, stride = w + 32, (w + 32) * h + 64
    rng = np.random.default_rng(5)
    host = rng.integers(0, 256, (n, stride), dtype=np.uint8)
    d_frames = torch.from_numpy(host).cuda()
    d_out = torch.full((n, 24), 0xEE, dtype=torch.uint8, device="cuda")
    d = xdm.EdgeLineBatch()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.lineLength = n, w, h, line
    d.framesMem, d.outArgsMem, d.outArgsStride = xdm.MEM_DEVICE, xdm.MEM_DEVICE, 24
    d.frames, d.frameStride, d.outArgsAlg = d_frames.data_ptr(), stride, d_out.data_ptr()
    assert lib().trikb200_edgeLineBatch(C.byref(d)) == 0, sensors.last_error()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for i in range(n):
        want = oracle.edge_line(np.ascontiguousarray(host[i]), w, h, line)
        assert (int(got[i, 0].view(np.int8)), int(got[i, 1].view(np.int8)), int(got[i, 2])) == rec(want)
        assert (got[i, 16:] == 0xEE).all()
    d.width = 100
    assert lib().trikb200_edgeLineBatch(C.byref(d)) == xdm.XDM_EFAIL
    d.width, d.size = w, 8
    assert lib().trikb200_edgeLineBatch(C.byref(d)) == xdm.XDM_EFAIL
