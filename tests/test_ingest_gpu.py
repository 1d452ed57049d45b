"""RGB565 -> YUV422P ingest front end (csrc/trik_kernels_ingest.cu).

PARITY UNPINNED: the reference has no RGB565 input path, so there is no reference arithmetic to compare with.  What is
checked: the kernel against a numpy restatement of the conversion include/trik_b200.h defines (every RGB565 word), the
layout the ov7670 sensors expect (a converted frame run through a sensor equals the oracle on the same bytes), and a
round trip through the sensors' own RGB565X preview (the reference's YUV -> RGB, then this RGB -> YUV, lands close to
the frame it started from)."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import lib, open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu


def restate(words, w, h, bgr):
    """numpy restatement of the documented conversion: (n, h, w) uint16 -> (n, 2*w*h) uint8 YUV422P."""
    px = words.astype(np.int32)
    hi, g6, lo = (px >> 11) & 31, (px >> 5) & 63, px & 31
    r5, b5 = (lo, hi) if bgr else (hi, lo)
    R, G, B = (r5 << 3) | (r5 >> 2), (g6 << 2) | (g6 >> 4), (b5 << 3) | (b5 >> 2)
    y = ((66 * R + 129 * G + 25 * B + 128) >> 8) + 16
    Rs, Gs, Bs = R[:, :, 0::2] + R[:, :, 1::2], G[:, :, 0::2] + G[:, :, 1::2], B[:, :, 0::2] + B[:, :, 1::2]
    u = ((-38 * Rs - 74 * Gs + 112 * Bs + 256) >> 9) + 128
    v = ((112 * Rs - 94 * Gs - 18 * Bs + 256) >> 9) + 128
    n = words.shape[0]
    out = np.empty((n, 2 * w * h), np.uint8)
    out[:, :w * h] = y.reshape(n, -1)
    chroma = np.empty((n, h, w), np.uint8)
    chroma[:, :, 0::2] = v
    chroma[:, :, 1::2] = u
    out[:, w * h:] = chroma.reshape(n, -1)
    return out


@pytest.mark.parametrize("fmt", [xdm.PIXEL_RGB565, xdm.PIXEL_RGB565X])
def test_every_rgb565_word(fmt):
    # all 65536 words as left pixels against 4 right-pixel patterns: a 512 x 512 image of pairs
    w, h = 512, 256
    left = np.arange(65536, dtype=np.uint16).reshape(h, w // 2)
    frames = []
    l32 = left.astype(np.uint32)
    for right in (l32, (l32 * 40503) & 0xFFFF, np.full_like(l32, 0x1234), l32 ^ 0xFFFF):
        img = np.empty((h, w), np.uint16)
        img[:, 0::2], img[:, 1::2] = left, right.astype(np.uint16)
        frames.append(img)
    words = np.stack(frames)
    ret, got = sensors.ingest_rgb565(words.view(np.uint8).reshape(len(frames), -1), w, h, fmt)
    assert ret == 0, sensors.last_error()
    assert np.array_equal(got, restate(words, w, h, fmt == xdm.PIXEL_RGB565X))
    assert got[:, :w * h].min() >= 16 and got[:, :w * h].max() <= 235
    assert got[:, w * h:].min() >= 16 and got[:, w * h:].max() <= 240


def test_strides_and_device_memory():
    import torch
    w, h, n = 64, 8, 5
    rng = np.random.default_rng(3)
    src_line, dst_line = 2 * w + 32, w + 8
    src_stride, dst_stride = src_line * h + 48, 2 * dst_line * h + 24
    src = np.zeros((n, src_stride), np.uint8)
    words = rng.integers(0, 65536, (n, h, w), dtype=np.uint16)
    for i in range(n):
        rows = src[i, :src_line * h].reshape(h, src_line)
        rows[:, :2 * w] = words[i].view(np.uint8).reshape(h, 2 * w)
    d_src = torch.from_numpy(src).cuda()
    d_dst = torch.full((n, dst_stride), 0xEE, dtype=torch.uint8, device="cuda")
    d = xdm.Ingest()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.pixelFormat = n, w, h, xdm.PIXEL_RGB565
    d.srcMem, d.dstMem = xdm.MEM_DEVICE, xdm.MEM_DEVICE
    d.srcLineLength, d.dstLineLength = src_line, dst_line
    d.src, d.srcStride, d.dst, d.dstStride = d_src.data_ptr(), src_stride, d_dst.data_ptr(), dst_stride
    assert lib().trikb200_ingestRgb565(C.byref(d)) == 0, sensors.last_error()
    torch.cuda.synchronize()
    got = d_dst.cpu().numpy()
    want = restate(words, w, h, False)
    for i in range(n):
        planes = got[i, :2 * dst_line * h].reshape(2 * h, dst_line)
        assert np.array_equal(planes[:h, :w].reshape(-1), want[i, :w * h])
        assert np.array_equal(planes[h:, :w].reshape(-1), want[i, w * h:])
        assert (planes[:, w:] == 0xEE).all() and (got[i, 2 * dst_line * h:] == 0xEE).all()      # padding untouched


def test_rejects_bad_descriptors():
    d = xdm.Ingest()
    d.size = C.sizeof(d) - 4
    assert lib().trikb200_ingestRgb565(C.byref(d)) == xdm.XDM_EFAIL
    src = np.zeros((1, 2 * 12 * 4), np.uint8)
    ret, _ = sensors.ingest_rgb565(src, 12, 4)                  # width % 8 != 0
    assert ret == xdm.XDM_EFAIL


def test_converted_frames_feed_the_ov7670_sensors():
    w, h = 320, 240
    rng = np.random.default_rng(11)
    img = np.zeros((6, h, w), np.uint16)
    for i in range(6):
        img[i] = rng.integers(0, 65536, dtype=np.uint16)                       # background colour
        x0, y0 = int(rng.integers(0, w - 80)), int(rng.integers(0, h - 60))
        img[i, y0:y0 + 60, x0:x0 + 80] = 0xF800                                # a red block
        img[i, :, 150:170] = 0x0000                                            # a dark band
    ret, frames = sensors.ingest_rgb565(img.view(np.uint8).reshape(6, -1), w, h)
    assert ret == 0
    for kind, ia, oia in (("ol", xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 30, 0), oracle.RangeInArgs(0, 359, 0, 100, 0, 30, 0)),
                          ("om", xdm.MxnInArgsAlg(3, 3), oracle.MxnInArgs(3, 3)),
                          ("oo", xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 40, 0), oracle.ObjInArgs(1, 0, 20, 80, 20, 50, 40, 0))):
        codec = open_sensor(kind, w, h)
        orc = oracle.OracleSensor(kind, w, h)
        ret, outs = codec.process_batch(frames, ia)
        assert ret == 0, sensors.last_error()
        for i in range(frames.shape[0]):
            ok, exp = orc.process(frames[i], oia)
            if kind == "oo" and (orc.last_flags() & 2):
                continue                                   # fewer than 8 labels: undefined in the reference
            if kind == "ol" and i == 0:
                continue                                   # first call of a fresh line sensor: indeterminate band
            n = 9 * 4 if kind == "om" else (24 if kind == "oo" else 3)
            assert bytes(memoryview(outs[i]))[:n] == bytes(memoryview(exp))[:n], (kind, i)
        codec.close()


def test_round_trip_through_the_sensor_preview():
    """frame --(reference's YUV -> RGB565X preview, 1:1)--> words --(this front end)--> frame': the two matrices are
    inverses up to quantisation (5/6/5 bits: up to 7 of 255 per channel, which the chroma rows of the matrix amplify to about 10), so away from saturated colours the luma
    comes back within a few levels."""
    w, h = 320, 240
    y = np.tile(np.linspace(40, 200, w).astype(np.uint8), (h, 1))
    u = np.full((h, w // 2), 120, np.uint8)
    v = np.full((h, w // 2), 135, np.uint8)
    frame = synth.pack(y, u, v, "yuv422p")[None, :]
    codec = open_sensor("oo", w, h, out_w=w, out_h=h)
    previews = np.zeros((1, w * h * 2), np.uint8)
    # a range nothing in this frame falls into: the preview is the picture itself plus thin overlays
    ret, _ = codec.process_batch(frame, xdm.ObjInArgsAlg(1, 180, 1, 100, 1, 1, 1, 0), previews=previews)
    assert ret == 0, sensors.last_error()
    codec.close()
    ret, back = sensors.ingest_rgb565(previews, w, h, xdm.PIXEL_RGB565X)
    assert ret == 0
    yb = back[0, :w * h].reshape(h, w).astype(np.int32)
    ub = back[0, w * h:].reshape(h, w)[:, 1::2].astype(np.int32)
    vb = back[0, w * h:].reshape(h, w)[:, 0::2].astype(np.int32)
    # overlays (centre lines, marks) cover a small part of the picture: ask for 97 % of the pixels
    assert (np.abs(yb - y.astype(np.int32)) <= 6).mean() >= 0.97
    assert (np.abs(ub - 120) <= 12).mean() >= 0.95 and (np.abs(vb - 135) <= 12).mean() >= 0.95
