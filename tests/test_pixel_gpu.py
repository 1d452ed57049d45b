"""Exhaustive device pixel functions: all 2^24 (Y,U,V) triples and all 2^24 RGB triples through the
DEVICE code (both the YUYV and the YUV422P lane paths) against the oracle's closed forms."""
import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import sensors

pytestmark = pytest.mark.gpu
N24 = 1 << 24


def test_yuv_to_rgb_all_inputs():
    lib = sensors.lib()
    got = np.empty(N24, np.uint32)
    assert lib.trikb200_probePixels(0, 0, N24, got.ctypes.data) == 0, sensors.last_error()
    want = np.empty(N24, np.uint32)
    oracle.port_lib().trik_oracle_yuv_to_rgb888_range(0, N24, want.ctypes.data)
    assert not (got & 0x80000000).any(), "YUYV and YUV422P lane paths disagree"
    assert np.array_equal(got, want)


def test_rgb_to_hsv_all_inputs():
    lib = sensors.lib()
    got = np.empty(N24, np.uint32)
    assert lib.trikb200_probePixels(1, 0, N24, got.ctypes.data) == 0, sensors.last_error()
    want = np.empty(N24, np.uint32)
    oracle.port_lib().trik_oracle_rgb888_to_hsv_range(0, N24, want.ctypes.data)
    assert np.array_equal(got, want)


def test_packed_pair_hsv_all_inputs():
    """All 2^24 (Y,U,V) triples through the packed two-pixel HSV path the sensors use (64-scaled channels,
    lane-packed sector select), in both lane positions and both layouts, against hsv(rgb(yuv)) of the oracle."""
    lib = sensors.lib()
    got = np.empty(N24, np.uint32)
    assert lib.trikb200_probePixels(2, 0, N24, got.ctypes.data) == 0, sensors.last_error()
    olib = oracle.port_lib()
    rgb = np.empty(N24, np.uint32)
    olib.trik_oracle_yuv_to_rgb888_range(0, N24, rgb.ctypes.data)
    hsv_of_rgb = np.empty(N24, np.uint32)
    olib.trik_oracle_rgb888_to_hsv_range(0, N24, hsv_of_rgb.ctypes.data)
    want = hsv_of_rgb[rgb]
    assert not (got & 0x80000000).any(), "lane / layout paths disagree"
    assert np.array_equal(got, want)
