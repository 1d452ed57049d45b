"""trik-media-sensors-dsp_b200: B200-native drop-in for the per-frame pixel pipeline of
trikset/trik-media-sensors-dsp (webcam object/line sensors, ov7670 object/line/mxn sensors).

  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/trik_b200.h)
  build.py   in-tree nvcc build of libtrikb200.so
  xdm.py     ctypes mirror of the xDM / sensor argument structs
  sensors.py host-side mirror of the reference's codec interface (Codec, open_sensor)
  synth.py   seeded synthetic camera frames shared by tests and bench
"""
from . import xdm  # noqa: F401
from .sensors import Codec, TrikB200Error, open_sensor, lib, launch_count, last_error, process_mixed  # noqa: F401
