"""TEST INFRASTRUCTURE (oracle/): ctypes access to the host build of the reference
(oracle/_ref/libtrikref_<kind>.so, built by oracle/build_ref.sh from /root/reference) and to
the C restatement (oracle/liboracle.so, built from oracle/trik_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under the product package does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

KINDS = ("wo", "wl", "oo", "ol", "om")
# TRIK_VIDTRANSCODE_CV_VideoFormat (ov7670/object_sensor/trik_vidtranscode_cv.h:20-32)
FORMAT_YUV422 = 0x100 + 4
FORMAT_YUV422P = 0x100 + 5
FORMAT_OF = {"wo": FORMAT_YUV422, "wl": FORMAT_YUV422, "oo": FORMAT_YUV422P,
             "ol": FORMAT_YUV422P, "om": FORMAT_YUV422P, "oe": FORMAT_YUV422P}
# "oe" = ov7670/edge_line_sensor: only as a RefSensor (the reference's own sensor code built against
# the open IMGLIB restatement oracle/imglib_open.c) and through edge_line() of the C restatement -- not an OracleSensor kind
REF_KINDS = KINDS + ("oe",)


class RangeInArgs(C.Structure):
    """InArgsAlg of WO / WL / OL (webcam/line_sensor/trik_vidtranscode_cv.h:48-56)."""
    _fields_ = [("detectHueFrom", C.c_uint16), ("detectHueTo", C.c_uint16),
                ("detectSatFrom", C.c_uint8), ("detectSatTo", C.c_uint8),
                ("detectValFrom", C.c_uint8), ("detectValTo", C.c_uint8),
                ("autoDetectHsv", C.c_uint8)]


class TargetOutArgs(C.Structure):
    """OutArgsAlg of WO / WL / OL (webcam/line_sensor/trik_vidtranscode_cv.h:64-74)."""
    _fields_ = [("targetX", C.c_int8), ("targetY", C.c_int8), ("targetSize", C.c_uint8),
                ("detectHue", C.c_uint16), ("detectHueTolerance", C.c_uint16),
                ("detectSat", C.c_uint16), ("detectSatTolerance", C.c_uint16),
                ("detectVal", C.c_uint16), ("detectValTolerance", C.c_uint16)]


class ObjInArgs(C.Structure):
    """InArgsAlg of OO (ov7670/object_sensor/trik_vidtranscode_cv.h:51-60)."""
    _fields_ = [("setHsvRange", C.c_uint8),
                ("detectHue", C.c_uint16), ("detectHueTol", C.c_uint16),
                ("detectSat", C.c_uint8), ("detectSatTol", C.c_uint8),
                ("detectVal", C.c_uint8), ("detectValTol", C.c_uint8),
                ("autoDetectHsv", C.c_uint8)]


class XdasTarget(C.Structure):
    _fields_ = [("x", C.c_int8), ("y", C.c_int8), ("size", C.c_uint8)]


class ObjOutArgs(C.Structure):
    """OutArgsAlg of OO (ov7670/object_sensor/trik_vidtranscode_cv.h:67-81)."""
    _fields_ = [("target", XdasTarget * 8),
                ("detectHue", C.c_uint16), ("detectHueTolerance", C.c_uint16),
                ("detectSat", C.c_uint16), ("detectSatTolerance", C.c_uint16),
                ("detectVal", C.c_uint16), ("detectValTolerance", C.c_uint16)]


class MxnInArgs(C.Structure):
    """InArgsAlg of OM (ov7670/mxn_sensor/trik_vidtranscode_cv.h:49-52)."""
    _fields_ = [("widthM", C.c_int32), ("heightN", C.c_int32)]


class MxnOutArgs(C.Structure):
    """OutArgsAlg of OM (ov7670/mxn_sensor/trik_vidtranscode_cv.h:60-62)."""
    _fields_ = [("outColor", C.c_int32 * 100)]


IN_ARGS = {"wo": RangeInArgs, "wl": RangeInArgs, "ol": RangeInArgs, "oo": ObjInArgs, "om": MxnInArgs, "oe": RangeInArgs}
OUT_ARGS = {"wo": TargetOutArgs, "wl": TargetOutArgs, "ol": TargetOutArgs, "oo": ObjOutArgs, "om": MxnOutArgs, "oe": TargetOutArgs}


def ref_available(kind="wo"):
    return os.path.exists(os.path.join(REF_DIR, "libtrikref_%s.so" % kind))


def aligned_bytes(n, align=64):
    """uint8 numpy array of n bytes whose data pointer is align-byte aligned
    (the reference wants 8-byte aligned rows, WO/.../cv_ball_detector_seqpass.hpp:263)."""
    raw = np.zeros(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n]


class RefSensor:
    """One live instance of the host-built reference codec of a given kind.

    The shared object keeps file-scope statics, so there is exactly one instance per kind per
    process; creating another RefSensor of the same kind re-creates the codec.
    """

    def __init__(self, kind, suffix=""):
        assert kind in REF_KINDS
        path = os.path.join(REF_DIR, "libtrikref_%s%s.so" % (kind, suffix))
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with: make -C oracle ref)")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.lib.trikref_create.argtypes = [C.c_int] * 9 + [C.POINTER(C.c_int)]
        self.lib.trikref_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.lib.trikref_set_time.argtypes = [C.c_longlong]
        self.lib.trikref_probe_yuv2rgb.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        self.lib.trikref_probe_rgb2hsv.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        self.lib.trikref_fxns.restype = C.c_void_p
        self.lib.trikref_handle.restype = C.c_void_p
        assert self.lib.trikref_sizeof_inargs_alg() == C.sizeof(IN_ARGS[kind]), "InArgsAlg layout"
        assert self.lib.trikref_sizeof_outargs_alg() == C.sizeof(OUT_ARGS[kind]), "OutArgsAlg layout"
        self.w = self.h = self.line = 0
        self.preview = None

    def setup(self, w, h, line_length=None, out_w=None, out_h=None, max_w=640, max_h=480):
        """alloc -> initObj -> control(XDM_SETPARAMS).  Returns (code, detail): code 0 = ok."""
        if line_length is None:
            line_length = 2 * w if FORMAT_OF[self.kind] == FORMAT_YUV422 else w
        out_w = w if out_w is None else out_w
        out_h = h if out_h is None else out_h
        detail = C.c_int(0)
        code = self.lib.trikref_create(FORMAT_OF[self.kind], max(max_w, w), max(max_h, h), w, h, line_length,
                                       out_w, out_h, out_w * 2, C.byref(detail))
        self.w, self.h, self.line = w, h, line_length
        self.preview = aligned_bytes(max(out_w * out_h * 2, 64))
        return code, detail.value

    def process(self, frame, in_args, out_args=None, seed=0, num_bytes=None):
        """One process() call.  Returns (ret, out_args, extendedError)."""
        assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"]
        if out_args is None:
            out_args = OUT_ARGS[self.kind]()
        self.lib.trikref_set_time(int(seed))
        ext = C.c_int(0)
        bits = C.c_int(0)
        n = frame.nbytes if num_bytes is None else num_bytes
        ret = self.lib.trikref_process(frame.ctypes.data, n, frame.nbytes, C.byref(in_args), C.byref(out_args),
                                       self.preview.ctypes.data, self.preview.nbytes, C.byref(ext), C.byref(bits))
        return ret, out_args, ext.value

    def close(self):
        self.lib.trikref_destroy()


# ---------------------------------------------------------------------------------------------
# the C restatement (oracle/trik_oracle.c)
# ---------------------------------------------------------------------------------------------
KIND_ID = {"wo": 0, "wl": 1, "oo": 2, "ol": 3, "om": 4}
PORT_PATH = os.path.join(HERE, "liboracle.so")
_port = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            raise FileNotFoundError(PORT_PATH + " (build with: make -C oracle port)")
        lib = C.CDLL(PORT_PATH)
        lib.trik_oracle_create.restype = C.c_void_p
        lib.trik_oracle_create.argtypes = [C.c_int] * 4
        lib.trik_oracle_destroy.argtypes = [C.c_void_p]
        lib.trik_oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
        lib.trik_oracle_last_hsv.restype = C.c_void_p
        lib.trik_oracle_last_hsv.argtypes = [C.c_void_p]
        lib.trik_oracle_last_flags.argtypes = [C.c_void_p]
        for f in ("trik_oracle_yuv_to_rgb888", "trik_oracle_rgb888_to_hsv", "trik_oracle_hsv_to_rgb_mxn"):
            getattr(lib, f).restype = C.c_uint32
        lib.trik_oracle_yuv_to_rgb888.argtypes = [C.c_uint32] * 3
        lib.trik_oracle_rgb888_to_hsv.argtypes = [C.c_uint32]
        lib.trik_oracle_hsv_to_rgb_mxn.argtypes = [C.c_int] * 3
        lib.trik_oracle_detect.argtypes = [C.c_uint32] * 4
        lib.trik_oracle_yuv_to_rgb888_range.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        lib.trik_oracle_rgb888_to_hsv_range.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        lib.trik_oracle_srand.argtypes = [C.c_void_p, C.c_uint]
        lib.trik_oracle_rand.argtypes = [C.c_void_p]
        lib.trik_oracle_edge_line.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _port = lib
    return _port


class OracleSensor:
    """CVAlgorithm object of the C restatement: setup() at construction, run() per frame."""

    def __init__(self, kind, w, h, line_length=None):
        self.kind = kind
        self.lib = port_lib()
        if line_length is None:
            line_length = 2 * w if FORMAT_OF[kind] == FORMAT_YUV422 else w
        self.w, self.h, self.line = w, h, line_length
        self.ptr = self.lib.trik_oracle_create(KIND_ID[kind], w, h, line_length)
        if not self.ptr:
            raise ValueError("setup() rejects %dx%d" % (w, h))

    def process(self, frame, in_args, out_args=None, seed=0, num_bytes=None):
        """Returns (ok, out_args)."""
        assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"]
        if out_args is None:
            out_args = OUT_ARGS[self.kind]()
        n = frame.nbytes if num_bytes is None else num_bytes
        ok = self.lib.trik_oracle_run(self.ptr, frame.ctypes.data, n, C.byref(in_args), C.byref(out_args), int(seed))
        return ok, out_args

    def last_flags(self):
        return self.lib.trik_oracle_last_flags(self.ptr)

    def last_hsv(self):
        p = self.lib.trik_oracle_last_hsv(self.ptr)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.h, self.w)).copy()

    def close(self):
        if self.ptr:
            self.lib.trik_oracle_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def struct_bytes(s):
    return bytes(memoryview(s))


def edge_line(frame, w, h, line_length=None):
    """ov7670/edge_line_sensor restated (trik_oracle_edge_line): returns a TargetOutArgs with targetX / targetY / targetSize."""
    assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"] and frame.nbytes >= (line_length or w) * h
    out = TargetOutArgs()
    port_lib().trik_oracle_edge_line(frame.ctypes.data, w, h, line_length or w, C.byref(out))
    return out
