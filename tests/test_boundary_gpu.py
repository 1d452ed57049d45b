"""The drop-in boundary: every return code, error bit and bookkeeping field of process()/control()
(<sensor>/src/vidtranscode_cv_fxns.c:174-334), driven through the function
table exactly as Codec Engine would -- and, where oracle/_ref is present, the SAME driver run against
the host-built reference's table with the results compared field by field."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import Codec, open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu

_keep_alive = []          # the reference caches LUT pointers into its FIRST instance's fast RAM


def both(kind, params=None):
    """[(name, Codec)] for this library and, if built, the reference."""
    out = [("b200", Codec(kind, params))]
    if oracle.ref_available(kind):
        rlib = C.CDLL(oracle.os.path.join(oracle.REF_DIR, "libtrikref_%s.so" % kind))
        rlib.trikref_fxns.restype = C.c_void_p
        fx = C.cast(rlib.trikref_fxns(), C.POINTER(xdm.IVIDTRANSCODE_Fxns))
        c = Codec(kind, params, fxns=fx)
        _keep_alive.append((rlib, c))
        out.append(("ref", c))
    return out


def frame_for(kind, w=320, h=240):
    # OO: a frame with >= 8 labels -- with fewer the reference reads past its cluster vector (undefined, can SIGFPE)
    fam, seed = ("blobs", 1) if kind == "oo" else ("scene", 3)
    return synth.make_frame(fam, seed, w, h, sensors.layout_of(xdm.KIND_OF[kind]))


def default_in_alg(kind):
    if kind == "oo":
        return xdm.ObjInArgsAlg(1, 0, 40, 60, 40, 60, 40, 0)
    if kind == "om":
        return xdm.MxnInArgsAlg(3, 3)
    return xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)


def bufs(frame, preview, num_in=1):
    ib = xdm.XDM1_BufDesc()
    ib.numBufs = num_in
    ib.descs[0].buf = frame.ctypes.data if frame is not None else None
    ib.descs[0].bufSize = frame.nbytes if frame is not None else 0
    ptrs = (C.c_void_p * 1)(preview.ctypes.data)
    sizes = (C.c_int32 * 1)(preview.nbytes)
    ob = xdm.XDM_BufDesc(ptrs, 1, sizes)
    return ib, ob, (ptrs, sizes)


def call(c, frame, in_alg, num_bytes=None, in_size=None, out_size=None, num_in=1, preview=None, null_buf=False):
    preview = c.preview if preview is None else preview
    ib, ob, keep = bufs(None if null_buf else frame, preview, num_in)
    ia = c.InArgs()
    ia.base.size = C.sizeof(ia) if in_size is None else in_size
    ia.base.numBytes = frame.nbytes if num_bytes is None else num_bytes
    ia.base.inputID = 7
    ia.alg = in_alg
    oa = c.OutArgs()
    oa.base.size = C.sizeof(oa) if out_size is None else out_size
    ret = c.process_raw(ib, ob, ia, oa)
    return ret, oa, ib


@pytest.mark.parametrize("kind", xdm.KIND_NAMES)
def test_process_error_paths(kind):
    results = {}
    for name, c in both(kind, sensors.default_params(xdm.KIND_OF[kind])):
        assert c.init_result == 0
        assert c.set_params(320, 240) == 0
        fr = oracle.aligned_bytes(frame_for(kind).size)
        fr[:] = frame_for(kind)
        ia = default_in_alg(kind)
        r = []
        ret, oa, _ = call(c, fr, ia, in_size=4)                       # wrong InArgs size
        r.append(("in_size", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, out_size=8)                      # wrong OutArgs size
        r.append(("out_size", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, num_in=2)                        # numBufs != 1
        r.append(("num_bufs", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, null_buf=True, num_bytes=0)      # NULL input buffer
        r.append(("null_buf", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, num_bytes=-1)
        r.append(("neg_bytes", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, num_bytes=fr.nbytes + 1)         # numBytes > bufSize
        r.append(("too_many_bytes", ret, oa.base.extendedError))
        ret, oa, _ = call(c, fr, ia, num_bytes=100)                   # image does not fit numBytes
        r.append(("short_image", ret, oa.base.extendedError))
        small = np.zeros(64, dtype=np.uint8)
        ret, oa, _ = call(c, fr, ia, preview=small)                   # preview buffer too small
        r.append(("short_preview", ret, oa.base.extendedError))
        ret, oa, ib = call(c, fr, ia)                                 # the good call
        b = oa.base
        r.append(("ok", ret, b.extendedError, b.bitsConsumed, b.bitsGenerated[0], b.decodedPictureType,
                  b.decodedPictureStructure, b.encodedPictureType[0], b.encodedPictureStructure[0], b.decodedHeight,
                  b.decodedWidth, b.outputID[0], b.inputFrameSkipTranscodeFlag[0], b.outBufsInUseFlag,
                  b.encodedBuf[0].bufSize, b.encodedBuf[0].accessMask, ib.descs[0].accessMask,
                  b.encodedBuf[0].buf == c.preview.ctypes.data))
        results[name] = r
    mine = dict((x[0], x[1:]) for x in results["b200"])
    assert mine["in_size"] == (xdm.XDM_EUNSUPPORTED, 1 << xdm.XDM_UNSUPPORTEDPARAM)
    assert mine["out_size"] == (xdm.XDM_EUNSUPPORTED, 1 << xdm.XDM_UNSUPPORTEDPARAM)
    for k in ("num_bufs", "null_buf", "neg_bytes", "too_many_bytes"):
        assert mine[k] == (xdm.XDM_EFAIL, 1 << xdm.XDM_UNSUPPORTEDPARAM), k
    for k in ("short_image", "short_preview"):
        assert mine[k] == (xdm.XDM_EFAIL, 1 << xdm.XDM_CORRUPTEDDATA), k
    ok = mine["ok"]
    assert ok[0] == 0 and ok[1] == 0 and ok[2] == frame_for(kind).nbytes * 8 and ok[3] == 320 * 240 * 2 * 8
    assert ok[4] == -1 and ok[5] == -1 and ok[8] == 240 and ok[9] == 320 and ok[10] == 7 and ok[11] == 0 and ok[12] == 0
    assert ok[13] == 320 * 240 * 2 and ok[14] == 2 and ok[15] == 1 and ok[16]
    if "ref" in results:
        assert results["ref"] == results["b200"]


@pytest.mark.parametrize("kind", xdm.KIND_NAMES)
def test_control_commands(kind):
    results = {}
    for name, c in both(kind, sensors.default_params(xdm.KIND_OF[kind])):
        r = []
        for cmd in (xdm.XDM_GETSTATUS, xdm.XDM_GETBUFINFO):
            ret, st = c.control(cmd)
            r.append((cmd, ret, st.extendedError, st.bufInfo.minNumInBufs, st.bufInfo.minNumOutBufs,
                      st.bufInfo.minInBufSize[0], st.bufInfo.minOutBufSize[0], st.data.accessMask))
        r.append(("version",) + c.get_version())
        st = xdm.IVIDTRANSCODE_Status()
        st.size = C.sizeof(st)
        small = C.create_string_buffer(4)
        st.data.buf, st.data.bufSize = C.addressof(small), 4
        r.append(("version_small", c.control(xdm.XDM_GETVERSION, None, st)[0]))
        r.append(("flush", c.control(xdm.XDM_FLUSH)[0]))
        r.append(("unknown", c.control(99)[0]))
        r.append(("setparams_size", c.set_params(320, 240, dyn_size=12)))
        r.append(("setparams_ok", c.set_params(320, 240)))
        r.append(("width_not_32", c.set_params(328, 240)))
        r.append(("height_not_4", c.set_params(320, 242)))
        r.append(("too_wide", c.set_params(672, 480)))
        r.append(("too_high", c.set_params(640, 484)))
        r.append(("max", c.set_params(640, 480)))
        r.append(("reset", c.control(xdm.XDM_RESET)[0]))
        r.append(("setdefault", c.control(xdm.XDM_SETDEFAULT)[0]))
        results[name] = r
    mine = dict((x[0], x[1:]) for x in results["b200"])
    assert mine[xdm.XDM_GETSTATUS] == (0, 0, 1, 1, 0, 0, 2)
    assert mine["version"] == (0, "1.00.00.00") and mine["version_small"] == (xdm.XDM_EFAIL,)
    assert mine["flush"] == (0,) and mine["unknown"] == (xdm.XDM_EFAIL,)
    assert mine["setparams_size"] == (xdm.XDM_EUNSUPPORTED,) and mine["setparams_ok"] == (0,)
    for k in ("width_not_32", "height_not_4", "too_wide", "too_high"):
        assert mine[k] == (xdm.IALG_EFAIL,), k
    assert mine["max"] == (0,) and mine["reset"] == (0,) and mine["setdefault"] == (0,)
    if "ref" in results:
        assert results["ref"] == results["b200"]


@pytest.mark.parametrize("kind", ["wl", "ol"])
def test_line_sensor_init_needs_320_high_output(kind):
    """The line sensors default to a 240x320 preview, so initObj fails when maxHeightOutput < 320
    (SURVEY.md section 3.1)."""
    p = sensors.default_params(xdm.KIND_OF[kind], 640, 480, 640)
    p.base.maxHeightOutput[0] = 240
    for name, c in both(kind, p):
        assert c.init_result == xdm.IALG_EFAIL, name


def test_free_returns_the_alloc_table():
    c = open_sensor("wo", 320, 240)
    c.close()
    assert c.free_records == 2
    assert c.free_table[0].base == c.handle and c.free_table[0].size == C.sizeof(c._bufs[0])
    assert c.free_table[1].base == C.addressof(c._bufs[1]) and c.free_table[1].size == 0x1000
    assert c.free_table[0].space == xdm.IALG_EXTERNAL and c.free_table[1].space == xdm.IALG_DARAM0


def test_setparams_rebuilds_the_algorithm_object():
    """control(XDM_SETPARAMS) constructs a new algorithm object (vidtranscode_cv.cpp:58-59): the OL
    cross band lag starts again from 0,0."""
    w, h = 320, 240
    fr = synth.make_frame("halves", 0, w, h, "yuv422p")
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)
    c = open_sensor("ol", w, h)
    ys = [c.process(fr, ia)[1].alg.targetY for _ in range(3)]
    assert ys[0] != ys[1] and ys[1] == ys[2]
    assert c.set_params(w, h) == 0
    assert c.process(fr, ia)[1].alg.targetY == ys[0]
    c.close()


def test_preview_size_is_reported():
    """process() reports bufSize = outHeight * outLineLength and zero-fills what lies beyond the image
    (vidtranscode_cv_fxns.c:234,251)."""
    c = open_sensor("wl", 320, 240, out_w=160, out_h=120)
    big = np.full(160 * 120 * 2 + 512, 0xAB, dtype=np.uint8)
    ret, oa = c.process(synth.make_frame("scene", 0, 320, 240, "yuyv"), xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0), preview=big)
    assert ret == 0 and not big[160 * 120 * 2:].any() and big[:160 * 120 * 2].any()
    assert oa.base.encodedBuf[0].bufSize == 120 * 160 * 2 and oa.base.bitsGenerated[0] == 120 * 160 * 2 * 8
    c.close()


@pytest.mark.parametrize("kind", xdm.KIND_NAMES)
def test_line_length_is_validated(kind):
    """Documented deviation (DESIGN.md "Boundary"): the kernels read rows with 16-byte loads, so control(XDM_SETPARAMS)
    refuses an inputLineLength that is not a multiple of 16 or does not cover a row -- as IALG_EFAIL, never as a CUDA
    fault -- and the handle stays usable afterwards."""
    w, h = 320, 240
    c = open_sensor(kind, w, h)
    row = w if sensors.layout_of(xdm.KIND_OF[kind]) == "yuv422p" else 2 * w
    assert c.set_params(w, h, line_length=row + 8) == xdm.IALG_EFAIL      # misaligned rows
    assert c.set_params(w, h, line_length=0) == xdm.IALG_EFAIL            # explicit size, no stride
    assert c.set_params(w, h, line_length=row - 16) == xdm.IALG_EFAIL     # stride shorter than a row
    ret, _ = c.process(frame_for(kind), default_in_alg(kind))             # no valid geometry: EFAIL, no crash
    assert ret == xdm.XDM_EFAIL
    assert c.set_params(w, h, line_length=row + 16) == 0                  # padded rows are fine
    assert c.set_params(w, h) == 0
    ret, _ = c.process(frame_for(kind), default_in_alg(kind))
    assert ret == 0
    c.close()


def test_mxn_grid_arguments_are_validated_as_int32():
    """widthM / heightN are XDAS_Int32: 257 must not be taken for 1 (nor 256 for 0); non-positive values and
    products above 100 (outColor[100]) are refused."""
    w, h = 320, 240
    c = open_sensor("om", w, h)
    fr = frame_for("om")
    for m, n in ((257, 1), (1, 258), (256, 1), (0, 3), (3, 0), (-1, 3), (11, 10), (101, 1)):
        ret, _ = c.process(fr, xdm.MxnInArgsAlg(m, n))
        assert ret == xdm.XDM_EFAIL, (m, n)
    for m, n in ((1, 1), (10, 10), (100, 1), (1, 100)):
        ret, _ = c.process(fr, xdm.MxnInArgsAlg(m, n))
        assert ret == 0, (m, n)
    c.close()
