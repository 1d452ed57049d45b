"""Host-side mirror of the reference's codec interface over libtrikb200's C ABI.

`Codec` drives ONE sensor instance exactly the way Codec Engine drives the reference
(SURVEY.md section 3): alloc -> initObj -> control(XDM_SETPARAMS) -> process() per frame -> free,
through the IVIDTRANSCODE_Fxns table the library exports for that sensor kind.  `process_batch`
is the additive batch entry point (trikb200_processBatch).

There is no CPU path here: if the CUDA library is missing or fails to load, importing a Codec
raises; it never falls back to the oracle.
"""
import ctypes as C
import os

import numpy as np

from . import xdm

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtrikb200.so")
_lib = None


class TrikB200Error(RuntimeError):
    pass


def lib():
    """The loaded libtrikb200.so (built in-tree by build.py).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TrikB200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.trikb200_fxns.restype = C.POINTER(xdm.IVIDTRANSCODE_Fxns)
        l.trikb200_fxns.argtypes = [C.c_int32]
        l.trikb200_create.restype = C.c_void_p
        l.trikb200_create.argtypes = [C.c_int32] * 6
        l.trikb200_delete.argtypes = [C.c_void_p]
        l.trikb200_processBatch.argtypes = [C.c_void_p, C.POINTER(xdm.Batch)]
        l.trikb200_processBatch.restype = C.c_int32
        l.trikb200_processBatchMulti.argtypes = [C.c_void_p, C.c_int32, C.POINTER(xdm.Batch)]
        l.trikb200_processBatchMulti.restype = C.c_int32
        l.trikb200_processMixed.argtypes = [C.c_void_p, C.c_int32]
        l.trikb200_processMixed.restype = C.c_int32
        l.trikb200_synchronize.argtypes = [C.c_void_p]
        l.trikb200_setSeed.argtypes = [C.c_void_p, C.c_int64]
        l.trikb200_setDevice.argtypes = [C.c_int32]
        l.trikb200_launchCount.restype = C.c_int64
        l.trikb200_setSlabsPerFrame.argtypes = [C.c_int32]
        l.trikb200_setLoadStages.argtypes = [C.c_int32]
        l.trikb200_setBlockThreads.argtypes = [C.c_int32]
        l.trikb200_setOverlapLaunch.argtypes = [C.c_int32]
        l.trikb200_setLutMode.argtypes = [C.c_int32]
        l.trikb200_setMxnTableMode.argtypes = [C.c_int32]
        l.trikb200_setGatherMode.argtypes = [C.c_int32]
        l.trikb200_setLutSkew.argtypes = [C.c_int32]
        l.trikb200_setLutParts.argtypes = [C.c_int32]
        l.trikb200_setPreviewChunkMB.argtypes = [C.c_int32]
        l.trikb200_setPreviewSectorOverlay.argtypes = [C.c_int32]
        l.trikb200_setPreviewTable.argtypes = [C.c_int32]
        l.trikb200_setEdgeLineVariant.argtypes = [C.c_int32]
        l.trikb200_setMxnTableThreads.argtypes = [C.c_int32]
        l.trikb200_setZeroCopyBytes.argtypes = [C.c_int32]
        l.trikb200_setFramesPerCta.argtypes = [C.c_int32]
        l.trikb200_lastError.restype = C.c_char_p
        l.trikb200_probePixels.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p]
        l.trikb200_probeLut.argtypes = [C.c_void_p, C.c_void_p]
        l.trikb200_ingestRgb565.argtypes = [C.POINTER(xdm.Ingest)]
        l.trikb200_ingestRgb565.restype = C.c_int32
        l.trikb200_edgeLineBatch.argtypes = [C.POINTER(xdm.EdgeLineBatch)]
        l.trikb200_edgeLineBatch.restype = C.c_int32
        for f in ("trikb200_sizeofInArgsAlg", "trikb200_sizeofOutArgsAlg", "trikb200_sizeofInArgs",
                  "trikb200_sizeofOutArgs"):
            getattr(l, f).argtypes = [C.c_int32]
        _lib = l
    return _lib


def ingest_rgb565(src, width, height, pixel_format=xdm.PIXEL_RGB565, src_line_length=None, out=None):
    """Host-memory form of trikb200_ingestRgb565: src (n, src_line_length * height) uint8 holding packed RGB565 (or
    RGB565X) rows -> (n, 2 * width * height) uint8 YUV422P frames as the ov7670 sensors read them."""
    assert src.dtype == np.uint8 and src.ndim == 2 and src.flags["C_CONTIGUOUS"]
    line = 2 * width if src_line_length is None else src_line_length
    n = src.shape[0]
    if out is None:
        out = np.empty((n, 2 * width * height), np.uint8)
    d = xdm.Ingest()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.pixelFormat = n, width, height, pixel_format
    d.srcMem, d.dstMem = xdm.MEM_HOST, xdm.MEM_HOST
    d.srcLineLength, d.dstLineLength = line, width
    d.src, d.srcStride = src.ctypes.data, src.strides[0]
    d.dst, d.dstStride = out.ctypes.data, out.strides[0]
    ret = lib().trikb200_ingestRgb565(C.byref(d))
    return ret, out


def edge_line_batch(frames, width, height, line_length=None):
    """Host-memory form of trikb200_edgeLineBatch: frames (n, >= line_length * height) uint8 whose first plane is the luma
    -> (ret, TargetOutArgsAlg array)."""
    assert frames.dtype == np.uint8 and frames.ndim == 2 and frames.flags["C_CONTIGUOUS"]
    n = frames.shape[0]
    outs = (xdm.TargetOutArgsAlg * n)()
    d = xdm.EdgeLineBatch()
    d.size = C.sizeof(d)
    d.numFrames, d.width, d.height, d.lineLength = n, width, height, width if line_length is None else line_length
    d.framesMem, d.outArgsMem, d.outArgsStride = xdm.MEM_HOST, xdm.MEM_HOST, C.sizeof(xdm.TargetOutArgsAlg)
    d.frames, d.frameStride, d.outArgsAlg = frames.ctypes.data, frames.strides[0], C.addressof(outs)
    ret = lib().trikb200_edgeLineBatch(C.byref(d))
    return ret, outs


def last_error():
    return lib().trikb200_lastError().decode()


def launch_count():
    return int(lib().trikb200_launchCount())


def default_line_length(kind, width):
    return width if kind in (xdm.KIND_OO, xdm.KIND_OL, xdm.KIND_OM) else 2 * width


def layout_of(kind):
    return "yuv422p" if kind in (xdm.KIND_OO, xdm.KIND_OL, xdm.KIND_OM) else "yuyv"


def default_params(kind, max_w=640, max_h=480, max_out=640):
    p = xdm.TRIK_VIDTRANSCODE_CV_Params()
    b = p.base
    b.size = C.sizeof(p)
    b.numOutputStreams = 1
    b.formatInput = xdm.FORMAT_YUV422P if layout_of(kind) == "yuv422p" else xdm.FORMAT_YUV422
    b.formatOutput[0] = xdm.FORMAT_RGB565X
    b.formatOutput[1] = xdm.FORMAT_UNKNOWN
    b.maxHeightInput, b.maxWidthInput = max_h, max_w
    b.maxFrameRateInput, b.maxBitRateInput = 60000, -1
    b.maxHeightOutput[0], b.maxHeightOutput[1] = max_out, -1
    b.maxWidthOutput[0], b.maxWidthOutput[1] = max_out, -1
    b.maxFrameRateOutput[0] = b.maxFrameRateOutput[1] = -1
    b.maxBitRateOutput[0] = b.maxBitRateOutput[1] = -1
    b.dataEndianness = xdm.XDM_BYTE
    return p


def dynamic_params(width, height, line_length, out_w, out_h, out_line):
    d = xdm.TRIK_VIDTRANSCODE_CV_DynamicParams()
    b = d.base
    b.size = C.sizeof(d)
    b.keepInputResolutionFlag[0], b.keepInputResolutionFlag[1] = 0, 1
    b.outputHeight[0], b.outputWidth[0] = out_h, out_w
    b.keepInputFrameRateFlag[0] = b.keepInputFrameRateFlag[1] = 1
    b.inputFrameRate = -1
    b.outputFrameRate[0] = b.outputFrameRate[1] = -1
    b.targetBitRate[0] = b.targetBitRate[1] = -1
    b.rateControl[0] = b.rateControl[1] = xdm.IVIDEO_NONE
    b.keepInputGOPFlag[0] = b.keepInputGOPFlag[1] = 1
    b.intraFrameInterval[0] = b.intraFrameInterval[1] = 1
    b.forceFrame[0] = b.forceFrame[1] = xdm.IVIDEO_NA_FRAME
    d.inputHeight, d.inputWidth, d.inputLineLength = height, width, line_length
    d.outputLineLength[0], d.outputLineLength[1] = out_line, -1
    return d


class Codec:
    """One codec instance driven through the function table, as Codec Engine would."""

    def __init__(self, kind, params=None, device=None, fxns=None):
        """fxns: optional POINTER(IVIDTRANSCODE_Fxns) of ANOTHER implementation of the same ABI (the
        boundary tests pass the host-built reference's TRIK_VIDTRANSCODE_CV_FXNS here, so that one
        driver exercises both sides); default: this library's table for `kind`."""
        if isinstance(kind, str):
            kind = xdm.KIND_OF[kind]
        self.kind = kind
        self.foreign = fxns is not None
        self.InArgs = xdm.in_args_type(kind)
        self.OutArgs = xdm.out_args_type(kind)
        if self.foreign:
            self.lib = None
            self.fxns = fxns.contents
        else:
            self.lib = lib()
            if device is not None and self.lib.trikb200_setDevice(device) != 0:
                raise TrikB200Error(last_error())
            self.fxns = self.lib.trikb200_fxns(kind).contents
            assert C.sizeof(self.InArgs) == self.lib.trikb200_sizeofInArgs(kind)
            assert C.sizeof(self.OutArgs) == self.lib.trikb200_sizeofOutArgs(kind)
        self.params = params if params is not None else default_params(kind)
        # alloc: the CALLER owns the records (vidtranscode_cv_fxns.c:85-102)
        self.memtab = (xdm.IALG_MemRec * 4)()
        self.nrec = self.fxns.ialg.algAlloc(C.addressof(self.params), None, self.memtab)
        self._bufs = []
        for i in range(self.nrec):
            buf = C.create_string_buffer(max(int(self.memtab[i].size), 1))
            self._bufs.append(buf)
            self.memtab[i].base = C.addressof(buf)
        self.handle = C.addressof(self._bufs[0])
        C.cast(self.handle, C.POINTER(xdm.IALG_Obj)).contents.fxns = C.addressof(self.fxns)
        self.init_result = self.fxns.ialg.algInit(self.handle, self.memtab, None, C.addressof(self.params))
        self.width = self.height = self.line_length = 0
        self.out_w = self.out_h = self.out_line = 0
        self.preview = None
        self._closed = False

    # ---- control ---------------------------------------------------------------------------------
    def control(self, cmd, dyn=None, status=None):
        if status is None:
            status = xdm.IVIDTRANSCODE_Status()
            status.size = C.sizeof(status)
        ret = self.fxns.control(self.handle, cmd, C.addressof(dyn) if dyn is not None else None, C.byref(status))
        return ret, status

    def set_params(self, width, height, line_length=None, out_w=None, out_h=None, out_line=None, dyn_size=None):
        line_length = default_line_length(self.kind, width) if line_length is None else line_length
        out_w = width if out_w is None else out_w
        out_h = height if out_h is None else out_h
        out_line = out_w * 2 if out_line is None else out_line
        dyn = dynamic_params(width, height, line_length, out_w, out_h, out_line)
        if dyn_size is not None:
            dyn.base.size = dyn_size
        ret, _ = self.control(xdm.XDM_SETPARAMS, dyn)
        if ret == 0:
            self.width, self.height, self.line_length = width, height, line_length
            self.out_w, self.out_h, self.out_line = out_w, out_h, out_line
            self.preview = np.zeros(max(out_h * out_line, 16), dtype=np.uint8)
        return ret

    def get_version(self):
        status = xdm.IVIDTRANSCODE_Status()
        status.size = C.sizeof(status)
        buf = C.create_string_buffer(32)
        status.data.buf = C.addressof(buf)
        status.data.bufSize = 32
        ret, _ = self.control(xdm.XDM_GETVERSION, None, status)
        return ret, buf.value.decode()

    def set_seed(self, seed):
        if not self.foreign:
            self.lib.trikb200_setSeed(self.handle, int(seed))

    # ---- process ---------------------------------------------------------------------------------
    def process_raw(self, in_bufs, out_bufs, in_args, out_args):
        return self.fxns.process(self.handle, C.byref(in_bufs), C.byref(out_bufs), C.addressof(in_args), C.addressof(out_args))

    def process(self, frame, in_alg, out_alg=None, seed=None, num_bytes=None, preview=None):
        """One process() call with HOST buffers.  Returns (ret, OutArgs) -- OutArgs.alg holds the
        sensor result, OutArgs.base the xDM bookkeeping."""
        assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"]
        if seed is not None:
            self.set_seed(seed)
        preview = self.preview if preview is None else preview
        in_bufs = xdm.XDM1_BufDesc()
        in_bufs.numBufs = 1
        in_bufs.descs[0].buf = frame.ctypes.data
        in_bufs.descs[0].bufSize = frame.nbytes
        ptrs = (C.c_void_p * 1)(preview.ctypes.data)
        sizes = (C.c_int32 * 1)(preview.nbytes)
        out_bufs = xdm.XDM_BufDesc(ptrs, 1, sizes)
        ia = self.InArgs()
        ia.base.size = C.sizeof(ia)
        ia.base.numBytes = frame.nbytes if num_bytes is None else num_bytes
        ia.base.inputID = 1
        ia.alg = in_alg
        oa = self.OutArgs()
        oa.base.size = C.sizeof(oa)
        if out_alg is not None:
            oa.alg = out_alg
        ret = self.process_raw(in_bufs, out_bufs, ia, oa)
        return ret, oa

    def process_batch(self, frames, in_algs, out_algs=None, seeds=None, frames_device=False, frame_stride=None,
                      num_frames=None, out_device_ptr=None, stream=None, flags=0, stream_ids=None, num_streams=0,
                      previews=None, previews_device_ptr=None, preview_stride=None, multi=None):
        """n frames == n sequential process() calls.  multi: a list of codecs of this kind and geometry (self first), one
        per GPU: the batch is cut into contiguous ranges, one per codec (trikb200_processBatchMulti).

        frames: (n, frame_bytes) uint8 numpy array, or a raw device pointer (int) with
        frames_device=True, frame_stride and num_frames.  in_algs: one InArgsAlg (broadcast) or a
        ctypes array of n.  Returns (ret, out_algs ctypes array)."""
        InAlg, OutAlg = xdm.IN_ARGS_ALG[self.kind], xdm.OUT_ARGS_ALG[self.kind]
        b = xdm.Batch()
        b.size = C.sizeof(b)
        if frames_device:
            b.frames, b.frameStride, b.numFrames = int(frames), int(frame_stride), int(num_frames)
            b.framesMem = xdm.MEM_DEVICE
        else:
            assert frames.dtype == np.uint8 and frames.ndim == 2 and frames.flags["C_CONTIGUOUS"]
            b.frames, b.frameStride, b.numFrames = frames.ctypes.data, frames.strides[0], frames.shape[0]
            b.framesMem = xdm.MEM_HOST
        n = b.numFrames
        if isinstance(in_algs, InAlg):
            b.inArgsAlg, b.inArgsStride = C.addressof(in_algs), 0
        else:
            assert len(in_algs) == n
            b.inArgsAlg, b.inArgsStride = C.addressof(in_algs), C.sizeof(InAlg)
        if out_device_ptr is not None:
            b.outArgsAlg, b.outArgsMem = int(out_device_ptr), xdm.MEM_DEVICE
            out_algs = None
        else:
            if out_algs is None:
                out_algs = (OutAlg * n)()
            b.outArgsAlg, b.outArgsMem = C.addressof(out_algs), xdm.MEM_HOST
        b.outArgsStride = C.sizeof(OutAlg)
        keep = None
        if seeds is not None:
            keep = (C.c_int64 * n)(*[int(s) for s in seeds])
            b.seeds = C.addressof(keep)
        b.stream = stream
        b.flags = flags
        if previews is not None:
            assert previews.dtype == np.uint8 and previews.ndim == 2 and previews.flags["C_CONTIGUOUS"]
            b.previews, b.previewStride, b.previewsMem = previews.ctypes.data, previews.strides[0], xdm.MEM_HOST
        elif previews_device_ptr is not None:
            b.previews, b.previewStride, b.previewsMem = int(previews_device_ptr), int(preview_stride), xdm.MEM_DEVICE
        keep_ids = None
        if stream_ids is not None:
            keep_ids = (C.c_int32 * n)(*[int(x) for x in stream_ids])
            b.streamIds, b.numStreams = C.addressof(keep_ids), int(num_streams)
        if multi is not None:
            tab = (C.c_void_p * len(multi))(*[c.handle for c in multi])
            ret = self.lib.trikb200_processBatchMulti(tab, len(multi), C.byref(b))
        else:
            ret = self.lib.trikb200_processBatch(self.handle, C.byref(b))
        return ret, out_algs

    def synchronize(self):
        return self.lib.trikb200_synchronize(self.handle)

    # ---- free ------------------------------------------------------------------------------------
    def close(self):
        if not self._closed:
            tab = (xdm.IALG_MemRec * 4)()
            self.free_records = self.fxns.ialg.algFree(self.handle, tab)
            self.free_table = tab
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open_sensor(kind, width, height, line_length=None, out_w=None, out_h=None, device=None):
    """Codec with Params large enough for width x height and the geometry already set."""
    if isinstance(kind, str):
        kind = xdm.KIND_OF[kind]
    c = Codec(kind, default_params(kind, max(640, width), max(480, height), max(640, width, height)), device=device)
    if c.init_result != 0:
        raise TrikB200Error("initObj failed: %s" % last_error())
    if c.set_params(width, height, line_length, out_w, out_h) != 0:
        raise TrikB200Error("control(XDM_SETPARAMS) rejected %dx%d" % (width, height))
    return c


def process_mixed(items):
    """items: list of (codec, frame ndarray, in_alg, out_alg, seed) -- one frame each, any mix of sensors.
    Equivalent to calling codec.process() on every item in order (trikb200_processMixed)."""
    n = len(items)
    entries = (xdm.MixedEntry * n)()
    for e, (codec, frame, ia, oa, seed) in zip(entries, items):
        assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"]
        e.handle, e.frame = codec.handle, frame.ctypes.data
        e.inArgsAlg, e.outArgsAlg, e.seed = C.addressof(ia), C.addressof(oa), -1 if seed is None else int(seed)
    return lib().trikb200_processMixed(entries, n)
