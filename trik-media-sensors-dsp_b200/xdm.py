"""ctypes mirror of include/trik_xdm.h and include/trik_b200.h (the reference-facing ABI).

Names and field order follow the C headers one to one; see those headers for the reference
file:line each type restates.
"""
import ctypes as C

XDM_MAX_IO_BUFFERS = 16
MAXOUT = 2

IALG_EOK, IALG_EFAIL = 0, -1
XDM_EOK, XDM_EFAIL, XDM_EUNSUPPORTED = 0, -1, -3
XDM_GETSTATUS, XDM_SETPARAMS, XDM_RESET, XDM_SETDEFAULT, XDM_FLUSH, XDM_GETBUFINFO, XDM_GETVERSION = range(7)
XDM_CORRUPTEDDATA, XDM_UNSUPPORTEDPARAM = 11, 14
XDM_BYTE = 1
XDM_CUSTOMENUMBASE = 0x100
IVIDEO_NA_FRAME = IVIDEO_NA_PICTURE = IVIDEO_CONTENTTYPE_NA = -1
IVIDEO_NONE = 4
IALG_EXTERNAL, IALG_DARAM0, IALG_PERSIST = 0x11, 0, 1

FORMAT_UNKNOWN = 0
FORMAT_RGB888 = XDM_CUSTOMENUMBASE
FORMAT_RGB565 = XDM_CUSTOMENUMBASE + 1
FORMAT_RGB565X = XDM_CUSTOMENUMBASE + 2
FORMAT_YUV444 = XDM_CUSTOMENUMBASE + 3
FORMAT_YUV422 = XDM_CUSTOMENUMBASE + 4
FORMAT_YUV422P = XDM_CUSTOMENUMBASE + 5

KIND_WO, KIND_WL, KIND_OO, KIND_OL, KIND_OM = range(5)
KIND_NAMES = ("wo", "wl", "oo", "ol", "om")
KIND_OF = {n: i for i, n in enumerate(KIND_NAMES)}

MEM_HOST, MEM_DEVICE = 0, 1
BATCH_ASYNC = 1
BATCH_DEVICE_TAIL = 2


class IALG_MemRec(C.Structure):
    _fields_ = [("size", C.c_uint), ("alignment", C.c_int), ("space", C.c_int), ("attrs", C.c_int),
                ("base", C.c_void_p)]


class IALG_Obj(C.Structure):
    _fields_ = [("fxns", C.c_void_p)]


class IALG_Fxns(C.Structure):
    _fields_ = [("implementationId", C.c_void_p),
                ("algActivate", C.c_void_p),
                ("algAlloc", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(IALG_MemRec))),
                ("algControl", C.c_void_p),
                ("algDeactivate", C.c_void_p),
                ("algFree", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(IALG_MemRec))),
                ("algInit", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(IALG_MemRec), C.c_void_p, C.c_void_p)),
                ("algMoved", C.c_void_p),
                ("algNumAlloc", C.c_void_p)]


class XDM_BufDesc(C.Structure):
    _fields_ = [("bufs", C.POINTER(C.c_void_p)), ("numBufs", C.c_int32), ("bufSizes", C.POINTER(C.c_int32))]


class XDM1_SingleBufDesc(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("bufSize", C.c_int32), ("accessMask", C.c_int32)]


class XDM1_BufDesc(C.Structure):
    _fields_ = [("numBufs", C.c_int32), ("descs", XDM1_SingleBufDesc * XDM_MAX_IO_BUFFERS)]


class XDM_AlgBufInfo(C.Structure):
    _fields_ = [("minNumInBufs", C.c_int32), ("minNumOutBufs", C.c_int32),
                ("minInBufSize", C.c_int32 * XDM_MAX_IO_BUFFERS), ("minOutBufSize", C.c_int32 * XDM_MAX_IO_BUFFERS)]


class IVIDTRANSCODE_Params(C.Structure):
    _fields_ = [("size", C.c_int32), ("numOutputStreams", C.c_int32), ("formatInput", C.c_int32),
                ("formatOutput", C.c_int32 * MAXOUT), ("maxHeightInput", C.c_int32), ("maxWidthInput", C.c_int32),
                ("maxFrameRateInput", C.c_int32), ("maxBitRateInput", C.c_int32),
                ("maxHeightOutput", C.c_int32 * MAXOUT), ("maxWidthOutput", C.c_int32 * MAXOUT),
                ("maxFrameRateOutput", C.c_int32 * MAXOUT), ("maxBitRateOutput", C.c_int32 * MAXOUT),
                ("dataEndianness", C.c_int32)]


class IVIDTRANSCODE_DynamicParams(C.Structure):
    _fields_ = [("size", C.c_int32), ("readHeaderOnlyFlag", C.c_int32),
                ("keepInputResolutionFlag", C.c_uint8 * MAXOUT),
                ("outputHeight", C.c_int32 * MAXOUT), ("outputWidth", C.c_int32 * MAXOUT),
                ("keepInputFrameRateFlag", C.c_uint8 * MAXOUT), ("inputFrameRate", C.c_int32),
                ("outputFrameRate", C.c_int32 * MAXOUT), ("targetBitRate", C.c_int32 * MAXOUT),
                ("rateControl", C.c_int32 * MAXOUT), ("keepInputGOPFlag", C.c_uint8 * MAXOUT),
                ("intraFrameInterval", C.c_int32 * MAXOUT), ("interFrameInterval", C.c_int32 * MAXOUT),
                ("forceFrame", C.c_int32 * MAXOUT), ("frameSkipTranscodeFlag", C.c_uint8 * MAXOUT)]


class TRIK_VIDTRANSCODE_CV_Params(C.Structure):
    _fields_ = [("base", IVIDTRANSCODE_Params)]


class TRIK_VIDTRANSCODE_CV_DynamicParams(C.Structure):
    _fields_ = [("base", IVIDTRANSCODE_DynamicParams), ("inputHeight", C.c_int32), ("inputWidth", C.c_int32),
                ("inputLineLength", C.c_int32), ("outputLineLength", C.c_int32 * MAXOUT)]


class IVIDTRANSCODE_InArgs(C.Structure):
    _fields_ = [("size", C.c_int32), ("numBytes", C.c_int32), ("inputID", C.c_int32)]


class IVIDTRANSCODE_Status(C.Structure):
    _fields_ = [("size", C.c_int32), ("extendedError", C.c_int32), ("data", XDM1_SingleBufDesc),
                ("bufInfo", XDM_AlgBufInfo)]


class IVIDTRANSCODE_OutArgs(C.Structure):
    _fields_ = [("size", C.c_int32), ("extendedError", C.c_int32), ("bitsConsumed", C.c_int32),
                ("bitsGenerated", C.c_int32 * MAXOUT), ("decodedPictureType", C.c_int32),
                ("decodedPictureStructure", C.c_int32), ("encodedPictureType", C.c_int32 * MAXOUT),
                ("encodedPictureStructure", C.c_int32 * MAXOUT), ("decodedHeight", C.c_int32),
                ("decodedWidth", C.c_int32), ("outputID", C.c_int32 * MAXOUT),
                ("inputFrameSkipTranscodeFlag", C.c_int32 * MAXOUT), ("encodedBuf", XDM1_SingleBufDesc * MAXOUT),
                ("outBufsInUseFlag", C.c_int32)]


PROCESS_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(XDM1_BufDesc), C.POINTER(XDM_BufDesc), C.c_void_p, C.c_void_p)
CONTROL_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(IVIDTRANSCODE_Status))


class IVIDTRANSCODE_Fxns(C.Structure):
    _fields_ = [("ialg", IALG_Fxns), ("process", PROCESS_FN), ("control", CONTROL_FN)]


# ---- per-sensor argument structs (include/trik_b200.h) -------------------------------------------
class RangeInArgsAlg(C.Structure):
    _fields_ = [("detectHueFrom", C.c_uint16), ("detectHueTo", C.c_uint16),
                ("detectSatFrom", C.c_uint8), ("detectSatTo", C.c_uint8),
                ("detectValFrom", C.c_uint8), ("detectValTo", C.c_uint8),
                ("autoDetectHsv", C.c_uint8)]


class TargetOutArgsAlg(C.Structure):
    _fields_ = [("targetX", C.c_int8), ("targetY", C.c_int8), ("targetSize", C.c_uint8),
                ("detectHue", C.c_uint16), ("detectHueTolerance", C.c_uint16),
                ("detectSat", C.c_uint16), ("detectSatTolerance", C.c_uint16),
                ("detectVal", C.c_uint16), ("detectValTolerance", C.c_uint16)]


class ObjInArgsAlg(C.Structure):
    _fields_ = [("setHsvRange", C.c_uint8),
                ("detectHue", C.c_uint16), ("detectHueTol", C.c_uint16),
                ("detectSat", C.c_uint8), ("detectSatTol", C.c_uint8),
                ("detectVal", C.c_uint8), ("detectValTol", C.c_uint8),
                ("autoDetectHsv", C.c_uint8)]


class XDAS_Target(C.Structure):
    _fields_ = [("x", C.c_int8), ("y", C.c_int8), ("size", C.c_uint8)]


class ObjOutArgsAlg(C.Structure):
    _fields_ = [("target", XDAS_Target * 8),
                ("detectHue", C.c_uint16), ("detectHueTolerance", C.c_uint16),
                ("detectSat", C.c_uint16), ("detectSatTolerance", C.c_uint16),
                ("detectVal", C.c_uint16), ("detectValTolerance", C.c_uint16)]


class MxnInArgsAlg(C.Structure):
    _fields_ = [("widthM", C.c_int32), ("heightN", C.c_int32)]


class MxnOutArgsAlg(C.Structure):
    _fields_ = [("outColor", C.c_int32 * 100)]


IN_ARGS_ALG = {KIND_WO: RangeInArgsAlg, KIND_WL: RangeInArgsAlg, KIND_OL: RangeInArgsAlg,
               KIND_OO: ObjInArgsAlg, KIND_OM: MxnInArgsAlg}
OUT_ARGS_ALG = {KIND_WO: TargetOutArgsAlg, KIND_WL: TargetOutArgsAlg, KIND_OL: TargetOutArgsAlg,
                KIND_OO: ObjOutArgsAlg, KIND_OM: MxnOutArgsAlg}


def in_args_type(kind):
    class InArgs(C.Structure):
        _fields_ = [("base", IVIDTRANSCODE_InArgs), ("alg", IN_ARGS_ALG[kind])]
    return InArgs


def out_args_type(kind):
    class OutArgs(C.Structure):
        _fields_ = [("base", IVIDTRANSCODE_OutArgs), ("alg", OUT_ARGS_ALG[kind])]
    return OutArgs


class Batch(C.Structure):
    _fields_ = [("size", C.c_int32), ("numFrames", C.c_int32), ("frames", C.c_void_p), ("frameStride", C.c_int64),
                ("framesMem", C.c_int32), ("inArgsAlg", C.c_void_p), ("inArgsStride", C.c_int32),
                ("outArgsAlg", C.c_void_p), ("outArgsStride", C.c_int32), ("outArgsMem", C.c_int32),
                ("seeds", C.c_void_p), ("stream", C.c_void_p), ("flags", C.c_int32),
                ("streamIds", C.c_void_p), ("numStreams", C.c_int32),
                ("previews", C.c_void_p), ("previewStride", C.c_int64), ("previewsMem", C.c_int32)]


class Ingest(C.Structure):
    """TRIKB200_Ingest (include/trik_b200.h): RGB565 -> YUV422P front end."""
    _fields_ = [("size", C.c_int32), ("numFrames", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("pixelFormat", C.c_int32), ("srcMem", C.c_int32), ("dstMem", C.c_int32),
                ("srcLineLength", C.c_int32), ("dstLineLength", C.c_int32),
                ("src", C.c_void_p), ("srcStride", C.c_int64), ("dst", C.c_void_p), ("dstStride", C.c_int64),
                ("stream", C.c_void_p)]


PIXEL_RGB565, PIXEL_RGB565X = 0, 1


class EdgeLineBatch(C.Structure):
    """TRIKB200_EdgeLineBatch (include/trik_b200.h): ov7670/edge_line_sensor as a batch operation."""
    _fields_ = [("size", C.c_int32), ("numFrames", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("lineLength", C.c_int32), ("framesMem", C.c_int32), ("outArgsMem", C.c_int32), ("outArgsStride", C.c_int32),
                ("frames", C.c_void_p), ("frameStride", C.c_int64), ("outArgsAlg", C.c_void_p), ("stream", C.c_void_p)]


class MixedEntry(C.Structure):
    _fields_ = [("handle", C.c_void_p), ("frame", C.c_void_p), ("inArgsAlg", C.c_void_p), ("outArgsAlg", C.c_void_p),
                ("seed", C.c_int64)]
