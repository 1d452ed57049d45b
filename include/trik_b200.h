/*
 * trik_b200.h -- C ABI of libtrikb200: the B200 drop-in for the per-frame pixel pipeline of
 * trikset/trik-media-sensors-dsp (webcam object/line sensors, ov7670 object/line/mxn sensors).
 *
 * Two surfaces:
 *
 *  1. The reference's own codec surface.  Every sensor of the reference exports the SAME two
 *     symbols, TRIK_VIDTRANSCODE_CV_FXNS and TRIK_VIDTRANSCODE_CV_IALG
 *     (<sensor>/trik_vidtranscode_cv.h:17-18, <sensor>/src/vidtranscode_cv_fxns.c:36-40,63-65),
 *     and one DSP server image holds one sensor.  Here one library holds all five, so the tables
 *     carry the sensor in their name (TRIKB200_<KIND>_FXNS); the five thin alias libraries
 *     libtrik_vidtranscode_cv_<kind>.so re-export them under the reference's names for a caller
 *     that links exactly one sensor, as the reference's callers do.
 *     Signatures, ownership, error codes and bookkeeping follow vidtranscode_cv_fxns.c:85-334
 *     line by line (see DESIGN.md section "Boundary" for the table).
 *
 *  2. A batch extension (new, additive).  process() rejects numBufs != 1
 *     (vidtranscode_cv_fxns.c:199-204), so many frames per call need a new entry point:
 *     trikb200_processBatch().  Its results are defined as those of n sequential process()
 *     calls on the same handle, including the per-handle state the reference carries between
 *     calls (ov7670 line sensor: m_hStart/m_hStop lag one frame; ov7670 object sensor: the HSV
 *     range persists while setHsvRange == 0).
 *
 * Plain pointers and sizes only: no CUDA or torch types in any signature.  A "device pointer"
 * is a raw CUDA device address; a "stream" is a cudaStream_t passed as void* (NULL = the
 * handle's own stream).
 */
#ifndef TRIK_B200_H_
#define TRIK_B200_H_

#include "trik_xdm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- image formats: <sensor>/trik_vidtranscode_cv.h:20-32 --------------------------------- */
typedef enum TRIK_VIDTRANSCODE_CV_VideoFormat {
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_UNKNOWN = 0,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB888 = XDM_CUSTOMENUMBASE,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV444,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422,   /* webcam sensors: interleaved Y0 U Y1 V */
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P,  /* ov7670 sensors: luma plane + interleaved chroma plane */
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB888HSV,
    TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_METABITMAP
} TRIK_VIDTRANSCODE_CV_VideoFormat;

/* ---- create-time and run-time parameters: <sensor>/trik_vidtranscode_cv.h:35-48 ------------ */
typedef struct TRIK_VIDTRANSCODE_CV_Params {
    IVIDTRANSCODE_Params base;
} TRIK_VIDTRANSCODE_CV_Params;

typedef struct TRIK_VIDTRANSCODE_CV_DynamicParams {
    IVIDTRANSCODE_DynamicParams base;
    XDAS_Int32 inputHeight;
    XDAS_Int32 inputWidth;
    XDAS_Int32 inputLineLength;   /* byte stride of one row; YUV422P: of each plane, chroma plane at inputLineLength*inputHeight */
    XDAS_Int32 outputLineLength[IVIDTRANSCODE_MAXOUTSTREAMS];
} TRIK_VIDTRANSCODE_CV_DynamicParams;

/* ---- per-frame arguments --------------------------------------------------------------------
 * The reference gives every sensor's structs the same names; here the three distinct layouts
 * get distinct names.  Layouts are byte-identical to the reference's. */

/* webcam object / webcam line / ov7670 line: webcam/line_sensor/trik_vidtranscode_cv.h:48-74 */
typedef struct TRIKB200_RangeInArgsAlg {
    XDAS_UInt16 detectHueFrom; /* [0..359] */
    XDAS_UInt16 detectHueTo;   /* [0..359] */
    XDAS_UInt8  detectSatFrom; /* [0..100] */
    XDAS_UInt8  detectSatTo;   /* [0..100] */
    XDAS_UInt8  detectValFrom; /* [0..100] */
    XDAS_UInt8  detectValTo;   /* [0..100] */
    XDAS_Bool   autoDetectHsv;
} TRIKB200_RangeInArgsAlg;

typedef struct TRIKB200_TargetOutArgsAlg {
    XDAS_Int8   targetX;    /* [-100..100] */
    XDAS_Int8   targetY;    /* [-100..100] */
    XDAS_UInt8  targetSize; /* [0..100] */
    XDAS_UInt16 detectHue;
    XDAS_UInt16 detectHueTolerance;
    XDAS_UInt16 detectSat;
    XDAS_UInt16 detectSatTolerance;
    XDAS_UInt16 detectVal;
    XDAS_UInt16 detectValTolerance;
} TRIKB200_TargetOutArgsAlg;

/* ov7670 object sensor: ov7670/object_sensor/trik_vidtranscode_cv.h:51-81 */
typedef struct TRIKB200_ObjInArgsAlg {
    XDAS_Bool   setHsvRange;
    XDAS_UInt16 detectHue;     /* [0..359] */
    XDAS_UInt16 detectHueTol;
    XDAS_UInt8  detectSat;     /* [0..100] */
    XDAS_UInt8  detectSatTol;
    XDAS_UInt8  detectVal;     /* [0..100] */
    XDAS_UInt8  detectValTol;
    XDAS_Bool   autoDetectHsv;
} TRIKB200_ObjInArgsAlg;

typedef struct XDAS_Target {
    XDAS_Int8  x;    /* [-100..100] */
    XDAS_Int8  y;    /* [-100..100] */
    XDAS_UInt8 size; /* [0..100] */
} XDAS_Target;

typedef struct TRIKB200_ObjOutArgsAlg {
    XDAS_Target target[8];
    XDAS_UInt16 detectHue;
    XDAS_UInt16 detectHueTolerance;
    XDAS_UInt16 detectSat;
    XDAS_UInt16 detectSatTolerance;
    XDAS_UInt16 detectVal;
    XDAS_UInt16 detectValTolerance;
} TRIKB200_ObjOutArgsAlg;

/* ov7670 mxn sensor: ov7670/mxn_sensor/trik_vidtranscode_cv.h:49-62 */
typedef struct TRIKB200_MxnInArgsAlg {
    XDAS_Int32 widthM;   /* number of cell ROWS (the reference swaps the names, cv_ball_detector_seqpass.hpp:585-586) */
    XDAS_Int32 heightN;  /* number of cell COLUMNS */
} TRIKB200_MxnInArgsAlg;

typedef struct TRIKB200_MxnOutArgsAlg {
    XDAS_Int32 outColor[100]; /* 0x00RRGGBB per cell, row-major; entries past widthM*heightN are left untouched */
} TRIKB200_MxnOutArgsAlg;

/* full xDM argument blocks (base + alg), as the reference's TRIK_VIDTRANSCODE_CV_InArgs/OutArgs */
typedef struct { IVIDTRANSCODE_InArgs  base; TRIKB200_RangeInArgsAlg   alg; } TRIKB200_RangeInArgs;
typedef struct { IVIDTRANSCODE_OutArgs base; TRIKB200_TargetOutArgsAlg alg; } TRIKB200_TargetOutArgs;
typedef struct { IVIDTRANSCODE_InArgs  base; TRIKB200_ObjInArgsAlg     alg; } TRIKB200_ObjInArgs;
typedef struct { IVIDTRANSCODE_OutArgs base; TRIKB200_ObjOutArgsAlg    alg; } TRIKB200_ObjOutArgs;
typedef struct { IVIDTRANSCODE_InArgs  base; TRIKB200_MxnInArgsAlg     alg; } TRIKB200_MxnInArgs;
typedef struct { IVIDTRANSCODE_OutArgs base; TRIKB200_MxnOutArgsAlg    alg; } TRIKB200_MxnOutArgs;

/* ---- sensor kinds --------------------------------------------------------------------------- */
typedef enum TRIKB200_Kind {
    TRIKB200_KIND_WO = 0, /* trik/webcam/object_sensor : YUV422,  Range/Target args */
    TRIKB200_KIND_WL = 1, /* trik/webcam/line_sensor   : YUV422,  Range/Target args */
    TRIKB200_KIND_OO = 2, /* trik/ov7670/object_sensor : YUV422P, Obj args          */
    TRIKB200_KIND_OL = 3, /* trik/ov7670/line_sensor   : YUV422P, Range/Target args */
    TRIKB200_KIND_OM = 4, /* trik/ov7670/mxn_sensor    : YUV422P, Mxn args          */
    TRIKB200_KIND_COUNT = 5
} TRIKB200_Kind;

/* ---- surface 1: the codec function tables ---------------------------------------------------
 * Replaces TRIK_VIDTRANSCODE_CV_FXNS / TRIK_VIDTRANSCODE_CV_IALG of <sensor>/src/vidtranscode_cv_fxns.c:36-65.
 *   ialg.algAlloc  (fxns.c:85-102)  2 records: {sizeof(handle), EXTERNAL, PERSIST}, {0x1000, DARAM0, PERSIST}
 *   ialg.algInit   (fxns.c:146-166) + creates the device buffers and the CUDA stream owned by the handle
 *   ialg.algFree   (fxns.c:114-136) returns the same two records with base filled
 *   process        (fxns.c:174-264) one frame, HOST buffers; writes the RGB565X preview with overlays into outBufs->bufs[0]
 *   control        (fxns.c:272-334) XDM_GETSTATUS/GETBUFINFO/SETPARAMS/RESET/SETDEFAULT/FLUSH/GETVERSION */
extern IVIDTRANSCODE_Fxns TRIKB200_WO_FXNS;
extern IVIDTRANSCODE_Fxns TRIKB200_WL_FXNS;
extern IVIDTRANSCODE_Fxns TRIKB200_OO_FXNS;
extern IVIDTRANSCODE_Fxns TRIKB200_OL_FXNS;
extern IVIDTRANSCODE_Fxns TRIKB200_OM_FXNS;
extern IALG_Fxns TRIKB200_WO_IALG;
extern IALG_Fxns TRIKB200_WL_IALG;
extern IALG_Fxns TRIKB200_OO_IALG;
extern IALG_Fxns TRIKB200_OL_IALG;
extern IALG_Fxns TRIKB200_OM_IALG;

/* table lookup by kind (NULL if kind is out of range) */
IVIDTRANSCODE_Fxns* trikb200_fxns(XDAS_Int32 kind);

/* sizes a caller needs without including this header (ctypes / cgo / JNI bindings) */
XDAS_Int32 trikb200_sizeofInArgsAlg(XDAS_Int32 kind);
XDAS_Int32 trikb200_sizeofOutArgsAlg(XDAS_Int32 kind);
XDAS_Int32 trikb200_sizeofInArgs(XDAS_Int32 kind);
XDAS_Int32 trikb200_sizeofOutArgs(XDAS_Int32 kind);
XDAS_Int32 trikb200_sizeofHandle(void);

/* ---- surface 2: batch extension --------------------------------------------------------------
 * Convenience constructor: alloc -> malloc the records -> algInit -> control(XDM_SETPARAMS) for a
 * width x height input whose preview output is outWidth x outHeight (0,0 = same as input).
 * lineLength 0 = tight (2*width for YUV422, width for YUV422P).  Limits of the reference's static
 * buffers (640x480) do not apply: maxWidth/maxHeight default to the requested size.
 * Returns NULL on failure (same conditions as initObj/control failing). */
IVIDTRANSCODE_Handle trikb200_create(XDAS_Int32 kind, XDAS_Int32 width, XDAS_Int32 height,
                                     XDAS_Int32 lineLength, XDAS_Int32 outWidth, XDAS_Int32 outHeight);
void trikb200_delete(IVIDTRANSCODE_Handle handle);

#define TRIKB200_MEM_HOST   0   /* frames / results are host memory (pageable or pinned) */
#define TRIKB200_MEM_DEVICE 1   /* frames / results are CUDA device memory on the handle's device */

#define TRIKB200_BATCH_ASYNC 1  /* enqueue only: results land in stream order; needs results in
                                   device or pinned host memory and no frame that needs the host
                                   annealing tail (autoDetectHsv on WL/OL/OO), else XDM_EFAIL */

#define TRIKB200_BATCH_DEVICE_TAIL 2  /* run the annealing tail of autoDetectHsv (WL/OL/OO) on the device instead of the
                                   host: lifts the TRIKB200_BATCH_ASYNC restriction above.  Same generator and
                                   arithmetic; only pow() is the device's (<= 2 ulp) instead of libm's, see
                                   DESIGN.md 3.6 for the measured agreement */

typedef struct TRIKB200_Batch {
    XDAS_Int32  size;          /* sizeof(TRIKB200_Batch) */
    XDAS_Int32  numFrames;
    const void* frames;        /* frame i at frames + i*frameStride */
    int64_t     frameStride;   /* bytes, multiple of 16 */
    XDAS_Int32  framesMem;     /* TRIKB200_MEM_* */
    const void* inArgsAlg;     /* numFrames InArgsAlg of the handle's kind, or ONE when inArgsStride == 0 (host memory) */
    XDAS_Int32  inArgsStride;  /* bytes between consecutive InArgsAlg, 0 = broadcast the first */
    void*       outArgsAlg;    /* numFrames OutArgsAlg of the handle's kind */
    XDAS_Int32  outArgsStride; /* bytes between consecutive OutArgsAlg (>= sizeof) */
    XDAS_Int32  outArgsMem;    /* TRIKB200_MEM_*; with MEM_DEVICE every field the sensor does not produce is written as 0 */
    const int64_t* seeds;      /* per-frame srand() seed for the annealed auto-detect, NULL = time(NULL) as the reference */
    void*       stream;        /* cudaStream_t, NULL = the handle's stream */
    XDAS_Int32  flags;         /* TRIKB200_BATCH_* */
    const XDAS_Int32* streamIds; /* optional: frame i belongs to logical stream streamIds[i] in [0, numStreams).  Every
                                  stream has its own carried state inside the handle (as if it were its own codec
                                  instance of this geometry), frames of one stream are taken in batch order; one launch
                                  serves all streams.  NULL = one stream, the handle itself.  Stream states are reset
                                  by control(XDM_SETPARAMS) like the handle's own. */
    XDAS_Int32  numStreams;
    void*       previews;      /* optional: RGB565X preview images with overlays, frame i at previews + i*previewStride,
                                  each outputHeight * outputLineLength bytes as process() produces them; NULL = none */
    int64_t     previewStride;
    XDAS_Int32  previewsMem;   /* TRIKB200_MEM_* */
} TRIKB200_Batch;

/* n frames through one handle == n sequential process() calls (without the preview image).
 * Returns IVIDTRANSCODE_EOK, or IVIDTRANSCODE_EFAIL (bad pointers / sizes / CUDA error) with
 * nothing written to outArgsAlg. */
XDAS_Int32 trikb200_processBatch(IVIDTRANSCODE_Handle handle, const TRIKB200_Batch* batch);

/* The same over several GPUs of one box (frames shard by batch, no collective on the per-pixel path):
 * handles[0..numHandles) are distinct instances of ONE sensor kind and geometry, each created after
 * trikb200_setDevice(d) for the device it shall run on (two handles may share a device).  The batch is cut into
 * numHandles contiguous frame ranges, range d runs on handles[d] from its own host thread and stream, and every
 * result lands in the caller's one outArgsAlg array.  Carried state (ov7670 line sensor band, object sensor range)
 * is walked over the whole batch first, so the results are those of numFrames sequential process() calls on
 * handles[0]; afterwards every handle holds the state after the last frame.  Host frames / results / previews
 * only, synchronous, no caller stream, no streamIds (IVIDTRANSCODE_EFAIL otherwise).  Replaces N instances of the
 * reference's codec behind N Codec Engine servers (dsp_server/server.cfg:139-142). */
XDAS_Int32 trikb200_processBatchMulti(const IVIDTRANSCODE_Handle* handles, XDAS_Int32 numHandles,
                                      const TRIKB200_Batch* batch);

/* Mixed sensors in one call (BASELINE.json config 4: line + object + mxn instances over many streams):
 * a table of {handle, frame} pairs with HOST frames.  Entries of one handle are processed in table order
 * (== sequential process() calls on that handle).  Handles of the same sensor kind, device and geometry
 * (the reference's model is one codec instance per camera, dsp_server/server.cfg:139-142) share ONE launch
 * per kernel, each frame judged with the carried state of its own handle; the classes run concurrently on
 * their own CUDA streams.  Pinned (cudaHostAlloc / cudaHostRegister) frames are read in place by one gather
 * kernel per class, pageable ones go through one copy per frame (neighbours in memory merged).
 * seed: srand() seed for an annealed auto-detect (negative = time(NULL)). */
typedef struct TRIKB200_MixedEntry {
    IVIDTRANSCODE_Handle handle;
    const void*          frame;       /* host pointer to one frame of the handle's geometry */
    const void*          inArgsAlg;   /* InArgsAlg of the handle's kind */
    void*                outArgsAlg;  /* OutArgsAlg of the handle's kind */
    int64_t              seed;
} TRIKB200_MixedEntry;
XDAS_Int32 trikb200_processMixed(const TRIKB200_MixedEntry* entries, XDAS_Int32 numEntries);

/* Ingest front end: packed RGB565 camera frames -> the YUV422P layout the ov7670 sensors read
 * (luma plane, then at dstLineLength * height the plane of interleaved chroma bytes V U V U ...), so that an RGB565 stream
 * can feed trikb200_processBatch of an ov7670 handle (frames = dst, frameStride = dstStride, framesMem = dstMem).
 * Runs on the calling thread's current CUDA device (trikb200_setDevice); needs no handle.
 * The reference itself has no RGB565 INPUT (its ov7670 sensors accept YUV422P only, src/vidtranscode_cv.cpp:76-84;
 * RGB565X is the format of the preview they write), so the arithmetic is defined here and pinned by nothing in the
 * reference: integer BT.601 studio swing, the inverse of the reference's own YUV -> RGB matrix --
 *   R8 = r5<<3 | r5>>2, G8 = g6<<2 | g6>>4, B8 = b5<<3 | b5>>2,
 *   Y = ((66R + 129G + 25B + 128) >> 8) + 16; chroma of a pixel pair from its summed channels Rs = R0 + R1, Gs, Bs:
 *   Cb = ((-38Rs - 74Gs + 112Bs + 256) >> 9) + 128, Cr = ((112Rs - 94Gs - 18Bs + 256) >> 9) + 128 (shifts floor). */
#define TRIKB200_PIXEL_RGB565   0   /* 16-bit little-endian words: R in bits 15..11, G 10..5, B 4..0 */
#define TRIKB200_PIXEL_RGB565X  1   /* B in bits 15..11, G 10..5, R 4..0: the words the sensors' preview images hold
                                       (writeOutputPixel, webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:66-70) */
typedef struct TRIKB200_Ingest {
    XDAS_Int32  size;            /* sizeof(TRIKB200_Ingest) */
    XDAS_Int32  numFrames;
    XDAS_Int32  width, height;   /* pixels, width % 8 == 0 */
    XDAS_Int32  pixelFormat;     /* TRIKB200_PIXEL_* */
    XDAS_Int32  srcMem, dstMem;  /* TRIKB200_MEM_* */
    XDAS_Int32  srcLineLength;   /* bytes per source row, >= 2 * width, multiple of 16 */
    XDAS_Int32  dstLineLength;   /* bytes per row of each destination plane, >= width, multiple of 8 */
    const void* src;             /* frame i at src + i * srcStride (16-byte aligned) */
    int64_t     srcStride;
    void*       dst;             /* frame i at dst + i * dstStride (8-byte aligned), 2 * dstLineLength * height bytes each */
    int64_t     dstStride;
    void*       stream;          /* cudaStream_t, NULL = the default stream.  With device memory on both sides the
                                    conversion is only enqueued; with host memory the call returns when dst is complete */
} TRIKB200_Ingest;
XDAS_Int32 trikb200_ingestRgb565(const TRIKB200_Ingest* ingest);

/* ov7670/edge_line_sensor as a batch operation: targetX / targetY / targetSize of
 * ov7670/edge_line_sensor/include/internal/cv_ball_detector_seqpass.hpp (Sobel 3x3 of the luma plane, threshold 50, centroid
 * column of the edge pixels in columns 16 .. width-16; :176-205, :386-414) for every frame, written as
 * TRIKB200_TargetOutArgsAlg records (the six detect* fields are written as 0; the reference leaves them alone).
 * Only the luma plane is read: frames may be whole YUV422P frames (frameStride = 2 * lineLength * height) or bare planes.
 * PARITY UNPINNED: the sensor's Sobel / threshold kernels are TI IMGLIB (closed, absent from the reference tree); they
 * are taken from TI's published natural-C models (oracle/imglib_open.c).  Not a codec handle: the reference's preview for
 * this sensor goes through a third IMGLIB kernel (IMG_ycbcr422pl_to_rgb565) and is not produced.
 * width % 32 == 0, 32 <= width <= 1040, height % 4 == 0, lineLength >= width (the reference itself ignores inputLineLength
 * in its Sobel call, i.e. assumes lineLength == width). */
typedef struct TRIKB200_EdgeLineBatch {
    XDAS_Int32  size;            /* sizeof(TRIKB200_EdgeLineBatch) */
    XDAS_Int32  numFrames;
    XDAS_Int32  width, height;
    XDAS_Int32  lineLength;      /* bytes per luma row */
    XDAS_Int32  framesMem;       /* TRIKB200_MEM_* */
    XDAS_Int32  outArgsMem;      /* TRIKB200_MEM_* */
    XDAS_Int32  outArgsStride;   /* bytes between records, >= sizeof(TRIKB200_TargetOutArgsAlg) */
    const void* frames;          /* frame i at frames + i * frameStride */
    int64_t     frameStride;
    void*       outArgsAlg;
    void*       stream;          /* cudaStream_t, NULL = the default stream; device memory on both sides: enqueued only */
} TRIKB200_EdgeLineBatch;
XDAS_Int32 trikb200_edgeLineBatch(const TRIKB200_EdgeLineBatch* batch);

/* wait for everything enqueued on the handle (TRIKB200_BATCH_ASYNC) */
XDAS_Int32 trikb200_synchronize(IVIDTRANSCODE_Handle handle);

/* seed used by the next process() call that runs an annealed auto-detect (the reference calls
 * srand(time(NULL)), ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:180);
 * negative = back to time(NULL). */
void trikb200_setSeed(IVIDTRANSCODE_Handle handle, int64_t seed);

/* device selection for handles created afterwards on this thread (default: current CUDA device) */
XDAS_Int32 trikb200_deviceCount(void);
XDAS_Int32 trikb200_setDevice(XDAS_Int32 device);

/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t trikb200_launchCount(void);
/* tuning knob: CTAs per frame for the sum kernels (0 = heuristic) */
void trikb200_setSlabsPerFrame(XDAS_Int32 slabs);
/* tuning knob for A/B measurements of the sum kernels (WO/WL/OL); negative = the measured per-sensor defaults.
 * v % 100: load path, 0 = register prefetch, 2 or 4 = depth of the per-thread cp.async ring in shared memory;
 * v / 100: 0 = default kernel per sensor, 1 = first-version kernel, 2 = tuned line kernel (WL/OL).
 * Unsupported values make the next launch fail with XDM_EFAIL. */
void trikb200_setLoadStages(XDAS_Int32 stages);
/* tuning knob: target CTA size of the sum kernels (rounded to a whole number of rows per iteration), 0 = default */
void trikb200_setBlockThreads(XDAS_Int32 threads);
/* tuning knob: 1 (default) launches the line kernel with programmatic stream serialisation, so the next batch's
 * CTAs are placed while the previous batch drains (each kernel still waits for all earlier work of the stream
 * before its first memory access); 0 = plain launches */
void trikb200_setOverlapLaunch(XDAS_Int32 on);
/* tuning knob: webcam object sensor batches whose frames share one threshold set can run through a chroma-indexed
 * detection table (results identical): 0 = automatic (default), 1 = whenever possible, -1 = never */
void trikb200_setLutMode(XDAS_Int32 mode);
/* tuning knob: the mxn sensor's per-pixel colour bin (H>>3, S>>6, V>>6) is one fixed function of (Y,U,V), tabulated once
 * per device and process at the first mxn call (2^24 entries, 32 MB; results identical):
 * 0 (default) = majority pass first -- a cell in which one bin provably holds more than half of the pixels is decided
 * without a histogram -- and the table-gather histogram kernel for the cell rows that leaves undecided;
 * 1 = the histogram kernel for every cell row; -1 = no table at all (arithmetic kernel) */
void trikb200_setMxnTableMode(XDAS_Int32 mode);
/* tuning knob: scattered frames of trikb200_processMixed: 0 (default) = pinned host frames are read in place by one gather
 * kernel per class from 8 frames up, 1 = from one frame up, -1 = always one cudaMemcpyAsync per frame */
void trikb200_setGatherMode(XDAS_Int32 mode);
/* tuning knob: 1 (default) = the shared-memory detection table of the object sensors is laid out with skewed rows (260
 * instead of 256 bytes apart): no bank conflicts when the chroma of neighbouring pixels differs by a little (camera noise),
 * for one more instruction per pixel pair (+60..75 % on such frames, -4 % on frames with noise-free chroma); 0 = plain rows */
void trikb200_setLutSkew(XDAS_Int32 on);
/* tuning knob: work items of the webcam object sensor's table kernel: 0 (default) = whole frames, or 2 / 4 / 8 bands of rows
 * per frame when whole frames would leave the last round of the persistent groups mostly idle; 1 / 2 / 4 / 8 = fixed */
void trikb200_setLutParts(XDAS_Int32 parts);
/* tuning knob: a batch with 1:1 previews can go through the preview kernels in sub-batches whose images total about this
 * many MiB (so that the overlay kernel's scattered two-byte stores land on lines still in L2); measured slower than the whole
 * batch at once at every size tried, hence 0 (= the whole batch at once) is the default */
void trikb200_setPreviewChunkMB(XDAS_Int32 mb);
/* tuning knob, where the line sensors' overlays on a 1:1 preview are drawn: -1 (default) = inside the streaming preview
 * kernel, every CTA patching the 32-byte sectors the lines cross among the rows it has just written (still in L2, no second
 * launch); 1 = the same full-sector read-modify-writes as a kernel of their own; 0 = the generic overlay kernel's two-byte
 * stores; 2..8 = inside, by the last CTA of each frame to arrive (1, 2 .. 64 blocks of 256 items per CTA); 9..15 = inside,
 * CTA-local with a forced 1, 2 .. 64 blocks per CTA.  All byte-identical (tests/test_preview_gpu.py); measurements in
 * DESIGN.md 3.7 */
void trikb200_setPreviewSectorOverlay(XDAS_Int32 on);
/* tuning knob: 1 (default) = the 1:1 preview of a webcam object sensor batch that went through the chroma table of its one
 * threshold set detects through that table too (two byte look-ups per pixel pair), 0 = always the HSV arithmetic */
void trikb200_setPreviewTable(XDAS_Int32 on);
/* tuning knob: edge-line kernel, 0 = packed four-pixels-per-thread form (default, needs 4-byte aligned rows), 1 = one thread
 * per column (first version) */
void trikb200_setEdgeLineVariant(XDAS_Int32 variant);
/* tuning knob: target CTA size of the mxn table kernel, 0 = default */
void trikb200_setMxnTableThreads(XDAS_Int32 threads);
/* tuning knob: synchronous host-memory calls of up to this many frame bytes (default 1 MiB; process() is one frame)
 * are staged through the handle's pinned buffers by the CPU (DMA between pinned and device memory, result records
 * written in place by the kernels); 0 = hand the caller's pointers to cudaMemcpyAsync as they are */
void trikb200_setZeroCopyBytes(XDAS_Int32 bytes);
/* tuning knob: frames per CTA of the wide ov7670 line kernel on small frames: 0 = heuristic (default), 1 = one, n = n */
void trikb200_setFramesPerCta(XDAS_Int32 n);
/* last CUDA / argument error message of this thread ("" if none) */
const char* trikb200_lastError(void);

/* diagnostics for the parity tests: run the DEVICE pixel functions over a range of inputs.
 * which = 0: index = Y | U<<8 | V<<16 -> 0x00RRGGBB (bit 31 set if the YUYV and YUV422P lane
 * paths disagree); which = 1: index = 0x00RRGGBB -> 0x00VVSSHH (scalar form); which = 2: index = Y | U<<8 | V<<16
 * -> 0x00VVSSHH through the packed two-pixel path the sensors use (bit 31: lane paths disagree);
 * which = 3: index = Y | U<<8 | V<<16 -> the mxn sensor's colour bin (H>>3)<<4 | (S>>6)<<2 | V>>6 as its table holds it.
 * hostOut holds count words. */
XDAS_Int32 trikb200_probePixels(XDAS_Int32 which, uint32_t first, uint32_t count, uint32_t* hostOut);
/* diagnostics for the parity tests: build the chroma-indexed detection table of a webcam object sensor threshold
 * set and compare it with the arithmetic on all 2^24 (Y,U,V).  stats[0] = mismatching pixels (must be 0),
 * [1] / [2] / [3] = chroma entries that never pass / pass on one luma interval / need the mask, [4] = passing pixels. */
XDAS_Int32 trikb200_probeLut(const TRIKB200_RangeInArgsAlg* inArgsAlg, uint64_t stats[5]);

#ifdef __cplusplus
}
#endif
#endif /* TRIK_B200_H_ */
