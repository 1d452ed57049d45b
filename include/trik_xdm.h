/*
 * trik_xdm.h -- clean-room restatement of the TI XDAIS / xDM 1.x types that the
 * reference's codec surface touches.
 *
 * The reference compiles against TI XDAIS 7.24 / Codec Engine 3.23 headers
 * (trik/ov7670/common/makefile:7-8: <xdc/std.h>, <ti/xdais/ialg.h>,
 * <ti/xdais/dm/ividtranscode.h>), which are NOT vendored in the reference tree
 * and are not in this image.  Only the fields and constants that the reference
 * actually reads or writes are restated here:
 *   - src/vidtranscode_cv_fxns.c:20-40,85-334  (IALG_*, XDM*, IVIDTRANSCODE_*)
 *   - src/vidtranscode_cv.cpp:153-266          (Params / DynamicParams field order,
 *                                               pinned by the positional default
 *                                               initialisers there)
 * Numeric values of the enums follow the published xDM 1.x interface; nothing in
 * the reference pins them (SURVEY.md section 8(b) "dependency").
 *
 * The same header serves the drop-in library (libtrikb200) and, through the thin
 * forwarding stubs under oracle/stubs/, the host build of the reference itself, so
 * both sides of every parity test agree on the ABI.
 */
#ifndef TRIK_XDM_H_
#define TRIK_XDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- xdc/std.h ---------------------------------------------------------- */
typedef int            Int;
typedef unsigned int   Uns;
typedef unsigned int   UInt;
typedef char           Char;
typedef char*          String;
typedef void*          Ptr;
typedef unsigned short Bool;
typedef int8_t         Int8;
typedef int16_t        Int16;
typedef int32_t        Int32;
typedef uint8_t        UInt8;
typedef uint16_t       UInt16;
typedef uint32_t       UInt32;
typedef uint8_t        Uint8;
typedef uint16_t       Uint16;
typedef uint32_t       Uint32;
#ifndef Void
#define Void void
#endif
#ifndef TRUE
#define TRUE  1
#define FALSE 0
#endif

/* ---- ti/xdais/xdas.h ---------------------------------------------------- */
typedef void     XDAS_Void;
typedef uint8_t  XDAS_Bool;
typedef int8_t   XDAS_Int8;
typedef uint8_t  XDAS_UInt8;
typedef int16_t  XDAS_Int16;
typedef uint16_t XDAS_UInt16;
typedef int32_t  XDAS_Int32;
typedef uint32_t XDAS_UInt32;
#define XDAS_TRUE  1
#define XDAS_FALSE 0

/* ---- ti/xdais/ialg.h ---------------------------------------------------- */
#define IALG_DEFMEMRECS 4
#define IALG_OBJMEMREC  0
#define IALG_EOK        0
#define IALG_EFAIL      (-1)

typedef enum IALG_MemAttrs {
    IALG_SCRATCH   = 0,
    IALG_PERSIST   = 1,
    IALG_WRITEONCE = 2
} IALG_MemAttrs;

typedef enum IALG_MemSpace {
    IALG_EPROG    = 0x18,
    IALG_IPROG    = 0x08,
    IALG_ESDATA   = 0x10,
    IALG_EXTERNAL = 0x11,
    IALG_DARAM0   = 0,
    IALG_DARAM1   = 1,
    IALG_SARAM    = 2,
    IALG_SARAM0   = 2,
    IALG_SARAM1   = 3,
    IALG_DARAM2   = 4,
    IALG_SARAM2   = 5
} IALG_MemSpace;

typedef struct IALG_MemRec {
    Uns           size;
    Int           alignment;
    IALG_MemSpace space;
    IALG_MemAttrs attrs;
    Void*         base;
} IALG_MemRec;

struct IALG_Fxns;
typedef struct IALG_Obj {
    struct IALG_Fxns* fxns;
} IALG_Obj;
typedef struct IALG_Obj* IALG_Handle;

typedef struct IALG_Params {
    Int size;
} IALG_Params;

typedef struct IALG_Status {
    Int size;
} IALG_Status;

typedef unsigned int IALG_Cmd;

typedef struct IALG_Fxns {
    Void* implementationId;
    Void  (*algActivate)(IALG_Handle);
    Int   (*algAlloc)(const IALG_Params*, struct IALG_Fxns**, IALG_MemRec*);
    Int   (*algControl)(IALG_Handle, IALG_Cmd, IALG_Status*);
    Void  (*algDeactivate)(IALG_Handle);
    Int   (*algFree)(IALG_Handle, IALG_MemRec*);
    Int   (*algInit)(IALG_Handle, const IALG_MemRec*, IALG_Handle, const IALG_Params*);
    Void  (*algMoved)(IALG_Handle, const IALG_MemRec*, IALG_Handle, const IALG_Params*);
    Int   (*algNumAlloc)(Void);
} IALG_Fxns;

/* ---- ti/xdais/dm/xdm.h -------------------------------------------------- */
#define XDM_MAX_IO_BUFFERS 16

#define XDM_EOK          0
#define XDM_EFAIL        (-1)
#define XDM_EUNSUPPORTED (-3)

typedef struct XDM_BufDesc {
    XDAS_Int8** bufs;
    XDAS_Int32  numBufs;
    XDAS_Int32* bufSizes;
} XDM_BufDesc;

typedef struct XDM_SingleBufDesc {
    XDAS_Int8* buf;
    XDAS_Int32 bufSize;
} XDM_SingleBufDesc;

typedef struct XDM1_SingleBufDesc {
    XDAS_Int8* buf;
    XDAS_Int32 bufSize;
    XDAS_Int32 accessMask;
} XDM1_SingleBufDesc;

typedef struct XDM1_BufDesc {
    XDAS_Int32         numBufs;
    XDM1_SingleBufDesc descs[XDM_MAX_IO_BUFFERS];
} XDM1_BufDesc;

typedef enum XDM_AccessMode {
    XDM_ACCESSMODE_READ  = 0,
    XDM_ACCESSMODE_WRITE = 1
} XDM_AccessMode;

#define XDM_ISACCESSMODE_READ(x)    (((x) >> XDM_ACCESSMODE_READ) & 0x1)
#define XDM_ISACCESSMODE_WRITE(x)   (((x) >> XDM_ACCESSMODE_WRITE) & 0x1)
#define XDM_CLEARACCESSMODE_READ(x)  ((x) &= (~(0x1 << XDM_ACCESSMODE_READ)))
#define XDM_CLEARACCESSMODE_WRITE(x) ((x) &= (~(0x1 << XDM_ACCESSMODE_WRITE)))
#define XDM_SETACCESSMODE_READ(x)    ((x) |= (0x1 << XDM_ACCESSMODE_READ))
#define XDM_SETACCESSMODE_WRITE(x)   ((x) |= (0x1 << XDM_ACCESSMODE_WRITE))

typedef struct XDM_AlgBufInfo {
    XDAS_Int32 minNumInBufs;
    XDAS_Int32 minNumOutBufs;
    XDAS_Int32 minInBufSize[XDM_MAX_IO_BUFFERS];
    XDAS_Int32 minOutBufSize[XDM_MAX_IO_BUFFERS];
} XDM_AlgBufInfo;

#define XDM_CUSTOMENUMBASE 0x100

typedef enum XDM_CmdId {
    XDM_GETSTATUS      = 0,
    XDM_SETPARAMS      = 1,
    XDM_RESET          = 2,
    XDM_SETDEFAULT     = 3,
    XDM_FLUSH          = 4,
    XDM_GETBUFINFO     = 5,
    XDM_GETVERSION     = 6,
    XDM_GETCONTEXTINFO = 7
} XDM_CmdId;

typedef enum XDM_ErrorBit {
    XDM_PARAMSCHANGE       = 8,
    XDM_APPLIEDCONCEALMENT = 9,
    XDM_INSUFFICIENTDATA   = 10,
    XDM_CORRUPTEDDATA      = 11,
    XDM_CORRUPTEDHEADER    = 12,
    XDM_UNSUPPORTEDINPUT   = 13,
    XDM_UNSUPPORTEDPARAM   = 14,
    XDM_FATALERROR         = 15
} XDM_ErrorBit;

#define XDM_ISCORRUPTEDDATA(x)     (((x) >> XDM_CORRUPTEDDATA) & 0x1)
#define XDM_ISUNSUPPORTEDPARAM(x)  (((x) >> XDM_UNSUPPORTEDPARAM) & 0x1)
#define XDM_SETCORRUPTEDDATA(x)    ((x) |= (0x1 << XDM_CORRUPTEDDATA))
#define XDM_SETUNSUPPORTEDPARAM(x) ((x) |= (0x1 << XDM_UNSUPPORTEDPARAM))
#define XDM_SETFATALERROR(x)       ((x) |= (0x1 << XDM_FATALERROR))

typedef enum XDM_DataFormat {
    XDM_BYTE  = 1,
    XDM_LE_16 = 2,
    XDM_LE_32 = 3,
    XDM_LE_64 = 4,
    XDM_BE_16 = 5,
    XDM_BE_32 = 6,
    XDM_BE_64 = 7
} XDM_DataFormat;

/* ---- ti/xdais/dm/ivideo.h ----------------------------------------------- */
typedef enum IVIDEO_FrameType {
    IVIDEO_NA_FRAME = -1,
    IVIDEO_I_FRAME  = 0,
    IVIDEO_P_FRAME  = 1,
    IVIDEO_B_FRAME  = 2,
    IVIDEO_IDR_FRAME = 3
} IVIDEO_FrameType;

typedef enum IVIDEO_PictureType {
    IVIDEO_NA_PICTURE = -1,
    IVIDEO_I_PICTURE  = 0,
    IVIDEO_P_PICTURE  = 1,
    IVIDEO_B_PICTURE  = 2
} IVIDEO_PictureType;

typedef enum IVIDEO_ContentType {
    IVIDEO_CONTENTTYPE_NA = -1,
    IVIDEO_PROGRESSIVE    = 0,
    IVIDEO_INTERLACED     = 1
} IVIDEO_ContentType;

typedef enum IVIDEO_RateControlPreset {
    IVIDEO_LOW_DELAY    = 1,
    IVIDEO_STORAGE      = 2,
    IVIDEO_TWOPASS      = 3,
    IVIDEO_NONE         = 4,
    IVIDEO_USER_DEFINED = 5
} IVIDEO_RateControlPreset;

/* ---- ti/xdais/dm/ividtranscode.h ---------------------------------------- */
#define IVIDTRANSCODE_EOK          XDM_EOK
#define IVIDTRANSCODE_EFAIL        XDM_EFAIL
#define IVIDTRANSCODE_EUNSUPPORTED XDM_EUNSUPPORTED
#define IVIDTRANSCODE_MAXOUTSTREAMS 2

struct IVIDTRANSCODE_Fxns;
typedef struct IVIDTRANSCODE_Obj {
    struct IVIDTRANSCODE_Fxns* fxns;
} IVIDTRANSCODE_Obj;
typedef struct IVIDTRANSCODE_Obj* IVIDTRANSCODE_Handle;

/* field order pinned by the positional initialiser at src/vidtranscode_cv.cpp:153-184 */
typedef struct IVIDTRANSCODE_Params {
    XDAS_Int32 size;
    XDAS_Int32 numOutputStreams;
    XDAS_Int32 formatInput;
    XDAS_Int32 formatOutput[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 maxHeightInput;
    XDAS_Int32 maxWidthInput;
    XDAS_Int32 maxFrameRateInput;
    XDAS_Int32 maxBitRateInput;
    XDAS_Int32 maxHeightOutput[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 maxWidthOutput[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 maxFrameRateOutput[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 maxBitRateOutput[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 dataEndianness;
} IVIDTRANSCODE_Params;

/* field order pinned by the positional initialiser at src/vidtranscode_cv.cpp:204-266 */
typedef struct IVIDTRANSCODE_DynamicParams {
    XDAS_Int32 size;
    XDAS_Int32 readHeaderOnlyFlag;
    XDAS_Bool  keepInputResolutionFlag[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 outputHeight[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 outputWidth[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Bool  keepInputFrameRateFlag[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 inputFrameRate;
    XDAS_Int32 outputFrameRate[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 targetBitRate[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 rateControl[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Bool  keepInputGOPFlag[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 intraFrameInterval[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 interFrameInterval[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32 forceFrame[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Bool  frameSkipTranscodeFlag[IVIDTRANSCODE_MAXOUTSTREAMS];
} IVIDTRANSCODE_DynamicParams;

typedef struct IVIDTRANSCODE_InArgs {
    XDAS_Int32 size;
    XDAS_Int32 numBytes;
    XDAS_Int32 inputID;
} IVIDTRANSCODE_InArgs;

typedef struct IVIDTRANSCODE_Status {
    XDAS_Int32         size;
    XDAS_Int32         extendedError;
    XDM1_SingleBufDesc data;
    XDM_AlgBufInfo     bufInfo;
} IVIDTRANSCODE_Status;

/* fields touched at src/vidtranscode_cv_fxns.c:216-261 */
typedef struct IVIDTRANSCODE_OutArgs {
    XDAS_Int32         size;
    XDAS_Int32         extendedError;
    XDAS_Int32         bitsConsumed;
    XDAS_Int32         bitsGenerated[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32         decodedPictureType;
    XDAS_Int32         decodedPictureStructure;
    XDAS_Int32         encodedPictureType[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32         encodedPictureStructure[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32         decodedHeight;
    XDAS_Int32         decodedWidth;
    XDAS_Int32         outputID[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32         inputFrameSkipTranscodeFlag[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDM1_SingleBufDesc encodedBuf[IVIDTRANSCODE_MAXOUTSTREAMS];
    XDAS_Int32         outBufsInUseFlag;
} IVIDTRANSCODE_OutArgs;

typedef IALG_Cmd IVIDTRANSCODE_Cmd;

typedef struct IVIDTRANSCODE_Fxns {
    IALG_Fxns  ialg;
    XDAS_Int32 (*process)(IVIDTRANSCODE_Handle handle, XDM1_BufDesc* inBufs,
                          XDM_BufDesc* outBufs, IVIDTRANSCODE_InArgs* inArgs,
                          IVIDTRANSCODE_OutArgs* outArgs);
    XDAS_Int32 (*control)(IVIDTRANSCODE_Handle handle, IVIDTRANSCODE_Cmd id,
                          IVIDTRANSCODE_DynamicParams* params,
                          IVIDTRANSCODE_Status* status);
} IVIDTRANSCODE_Fxns;

#ifdef __cplusplus
}
#endif

#endif /* TRIK_XDM_H_ */
