/*
 * TEST INFRASTRUCTURE (oracle/): driver compiled INTO each oracle/_ref/libtrikref_<kind>.so
 * next to the reference's own, unmodified src/vidtranscode_cv.cpp and
 * src/vidtranscode_cv_fxns.c.  It walks the reference's codec function table exactly the
 * way Codec Engine would:
 *     alloc -> initObj -> control(XDM_SETPARAMS) -> process ... -> free
 * and flattens that into a few plain-C calls so that tests and bench.py can drive the
 * reference through ctypes without knowing the xDM struct layouts.
 *
 * One shared object hosts ONE live codec instance: the reference keeps frame-sized
 * file-scope statics (e.g. webcam/object_sensor/.../cv_ball_detector_seqpass.hpp:24-26)
 * and is neither re-entrant nor multi-instance safe.
 *
 * time() is redirected (-Wl,--wrap=time) so that the srand(time(NULL)) at
 * ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:180 becomes a
 * reproducible, caller-chosen seed.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "trik_vidtranscode_cv.h"

static time_t s_fake_time = 0;

time_t __wrap_time(time_t* t)
{
  if (t)
    *t = s_fake_time;
  return s_fake_time;
}

static IALG_MemRec s_memTab[IALG_DEFMEMRECS];
static IALG_Handle s_handle = NULL;
static int         s_numRecs = 0;
static void*       s_fastRam = NULL;
static size_t      s_fastRamSize = 0;
static TRIK_VIDTRANSCODE_CV_DynamicParams s_dyn;

void trikref_set_time(long long t) { s_fake_time = (time_t)t; }

int trikref_sizeof_inargs_alg(void)  { return (int)sizeof(TRIK_VIDTRANSCODE_CV_InArgsAlg); }
int trikref_sizeof_outargs_alg(void) { return (int)sizeof(TRIK_VIDTRANSCODE_CV_OutArgsAlg); }
int trikref_sizeof_inargs(void)      { return (int)sizeof(TRIK_VIDTRANSCODE_CV_InArgs); }
int trikref_sizeof_outargs(void)     { return (int)sizeof(TRIK_VIDTRANSCODE_CV_OutArgs); }
int trikref_sizeof_dynparams(void)   { return (int)sizeof(TRIK_VIDTRANSCODE_CV_DynamicParams); }
int trikref_sizeof_params(void)      { return (int)sizeof(TRIK_VIDTRANSCODE_CV_Params); }

void trikref_destroy(void)
{
  int i;
  if (s_handle)
  {
    IALG_MemRec freeTab[IALG_DEFMEMRECS];
    memset(freeTab, 0, sizeof(freeTab));
    TRIK_VIDTRANSCODE_CV_FXNS.ialg.algFree(s_handle, freeTab);
    s_handle = NULL;
  }
  /* memTab[1] (the 0x1000-byte "fast RAM") is deliberately never freed: the reference caches
   * pointers to its two division LUTs in class statics on FIRST instance creation
   * (webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:390-398), so they would
   * dangle if a later instance outlived the first one's fast RAM. */
  for (i = 0; i < s_numRecs; ++i)
  {
    if (i != 1)
      free(s_memTab[i].base);
    s_memTab[i].base = NULL;
  }
  s_numRecs = 0;
}

/* returns 0 on success; 1 alloc failure, 2 initObj failure (the value of init is in *detail),
 * 3 control(SETPARAMS) failure */
int trikref_create(int formatInput, int maxW, int maxH,
                   int w, int h, int lineLength,
                   int outW, int outH, int outLineLength, int* detail)
{
  TRIK_VIDTRANSCODE_CV_Params params;
  IVIDTRANSCODE_Status status;
  IALG_Fxns* parentFxns = NULL;
  int i, res;
  int outMax = maxH > maxW ? maxH : maxW;

  if (outMax < 320)
    outMax = 320;
  trikref_destroy();
  if (detail) *detail = 0;

  memset(&params, 0, sizeof(params));
  params.base.size               = sizeof(params);
  params.base.numOutputStreams   = 1;
  params.base.formatInput        = formatInput;
  params.base.formatOutput[0]    = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X;
  params.base.formatOutput[1]    = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_UNKNOWN;
  params.base.maxHeightInput     = maxH;
  params.base.maxWidthInput      = maxW;
  params.base.maxFrameRateInput  = 60000;
  params.base.maxBitRateInput    = -1;
  params.base.maxHeightOutput[0] = outMax; /* line sensors default to a 240x320 output */
  params.base.maxHeightOutput[1] = -1;
  params.base.maxWidthOutput[0]  = outMax;
  params.base.maxWidthOutput[1]  = -1;
  params.base.maxFrameRateOutput[0] = params.base.maxFrameRateOutput[1] = -1;
  params.base.maxBitRateOutput[0]   = params.base.maxBitRateOutput[1]   = -1;
  params.base.dataEndianness     = XDM_BYTE;

  memset(s_memTab, 0, sizeof(s_memTab));
  s_numRecs = TRIK_VIDTRANSCODE_CV_FXNS.ialg.algAlloc((const IALG_Params*)&params, &parentFxns, s_memTab);
  if (s_numRecs <= 0 || s_numRecs > IALG_DEFMEMRECS)
  {
    s_numRecs = 0;
    return 1;
  }
  for (i = 0; i < s_numRecs; ++i)
  {
    if (i == 1 && s_fastRam != NULL && s_fastRamSize >= s_memTab[i].size)
    {
      s_memTab[i].base = s_fastRam; /* keep the first instance's LUT storage alive, see trikref_destroy */
      continue;
    }
    if (posix_memalign(&s_memTab[i].base, 64, s_memTab[i].size ? s_memTab[i].size : 64) != 0)
      return 1;
    memset(s_memTab[i].base, 0, s_memTab[i].size);
    if (i == 1)
    {
      s_fastRam = s_memTab[i].base;
      s_fastRamSize = s_memTab[i].size;
    }
  }

  s_handle = (IALG_Handle)s_memTab[0].base;
  s_handle->fxns = &TRIK_VIDTRANSCODE_CV_FXNS.ialg;
  res = TRIK_VIDTRANSCODE_CV_FXNS.ialg.algInit(s_handle, s_memTab, NULL, (const IALG_Params*)&params);
  if (res != IALG_EOK)
  {
    if (detail) *detail = res;
    /* initObj failed after trikCvHandleInit allocated the C++ object: still release it */
    trikref_destroy();
    return 2;
  }

  memset(&s_dyn, 0, sizeof(s_dyn));
  s_dyn.base.size                       = sizeof(s_dyn);
  s_dyn.base.keepInputResolutionFlag[0] = XDAS_FALSE;
  s_dyn.base.keepInputResolutionFlag[1] = XDAS_TRUE;
  s_dyn.base.outputHeight[0]            = outH;
  s_dyn.base.outputWidth[0]             = outW;
  s_dyn.base.keepInputFrameRateFlag[0]  = XDAS_TRUE;
  s_dyn.base.keepInputFrameRateFlag[1]  = XDAS_TRUE;
  s_dyn.base.inputFrameRate             = -1;
  s_dyn.base.outputFrameRate[0] = s_dyn.base.outputFrameRate[1] = -1;
  s_dyn.base.targetBitRate[0]   = s_dyn.base.targetBitRate[1]   = -1;
  s_dyn.base.rateControl[0]     = s_dyn.base.rateControl[1]     = IVIDEO_NONE;
  s_dyn.base.keepInputGOPFlag[0] = s_dyn.base.keepInputGOPFlag[1] = XDAS_TRUE;
  s_dyn.base.intraFrameInterval[0] = s_dyn.base.intraFrameInterval[1] = 1;
  s_dyn.base.forceFrame[0] = s_dyn.base.forceFrame[1] = IVIDEO_NA_FRAME;
  s_dyn.inputHeight         = h;
  s_dyn.inputWidth          = w;
  s_dyn.inputLineLength     = lineLength;
  s_dyn.outputLineLength[0] = outLineLength;
  s_dyn.outputLineLength[1] = -1;

  memset(&status, 0, sizeof(status));
  status.size = sizeof(status);
  res = TRIK_VIDTRANSCODE_CV_FXNS.control((IVIDTRANSCODE_Handle)s_handle, XDM_SETPARAMS,
                                          (IVIDTRANSCODE_DynamicParams*)&s_dyn, &status);
  if (res != IVIDTRANSCODE_EOK)
  {
    if (detail) *detail = res;
    return 3;
  }
  return 0;
}

/* One process() call.  inAlg/outAlg point at the sensor's InArgsAlg/OutArgsAlg; outAlg is
 * read AND written (fields the algorithm leaves alone keep the caller's bytes, exactly as
 * with the real codec).  Returns the process() return value; *extendedError gets the xDM
 * error word. */
int trikref_process(const void* frame, int numBytes, int bufSize,
                    const void* inAlg, void* outAlg,
                    void* preview, int previewSize,
                    int* extendedError, int* bitsGenerated)
{
  XDM1_BufDesc inBufs;
  XDM_BufDesc  outBufs;
  XDAS_Int8*   outPtrs[1];
  XDAS_Int32   outSizes[1];
  TRIK_VIDTRANSCODE_CV_InArgs  inArgs;
  TRIK_VIDTRANSCODE_CV_OutArgs outArgs;
  int res;

  if (!s_handle)
    return -100;

  memset(&inBufs, 0, sizeof(inBufs));
  inBufs.numBufs          = 1;
  inBufs.descs[0].buf     = (XDAS_Int8*)frame;
  inBufs.descs[0].bufSize = bufSize;

  outPtrs[0]  = (XDAS_Int8*)preview;
  outSizes[0] = previewSize;
  outBufs.bufs     = outPtrs;
  outBufs.numBufs  = 1;
  outBufs.bufSizes = outSizes;

  memset(&inArgs, 0, sizeof(inArgs));
  inArgs.base.size     = sizeof(inArgs);
  inArgs.base.numBytes = numBytes;
  inArgs.base.inputID  = 1;
  memcpy(&inArgs.alg, inAlg, sizeof(inArgs.alg));

  memset(&outArgs, 0, sizeof(outArgs));
  outArgs.base.size = sizeof(outArgs);
  memcpy(&outArgs.alg, outAlg, sizeof(outArgs.alg));

  res = TRIK_VIDTRANSCODE_CV_FXNS.process((IVIDTRANSCODE_Handle)s_handle, &inBufs, &outBufs,
                                          &inArgs.base, &outArgs.base);

  memcpy(outAlg, &outArgs.alg, sizeof(outArgs.alg));
  if (extendedError) *extendedError = outArgs.base.extendedError;
  if (bitsGenerated) *bitsGenerated = outArgs.base.bitsGenerated[0];
  return res;
}

/* Raw access for the boundary tests: the caller builds every xDM struct itself. */
void* trikref_fxns(void)   { return &TRIK_VIDTRANSCODE_CV_FXNS; }
void* trikref_handle(void) { return s_handle; }
