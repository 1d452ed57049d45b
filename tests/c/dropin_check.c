/*
 * dropin_check.c -- a plain C caller, written the way a Codec Engine skeleton drives a sensor codec
 * (alloc -> initObj -> control(XDM_SETPARAMS) -> process -> free through an IVIDTRANSCODE_Fxns table),
 * run against TWO implementations of that table loaded with dlopen:
 *     argv[1]  the host build of the reference          (oracle/_ref/libtrikref_<kind>.so)
 *     argv[2]  this repository's drop-in alias library  (libtrik_vidtranscode_cv_<kind>.so)
 * Both export the reference's symbol TRIK_VIDTRANSCODE_CV_FXNS.  The SAME compiled driver code and the
 * SAME struct definitions (include/trik_b200.h) are used for both; OutArgs and the preview image must
 * come out byte for byte identical.  argv[3] = kind (wo|wl|ol|om|oo).  Exit code 0 = identical.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trik_b200.h"

#define W 320
#define H 240

typedef struct {
  int ret;
  unsigned char alg[400];
  int algSize;
  IVIDTRANSCODE_OutArgs base;
  unsigned char* preview;
} Result;

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

static void make_frame(unsigned char* f, int planar, int variant)
{
  uint32_t s = 12345u + (uint32_t)variant;
  int r, c;
  for (r = 0; r < H; ++r)
    for (c = 0; c < W; c += 2)
    {
      const int dark = (c > 100 + 20 * variant && c < 140 + 20 * variant);
      const unsigned char y0 = (unsigned char)(dark ? 10 + lcg(&s) % 20 : 150 + lcg(&s) % 60);
      const unsigned char y1 = (unsigned char)(dark ? 10 + lcg(&s) % 20 : 150 + lcg(&s) % 60);
      const unsigned char u = (unsigned char)(dark ? 128 : 90 + r / 4), v = (unsigned char)(dark ? 128 : 200 - c / 4);
      if (!planar)
      {
        unsigned char* p = f + (size_t)r * W * 2 + (size_t)c * 2;
        p[0] = y0; p[1] = u; p[2] = y1; p[3] = v;
      }
      else
      {
        f[(size_t)r * W + c] = y0; f[(size_t)r * W + c + 1] = y1;
        f[(size_t)(H + r) * W + c] = v; f[(size_t)(H + r) * W + c + 1] = u;
      }
    }
}

/* object sensor: a dark YUV422P frame with 14 saturated red blocks of different sizes (>= 8 labels: with fewer the
 * reference reads past its cluster vector, ov7670/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:575-589),
 * two of them of equal size so that the order std::sort leaves equal sizes in shows */
static void make_object_frame(unsigned char* f, int variant)
{
  int k, r, c;
  memset(f, 24, (size_t)W * H);
  memset(f + (size_t)W * H, 128, (size_t)W * H);
  for (k = 0; k < 14; ++k)
  {
    const int bh = 4 * (1 + (k == 13 ? 5 : k) % 7), bw = 4 * (1 + (k == 13 ? 5 : k) % 9);
    const int r0 = 4 * ((7 * k + 3 * variant) % ((H - 40) / 4)), c0 = 4 * ((11 * k * k + 5 * variant + 2 * k) % ((W - 48) / 4));
    for (r = r0; r < r0 + bh; ++r)
      for (c = c0; c < c0 + bw; c += 2)
      {
        f[(size_t)r * W + c] = 120; f[(size_t)r * W + c + 1] = 120;
        f[(size_t)(H + r) * W + c] = 240; f[(size_t)(H + r) * W + c + 1] = 90;
      }
  }
}

static int run(IVIDTRANSCODE_Fxns* fx, const char* kind, int nframes, Result* out)
{
  const int planar = strcmp(kind, "wo") != 0 && strcmp(kind, "wl") != 0;
  const int isMxn = strcmp(kind, "om") == 0;
  const int isObj = strcmp(kind, "oo") == 0;
  TRIK_VIDTRANSCODE_CV_Params params;
  TRIK_VIDTRANSCODE_CV_DynamicParams dyn;
  IVIDTRANSCODE_Status status;
  IALG_MemRec mem[IALG_DEFMEMRECS];
  IALG_Handle h;
  unsigned char* frame = NULL;
  int n, i, f;

  memset(&params, 0, sizeof(params));
  params.base.size = sizeof(params);
  params.base.numOutputStreams = 1;
  params.base.formatInput = planar ? TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P : TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422;
  params.base.formatOutput[0] = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X;
  params.base.maxHeightInput = 480; params.base.maxWidthInput = 640;
  params.base.maxFrameRateInput = 60000; params.base.maxBitRateInput = -1;
  params.base.maxHeightOutput[0] = 640; params.base.maxHeightOutput[1] = -1;
  params.base.maxWidthOutput[0] = 640; params.base.maxWidthOutput[1] = -1;
  params.base.maxFrameRateOutput[0] = params.base.maxFrameRateOutput[1] = -1;
  params.base.maxBitRateOutput[0] = params.base.maxBitRateOutput[1] = -1;
  params.base.dataEndianness = XDM_BYTE;

  memset(mem, 0, sizeof(mem));
  n = fx->ialg.algAlloc((const IALG_Params*)&params, NULL, mem);
  if (n != 2) return 10;
  for (i = 0; i < n; ++i)
  {
    if (posix_memalign(&mem[i].base, 64, mem[i].size ? mem[i].size : 64)) return 11;
    memset(mem[i].base, 0, mem[i].size);
  }
  h = (IALG_Handle)mem[0].base;
  h->fxns = &fx->ialg;
  if (fx->ialg.algInit(h, mem, NULL, (const IALG_Params*)&params) != IALG_EOK) return 12;

  memset(&dyn, 0, sizeof(dyn));
  dyn.base.size = sizeof(dyn);
  dyn.base.keepInputResolutionFlag[1] = XDAS_TRUE;
  dyn.base.outputHeight[0] = H / 2; dyn.base.outputWidth[0] = W / 2;
  dyn.base.keepInputFrameRateFlag[0] = dyn.base.keepInputFrameRateFlag[1] = XDAS_TRUE;
  dyn.base.inputFrameRate = -1;
  dyn.base.rateControl[0] = dyn.base.rateControl[1] = IVIDEO_NONE;
  dyn.base.forceFrame[0] = dyn.base.forceFrame[1] = IVIDEO_NA_FRAME;
  dyn.inputHeight = H; dyn.inputWidth = W; dyn.inputLineLength = planar ? W : 2 * W;
  dyn.outputLineLength[0] = W; dyn.outputLineLength[1] = -1;
  memset(&status, 0, sizeof(status));
  status.size = sizeof(status);
  if (fx->control((IVIDTRANSCODE_Handle)h, XDM_SETPARAMS, &dyn.base, &status) != IVIDTRANSCODE_EOK) return 13;

  if (posix_memalign((void**)&frame, 64, (size_t)W * H * 2)) return 14;
  for (f = 0; f < nframes; ++f)
  {
    XDM1_BufDesc inBufs;
    XDM_BufDesc outBufs;
    XDAS_Int8* outPtr[1];
    XDAS_Int32 outSize[1];
    union { TRIKB200_RangeInArgs r; TRIKB200_MxnInArgs m; TRIKB200_ObjInArgs o; } in;
    union { TRIKB200_TargetOutArgs t; TRIKB200_MxnOutArgs m; TRIKB200_ObjOutArgs o; } oa;
    Result* res = &out[f];
    if (isObj)
      make_object_frame(frame, f);
    else
      make_frame(frame, planar, f);
    res->preview = (unsigned char*)malloc((size_t)(W / 2) * (H / 2) * 2);
    memset(res->preview, 0x5A, (size_t)(W / 2) * (H / 2) * 2);
    memset(&inBufs, 0, sizeof(inBufs));
    inBufs.numBufs = 1; inBufs.descs[0].buf = (XDAS_Int8*)frame; inBufs.descs[0].bufSize = W * H * 2;
    outPtr[0] = (XDAS_Int8*)res->preview; outSize[0] = (W / 2) * (H / 2) * 2;
    outBufs.bufs = outPtr; outBufs.numBufs = 1; outBufs.bufSizes = outSize;
    memset(&in, 0, sizeof(in));
    memset(&oa, 0, sizeof(oa));
    if (isMxn)
    {
      in.m.base.size = sizeof(in.m); in.m.base.numBytes = W * H * 2; in.m.base.inputID = f + 1;
      in.m.alg.widthM = 3; in.m.alg.heightN = 4;
      oa.m.base.size = sizeof(oa.m);
      res->ret = fx->process((IVIDTRANSCODE_Handle)h, &inBufs, &outBufs, &in.m.base, &oa.m.base);
      res->algSize = 12 * 4; memcpy(res->alg, &oa.m.alg, 48); res->base = oa.m.base;
    }
    else if (isObj)
    {
      in.o.base.size = sizeof(in.o); in.o.base.numBytes = W * H * 2; in.o.base.inputID = f + 1;
      in.o.alg.setHsvRange = (f == 0 || f == 2);       /* frames 1 and 3 run on the range the handle carries */
      in.o.alg.detectHue = (f == 2) ? 350 : 0; in.o.alg.detectHueTol = 40;      /* 350 +- 40 wraps around 0 */
      in.o.alg.detectSat = 60; in.o.alg.detectSatTol = 40; in.o.alg.detectVal = 60; in.o.alg.detectValTol = 40;
      oa.o.base.size = sizeof(oa.o);
      res->ret = fx->process((IVIDTRANSCODE_Handle)h, &inBufs, &outBufs, &in.o.base, &oa.o.base);
      res->algSize = 24; memcpy(res->alg, oa.o.alg.target, 24); res->base = oa.o.base;
    }
    else
    {
      in.r.base.size = sizeof(in.r); in.r.base.numBytes = W * H * 2; in.r.base.inputID = f + 1;
      in.r.alg.detectHueFrom = 0; in.r.alg.detectHueTo = 359; in.r.alg.detectSatFrom = 0; in.r.alg.detectSatTo = 100;
      in.r.alg.detectValFrom = 0; in.r.alg.detectValTo = 30;
      oa.t.base.size = sizeof(oa.t);
      res->ret = fx->process((IVIDTRANSCODE_Handle)h, &inBufs, &outBufs, &in.r.base, &oa.t.base);
      res->algSize = 3; memcpy(res->alg, &oa.t.alg, 3); res->base = oa.t.base;
    }
  }
  free(frame);
  n = fx->ialg.algFree(h, mem);
  /* memTab[1] (fast RAM) is leaked on purpose: the reference keeps LUT pointers into it in class statics */
  free(mem[0].base);
  return n == 2 ? 0 : 15;
}

int main(int argc, char** argv)
{
  enum { NF = 4 };
  Result a[NF], b[NF];
  void *la, *lb;
  IVIDTRANSCODE_Fxns *fa, *fb;
  int f, rc, bad = 0;
  if (argc < 4) { fprintf(stderr, "usage: %s <reference.so> <dropin.so> <wo|wl|ol|om|oo>\n", argv[0]); return 2; }
  la = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  lb = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
  if (!la || !lb) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  fa = (IVIDTRANSCODE_Fxns*)dlsym(la, "TRIK_VIDTRANSCODE_CV_FXNS");
  fb = (IVIDTRANSCODE_Fxns*)dlsym(lb, "TRIK_VIDTRANSCODE_CV_FXNS");
  if (!fa || !fb) { fprintf(stderr, "TRIK_VIDTRANSCODE_CV_FXNS not exported\n"); return 4; }
  memset(a, 0, sizeof(a)); memset(b, 0, sizeof(b));
  if ((rc = run(fa, argv[3], NF, a)) != 0) { fprintf(stderr, "reference run failed: %d\n", rc); return 5; }
  if ((rc = run(fb, argv[3], NF, b)) != 0) { fprintf(stderr, "drop-in run failed: %d\n", rc); return 6; }
  for (f = 0; f < NF; ++f)
  {
    const size_t psize = (size_t)(W / 2) * (H / 2) * 2;
    if (a[f].ret != b[f].ret || a[f].ret != 0) { fprintf(stderr, "frame %d: ret %d vs %d\n", f, a[f].ret, b[f].ret); ++bad; }
    if (memcmp(a[f].alg, b[f].alg, (size_t)a[f].algSize) != 0) { fprintf(stderr, "frame %d: OutArgsAlg differ\n", f); ++bad; }
    if (memcmp(a[f].preview, b[f].preview, psize) != 0) { fprintf(stderr, "frame %d: preview differs\n", f); ++bad; }
    if (a[f].base.bitsConsumed != b[f].base.bitsConsumed || a[f].base.bitsGenerated[0] != b[f].base.bitsGenerated[0]
        || a[f].base.outputID[0] != b[f].base.outputID[0] || a[f].base.decodedWidth != b[f].base.decodedWidth
        || a[f].base.encodedBuf[0].bufSize != b[f].base.encodedBuf[0].bufSize
        || a[f].base.extendedError != b[f].base.extendedError)
    { fprintf(stderr, "frame %d: xDM bookkeeping differs\n", f); ++bad; }
    printf("frame %d: ret %d  alg %02x %02x %02x  bitsGenerated %d\n", f, b[f].ret, b[f].alg[0], b[f].alg[1], b[f].alg[2],
           (int)b[f].base.bitsGenerated[0]);
  }
  printf(bad ? "MISMATCH (%d)\n" : "drop-in identical to the reference on %d frames\n", bad ? bad : NF);
  return bad ? 1 : 0;
}
