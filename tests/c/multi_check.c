/*
 * multi_check.c -- a plain C caller of the multi-GPU entry point of the C ABI (include/trik_b200.h,
 * trikb200_processBatchMulti): ONE process, one sensor handle per visible GPU (at least two handles, so the
 * splitting is exercised on a one-GPU box too), one batch of host frames, results in one array.  The results must
 * be identical to one handle taking the whole batch (== sequential process() calls).  Checked for the webcam line
 * sensor (no carried state) and the ov7670 line sensor (the cross band lags one frame,
 * ov7670/line_sensor/include/internal/cv_line_detector_seqpass.hpp:449-450).
 * argv[1] = libtrikb200.so.  Exit code 0 = identical.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trik_b200.h"

#define W 320
#define H 240
#define N 50

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

static void make_frame(unsigned char* f, int planar, int variant)
{
  uint32_t s = 777u + (uint32_t)variant;
  const int left = 20 + (variant * 5) % 240;
  int r, c;
  for (r = 0; r < H; ++r)
    for (c = 0; c < W; c += 2)
    {
      const int dark = (c > left && c < left + 40 + r / 8);
      const unsigned char y0 = (unsigned char)(dark ? 10 + lcg(&s) % 20 : 150 + lcg(&s) % 60);
      const unsigned char y1 = (unsigned char)(dark ? 10 + lcg(&s) % 20 : 150 + lcg(&s) % 60);
      const unsigned char u = 128, v = 128;
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

typedef IVIDTRANSCODE_Handle (*create_fn)(XDAS_Int32, XDAS_Int32, XDAS_Int32, XDAS_Int32, XDAS_Int32, XDAS_Int32);
typedef void (*delete_fn)(IVIDTRANSCODE_Handle);
typedef XDAS_Int32 (*batch_fn)(IVIDTRANSCODE_Handle, const TRIKB200_Batch*);
typedef XDAS_Int32 (*multi_fn)(const IVIDTRANSCODE_Handle*, XDAS_Int32, const TRIKB200_Batch*);
typedef XDAS_Int32 (*count_fn)(void);
typedef XDAS_Int32 (*setdev_fn)(XDAS_Int32);
typedef const char* (*err_fn)(void);

int main(int argc, char** argv)
{
  void* lib;
  create_fn create; delete_fn destroy; batch_fn processBatch; multi_fn processBatchMulti;
  count_fn deviceCount; setdev_fn setDevice; err_fn lastError;
  int kinds[2] = { TRIKB200_KIND_WL, TRIKB200_KIND_OL };
  int k, d, i, devices, handles;
  const size_t fbytes = (size_t)W * H * 2;
  unsigned char* frames = (unsigned char*)malloc(fbytes * N);
  if (argc < 2 || !frames) return 2;
  lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  create = (create_fn)dlsym(lib, "trikb200_create");
  destroy = (delete_fn)dlsym(lib, "trikb200_delete");
  processBatch = (batch_fn)dlsym(lib, "trikb200_processBatch");
  processBatchMulti = (multi_fn)dlsym(lib, "trikb200_processBatchMulti");
  deviceCount = (count_fn)dlsym(lib, "trikb200_deviceCount");
  setDevice = (setdev_fn)dlsym(lib, "trikb200_setDevice");
  lastError = (err_fn)dlsym(lib, "trikb200_lastError");
  if (!create || !destroy || !processBatch || !processBatchMulti || !deviceCount || !setDevice || !lastError) return 2;
  devices = deviceCount();
  if (devices < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
  handles = devices < 2 ? 2 : devices;
  for (k = 0; k < 2; ++k)
  {
    const int planar = kinds[k] == TRIKB200_KIND_OL;
    IVIDTRANSCODE_Handle one, many[64];
    TRIKB200_RangeInArgsAlg in;
    TRIKB200_TargetOutArgsAlg want[N], got[N];
    TRIKB200_Batch b;
    for (i = 0; i < N; ++i) make_frame(frames + fbytes * i, planar, i);
    memset(&in, 0, sizeof(in));
    in.detectHueFrom = 0; in.detectHueTo = 359; in.detectSatTo = 100; in.detectValTo = 40;
    memset(want, 0, sizeof(want)); memset(got, 0xEE, sizeof(got));
    memset(&b, 0, sizeof(b));
    b.size = (XDAS_Int32)sizeof(b);
    b.numFrames = N; b.frames = frames; b.frameStride = (int64_t)fbytes; b.framesMem = TRIKB200_MEM_HOST;
    b.inArgsAlg = &in; b.inArgsStride = 0;
    b.outArgsAlg = want; b.outArgsStride = (XDAS_Int32)sizeof(want[0]); b.outArgsMem = TRIKB200_MEM_HOST;
    setDevice(0);
    one = create(kinds[k], W, H, 0, 0, 0);
    if (!one || processBatch(one, &b) != IVIDTRANSCODE_EOK) { fprintf(stderr, "single: %s\n", lastError()); return 1; }
    for (d = 0; d < handles; ++d)
    {
      setDevice(d % devices);
      many[d] = create(kinds[k], W, H, 0, 0, 0);
      if (!many[d]) { fprintf(stderr, "create on device %d: %s\n", d % devices, lastError()); return 1; }
    }
    b.outArgsAlg = got;
    if (processBatchMulti(many, handles, &b) != IVIDTRANSCODE_EOK) { fprintf(stderr, "multi: %s\n", lastError()); return 1; }
    for (i = 0; i < N; ++i)
      if (want[i].targetX != got[i].targetX || want[i].targetY != got[i].targetY || want[i].targetSize != got[i].targetSize)
      {
        fprintf(stderr, "kind %d frame %d: (%d,%d,%d) vs (%d,%d,%d)\n", kinds[k], i, want[i].targetX, want[i].targetY,
                want[i].targetSize, got[i].targetX, got[i].targetY, got[i].targetSize);
        return 1;
      }
    printf("kind %d: %d frames over %d handles on %d device(s): frame 0 -> (%d,%d,%d), frame %d -> (%d,%d,%d)\n", kinds[k], N,
           handles, devices, got[0].targetX, got[0].targetY, got[0].targetSize, N - 1, got[N - 1].targetX, got[N - 1].targetY,
           got[N - 1].targetSize);
    for (d = 0; d < handles; ++d) destroy(many[d]);
    destroy(one);
  }
  printf("identical\n");
  free(frames);
  return 0;
}
