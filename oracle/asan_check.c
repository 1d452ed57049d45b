/* TEST INFRASTRUCTURE (oracle/): oracle hygiene (layer 5).  Built with
 * -fsanitize=address,undefined by `make -C oracle asan` and run by tests/test_oracle_hygiene.py:
 * every sensor of the C restatement over noise, flat and structured frames at several sizes, with and
 * without auto-calibration.  Any out-of-bounds access, signed overflow or bad shift aborts the run. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trik_oracle.h"

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void)
{
  s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17;
  return (uint32_t)(s_state >> 16);
}

static void fill(uint8_t* f, size_t n, int family, int w)
{
  size_t i;
  for (i = 0; i < n; ++i)
  {
    switch (family)
    {
      case 0: f[i] = (uint8_t)rnd(); break;                                   /* noise */
      case 1: f[i] = 0; break;
      case 2: f[i] = 255; break;
      case 3: f[i] = (uint8_t)((i % (size_t)w) * 255 / (size_t)w); break;     /* ramps */
      default: f[i] = (uint8_t)(((i / 64) & 1) ? 200 + (rnd() & 15) : 30 + (rnd() & 7)); break;
    }
  }
}

int main(void)
{
  static const int sizes[][2] = {{32, 4}, {96, 8}, {160, 120}, {320, 240}, {640, 480}, {1280, 720}};
  unsigned long checksum = 0;
  int si, kind, family, variant, runs = 0;
  for (si = 0; si < 6; ++si)
  {
    const int w = sizes[si][0], h = sizes[si][1];
    uint8_t* frame = (uint8_t*)malloc((size_t)w * h * 2);
    for (kind = TRIK_ORACLE_WO; kind <= TRIK_ORACLE_OM; ++kind)
    {
      const int line = (kind == TRIK_ORACLE_WO || kind == TRIK_ORACLE_WL) ? 2 * w : w;
      trik_oracle_sensor* s = trik_oracle_create(kind, w, h, line);
      if (!s) { fprintf(stderr, "create failed\n"); return 2; }
      for (family = 0; family < 5; ++family)
        for (variant = 0; variant < 3; ++variant)
        {
          const int autoDetect = (variant == 2 && w <= 640) ? 1 : 0;
          union { trik_oracle_range_in r; trik_oracle_obj_in o; trik_oracle_mxn_in m; } in;
          union { trik_oracle_target_out t; trik_oracle_obj_out o; trik_oracle_mxn_out m; } out;
          memset(&in, 0, sizeof(in));
          memset(&out, 0, sizeof(out));
          fill(frame, (size_t)w * h * 2, family, w);
          if (kind == TRIK_ORACLE_OO)
          {
            in.o.setHsvRange = (uint8_t)(variant != 1);
            in.o.detectHue = (uint16_t)(variant ? 350 : 120); in.o.detectHueTol = (uint16_t)(variant ? 359 : 25);
            in.o.detectSat = 60; in.o.detectSatTol = 100; in.o.detectVal = 55; in.o.detectValTol = 255;
            in.o.autoDetectHsv = (uint8_t)autoDetect;
          }
          else if (kind == TRIK_ORACLE_OM)
          {
            in.m.widthM = variant == 0 ? 3 : (variant == 1 ? 10 : 1);
            in.m.heightN = variant == 0 ? 3 : (variant == 1 ? 10 : 100);
            if (in.m.widthM > h || in.m.heightN > w) { in.m.widthM = 1; in.m.heightN = 1; }
          }
          else
          {
            in.r.detectHueFrom = (uint16_t)(variant ? 300 : 0); in.r.detectHueTo = (uint16_t)(variant ? 40 : 65535);
            in.r.detectSatFrom = 0; in.r.detectSatTo = 255; in.r.detectValFrom = (uint8_t)(variant ? 30 : 0);
            in.r.detectValTo = (uint8_t)(variant ? 255 : 40);
            in.r.autoDetectHsv = (uint8_t)autoDetect;
          }
          if (trik_oracle_run(s, frame, w * h * 2, &in, &out, 1234 + runs) != 1) { fprintf(stderr, "run failed\n"); return 3; }
          checksum = checksum * 31 + ((unsigned char*)&out)[0] + ((unsigned char*)&out)[2];
          ++runs;
        }
      trik_oracle_destroy(s);
    }
    free(frame);
  }
  {
    uint32_t i, acc = 0;
    for (i = 0; i < (1u << 24); i += 7)
      acc += trik_oracle_rgb888_to_hsv(trik_oracle_yuv_to_rgb888(i & 255, (i >> 8) & 255, (i >> 16) & 255));
    checksum += acc;
  }
  printf("oracle hygiene ok: %d runs, checksum %lu\n", runs, checksum);
  return 0;
}
