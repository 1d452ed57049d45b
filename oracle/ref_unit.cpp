/*
 * TEST INFRASTRUCTURE (oracle/): the translation unit that oracle/build_ref.sh compiles for each
 * sensor.  It #includes the reference's own, unmodified src/vidtranscode_cv.cpp (path given by
 * -DTRIKREF_SRC) and adds two probes that call the reference's PRIVATE static pixel functions
 *   convert2xYuyvToRgb888  (<sensor>/include/internal/cv_*_seqpass.hpp, e.g. WO :181-205)
 *   convertRgb888ToHsv     (e.g. WO :207-249)
 * directly, so that the oracle's closed forms can be checked against the intrinsic code on all
 * 2^24 inputs.  The probes must live in the SAME translation unit as the reference (its headers
 * define non-inline static data members), and "private" is opened with the usual test-only macro
 * after every standard header the reference uses has been included.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <memory>
#include <set>
#include <vector>

#include <c6x.h>

#define private public
#define protected public
#include TRIKREF_SRC
#undef private
#undef protected

typedef trik::cv::TRIKREF_CLASS<TRIKREF_FORMAT, TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X> TrikRefAlgorithm;

/* out[i] = 0x00RRGGBB of the FIRST pixel of the YUYV word built from index first+i = Y | U<<8 | V<<16;
 * the second pixel of the word carries luma 255-Y and is checked to land in the high word. */
#ifdef TRIKREF_NO_PIXEL_PROBES   /* edge_line_sensor: no HSV pipeline in this class */
extern "C" int trikref_probe_yuv2rgb(uint32_t, uint32_t, uint32_t*, uint32_t*) { return -1; }
extern "C" int trikref_probe_rgb2hsv(uint32_t, uint32_t, uint32_t*) { return -1; }
#else
extern "C" int trikref_probe_yuv2rgb(uint32_t first, uint32_t count, uint32_t* out, uint32_t* outSecond)
{
  for (uint32_t i = 0; i < count; ++i)
  {
    const uint32_t idx = first + i;
    const uint32_t y = idx & 0xffu, u = (idx >> 8) & 0xffu, v = (idx >> 16) & 0xffu;
    const uint32_t word = y | (u << 8) | ((255u - y) << 16) | (v << 24);
    const uint64_t rgb2 = TrikRefAlgorithm::convert2xYuyvToRgb888(word);
    out[i] = _loll(rgb2);
    if (outSecond)
      outSecond[i] = _hill(rgb2);
  }
  return 0;
}

/* out[i] = 0x00VVSSHH of 0x00RRGGBB = first+i.  Needs one codec instance to have been created
 * (the division LUTs are class statics filled by setup()). */
extern "C" int trikref_probe_rgb2hsv(uint32_t first, uint32_t count, uint32_t* out)
{
  if (TrikRefAlgorithm::s_mult43_div == NULL || TrikRefAlgorithm::s_mult255_div == NULL)
    return -1;
  for (uint32_t i = 0; i < count; ++i)
    out[i] = TrikRefAlgorithm::convertRgb888ToHsv(first + i);
  return 0;
}
#endif
