// trik_host.cpp -- see trik_host.hpp.
#include "trik_host.hpp"

#include <cstring>

namespace trikb200 {

namespace {

inline int32_t clampi(int32_t lo, int32_t v, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// V-only threshold of the line sensors expressed on the 16-bit keys of trik_pixel.cuh:
//   V = sat8(m >> 6), m = key - 0x8000 (int16).  V >= vf <=> vf == 0 or m >= vf*64;
//   V <= vt <=> vt == 255 or m <= vt*64 + 63.
void v_range_to_keys(uint32_t vf, uint32_t vt, FrameParams& fp)
{
  uint32_t klo, n;
  if (vf > vt)
  {
    // empty range: key 0 is unreachable (max(R,G) >= -14248 -> key >= 18520), so "key == 0" never passes
    klo = 0; n = 0;
  }
  else
  {
    klo = (vf == 0u) ? 0u : 0x8000u + vf * 64u;
    const uint32_t khi = (vt == 255u) ? 0xFFFFu : 0x8000u + vt * 64u + 63u;
    n = khi - klo;
  }
  const uint32_t negKlo = (0x10000u - klo) & 0xFFFFu;
  fp.negKlo2 = negKlo * 0x10001u;
  fp.n2 = n * 0x10001u;
}

int make_value_range(int v, int adj, int mn, int mx)      // ov7670/object_sensor/include/internal/stdcpp.hpp:65-74
{
  v += adj;
  return v > mx ? mx : (v < mn ? mn : v);
}
int make_value_wrap(int v, int adj, int mn, int mx)       // stdcpp.hpp:76-85
{
  v += adj;
  while (v > mx) v -= (mx - mn + 1);
  while (v < mn) v += (mx - mn + 1);
  return v;
}

} // namespace

void pack_hsv_range(uint32_t hf, uint32_t ht, uint32_t sf, uint32_t st, uint32_t vf, uint32_t vt,
                    uint32_t& from, uint32_t& to, uint32_t& expected)
{
  if (hf <= ht)
  {
    from = (vf << 16) | (sf << 8) | hf;
    to   = (vt << 16) | (st << 8) | ht;
    expected = 0u;
  }
  else
  {
    from = (vf << 16) | (sf << 8) | (ht + 1u);
    to   = (vt << 16) | (st << 8) | (hf - 1u);
    expected = 1u;
  }
}

void prepare_frame_params(int kind, const Geometry& g, const void* inArgsAlg, CarriedState& st, FrameParams& fp)
{
  std::memset(&fp, 0, sizeof(fp));
  switch (kind)
  {
    case KIND_WO:
    {
      // webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:425-445
      const TRIKB200_RangeInArgsAlg* a = static_cast<const TRIKB200_RangeInArgsAlg*>(inArgsAlg);
      const uint32_t hf = (uint32_t)clampi(0, ((int32_t)a->detectHueFrom * 255) / 359, 255);
      const uint32_t ht = (uint32_t)clampi(0, ((int32_t)a->detectHueTo   * 255) / 359, 255);
      const uint32_t sf = (uint32_t)clampi(0, ((int32_t)a->detectSatFrom * 255) / 100, 255);
      const uint32_t sT = (uint32_t)clampi(0, ((int32_t)a->detectSatTo   * 255) / 100, 255);
      const uint32_t vf = (uint32_t)clampi(0, ((int32_t)a->detectValFrom * 255) / 100, 255);
      const uint32_t vt = (uint32_t)clampi(0, ((int32_t)a->detectValTo   * 255) / 100, 255);
      pack_hsv_range(hf, ht, sf, sT, vf, vt, fp.from, fp.to, fp.expected);
      if (a->autoDetectHsv) fp.flags |= FP_AUTODETECT;
      break;
    }
    case KIND_WL:
    case KIND_OL:
    {
      // webcam/line_sensor/include/internal/cv_line_detector_seqpass.hpp:345-366: H and S arguments are ignored
      const TRIKB200_RangeInArgsAlg* a = static_cast<const TRIKB200_RangeInArgsAlg*>(inArgsAlg);
      const uint32_t vf = (uint32_t)clampi(0, ((int32_t)a->detectValFrom * 255) / 100, 255);
      const uint32_t vt = (uint32_t)clampi(0, ((int32_t)a->detectValTo   * 255) / 100, 255);
      v_range_to_keys(vf, vt, fp);
      pack_hsv_range(0, 255, 0, 255, vf, vt, fp.from, fp.to, fp.expected);
      if (a->autoDetectHsv) fp.flags |= FP_AUTODETECT;
      if (kind == KIND_OL)
      {
        fp.hStart = st.olHStart;           // this frame is judged with the PREVIOUS frame's band
        fp.hStop  = st.olHStop;
        st.olHStart = (uint32_t)(g.height / 2);                 // ov7670/line_sensor/.../cv_line_detector_seqpass.hpp:449-450
        st.olHStop  = (uint32_t)(g.height / 2 + 2 * 40);
      }
      break;
    }
    case KIND_OO:
    {
      // ov7670/object_sensor/include/internal/cv_bitmap_builder_reference.hpp:110-130
      const TRIKB200_ObjInArgsAlg* a = static_cast<const TRIKB200_ObjInArgsAlg*>(inArgsAlg);
      if (a->setHsvRange)
      {
        const int32_t hueFrom = make_value_wrap(a->detectHue, -(int)a->detectHueTol, 0, 359);
        const int32_t hueTo   = make_value_wrap(a->detectHue, +(int)a->detectHueTol, 0, 359);
        const int32_t satFrom = make_value_range(a->detectSat, -(int)a->detectSatTol, 0, 100);
        const int32_t satTo   = make_value_range(a->detectSat, +(int)a->detectSatTol, 0, 100);
        const int32_t valFrom = make_value_range(a->detectVal, -(int)a->detectValTol, 0, 100);
        const int32_t valTo   = make_value_range(a->detectVal, +(int)a->detectValTol, 0, 100);
        const uint32_t hf = (uint32_t)clampi(0, (int16_t)((hueFrom * 255) / 359), 255);
        const uint32_t ht = (uint32_t)clampi(0, (int16_t)((hueTo   * 255) / 359), 255);
        const uint32_t sf = (uint32_t)clampi(0, (int16_t)((satFrom * 255) / 100), 255);
        const uint32_t sT = (uint32_t)clampi(0, (int16_t)((satTo   * 255) / 100), 255);
        const uint32_t vf = (uint32_t)clampi(0, (int16_t)((valFrom * 255) / 100), 255);
        const uint32_t vt = (uint32_t)clampi(0, (int16_t)((valTo   * 255) / 100), 255);
        pack_hsv_range(hf, ht, sf, sT, vf, vt, st.ooFrom, st.ooTo, st.ooExpected);
      }
      fp.from = st.ooFrom; fp.to = st.ooTo; fp.expected = st.ooExpected;
      if (a->autoDetectHsv) fp.flags |= FP_AUTODETECT;
      break;
    }
    case KIND_OM:
    {
      // ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp:585-588 (names swapped there)
      const TRIKB200_MxnInArgsAlg* a = static_cast<const TRIKB200_MxnInArgsAlg*>(inArgsAlg);
      fp.gridRows = (uint8_t)a->widthM;
      fp.gridCols = (uint8_t)a->heightN;
      break;
    }
    default:
      break;
  }
}

// glibc 2.39 stdlib/random_r.c, TYPE_3: degree 31, separation 3.
void GlibcRand::seed(unsigned s)
{
  if (s == 0) s = 1;
  r[0] = (int32_t)s;
  int32_t word = (int32_t)s;
  for (int i = 1; i < 31; ++i)
  {
    const long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    r[i] = word;
  }
  f = 3; b = 0;
  for (int i = 0; i < 310; ++i)
    (void)next();
}

int GlibcRand::next()
{
  const uint32_t val = (uint32_t)r[f] + (uint32_t)r[b];
  r[f] = (int32_t)val;
  if (++f >= 31) f = 0;
  if (++b >= 31) b = 0;
  return (int)(val >> 1);
}

} // namespace trikb200
