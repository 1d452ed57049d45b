// trik_host.cpp -- see trik_host.hpp.
#include "trik_host.hpp"

#include <cmath>
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
    // keep N + 1 representable in a lane: key 65535 (value 32767) is unreachable -- 129*U + 74*Y = 50439 has
    // no solution in bytes -- so an all-inclusive range loses nothing by stopping at 65534
    if (n == 0xFFFFu)
      n = 0xFFFEu;
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
      // callers validate 1 <= widthM, heightN and widthM * heightN <= 100 on the int32 values first (trik_capi.cu)
      fp.gridRows = (uint32_t)a->widthM;
      fp.gridCols = (uint32_t)a->heightN;
      break;
    }
    default:
      break;
  }
}

// float divisions widened to double, v < 0.2 -> 0, s < 0.2 -> 0 else 1, six-sector conversion,
// truncation to int -- the arithmetic of HSVtoRGB() as written; plain IEEE, no libm.
static uint32_t mxn_hsv_to_rgb(int H, int S, int V)
{
  double r = 0, g = 0, b = 0;
  const double h = H / 255.0f;
  double s = S / 255.0f;
  double v = V / 255.0f;
  v = v < 0.2 ? 0 : v;
  s = s < 0.2 ? 0 : 1;
  const int i = (int)(h * 6);
  const double f = h * 6 - i;
  const double p = v * (1 - s);
  const double q = v * (1 - f * s);
  const double t = v * (1 - (1 - f) * s);
  switch (i % 6)
  {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    case 5: r = v; g = p; b = q; break;
  }
  const int ri = (int)(r * 255), gi = (int)(g * 255), bi = (int)(b * 255);
  return (uint32_t)((ri << 16) + (gi << 8) + bi);
}

void mxn_color_table(uint32_t table[512])
{
  for (int h = 0; h < 32; ++h)
    for (int s = 0; s < 4; ++s)
      for (int v = 0; v < 4; ++v)
        table[(h << 4) | (s << 2) | v] = mxn_hsv_to_rgb(h * 8, s * 64, v * 64);
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

// ---------------------------------------------------------------------------------------------
// annealing tails
// ---------------------------------------------------------------------------------------------
namespace {

const double kTEnd = 0.0005, kLambda = 0.76, kE = 2.718281828;
const double kRandMax = 2147483647.0;                      // glibc RAND_MAX

// do_getIncrement (WL/.../cv_hsv_range_detector.hpp:95-112): rejection sampling, inclusive bounds
int line_increment(GlibcRand& rng, int val, int mn, int mx, double base, double t)
{
  for (;;)
  {
    if (mn == mx)
      return mn;
    const double alpha = rng.next() / kRandMax;
    const double degree = 2 * alpha - 1;
    const int res = (int)(val + ((std::pow(base, degree) - 1) * t) * (double)(mx - mn));
    if (!((res < mn) || (res > mx)))
      return res;
  }
}

// F (:133-159): sum over v0..v1 of (bin != 0 ? bin : -K0), K0 = 1.  The 9200 evaluations per frame only
// differ in their bounds, so they are answered from an exact int64 prefix sum instead of a 256-step loop.
struct LinePrefix {
  int64_t p[257];
  explicit LinePrefix(const int32_t* h)
  {
    p[0] = 0;
    for (int v = 0; v < 256; ++v)
      p[v + 1] = p[v] + (h[v] != 0 ? h[v] : -1);
  }
  int64_t F(uint8_t v0, uint8_t v1) const { return v0 <= v1 ? p[v1 + 1] - p[v0] : 0; }
};

// getIncrement (OO/.../cv_hsv_range_detector.hpp:77-98): half-open upper bound
int oo_increment(GlibcRand& rng, int val, int mn, int mx, double t)
{
  for (;;)
  {
    if (mn == mx)
      return mn;
    const double base = 1 + 1 / t;
    const double alpha = rng.next() / kRandMax;
    const double degree = 2 * alpha - 1;
    const int res = (int)(val + ((std::pow(base, degree) - 1) * t) * (double)(mx - mn));
    if ((mn <= res) && (res < mx))
      return res;
  }
}

int oo_truncate_hue(int v) { int r = v % 32; if (r < 0) r += 32; return r; }

// m_foo (:109-153): sum over the hue x saturation rectangle (hue wraps when h1 > h2) of
// (cell != 0 ? cell : -K0), K0 = 2; answered from an exact int64 summed-area table.
struct HsPrefix {
  int64_t sat[33][33];
  explicit HsPrefix(const int32_t* hs)
  {
    for (int i = 0; i <= 32; ++i) sat[i][0] = sat[0][i] = 0;
    for (int h = 0; h < 32; ++h)
      for (int s = 0; s < 32; ++s)
        sat[h + 1][s + 1] = sat[h][s + 1] + sat[h + 1][s] - sat[h][s] + (hs[h * 32 + s] != 0 ? hs[h * 32 + s] : -2);
  }
  int64_t rect(int ha, int hb, int s1, int s2) const          // hue rows ha..hb, sat columns s1..s2, inclusive
  {
    if (ha > hb || s1 > s2) return 0;
    return sat[hb + 1][s2 + 1] - sat[ha][s2 + 1] - sat[hb + 1][s1] + sat[ha][s1];
  }
  int64_t foo(int h1, int h2, int s1, int s2) const
  {
    return h1 <= h2 ? rect(h1, h2, s1, s2) : rect(h1, 31, s1, s2) + rect(0, h2, s1, s2);
  }
};

} // namespace

void anneal_line(const int32_t hist[256], int seedBin, bool isOL, unsigned seed, uint16_t out[6])
{
  GlibcRand rng;
  rng.seed(seed);
  uint8_t v0 = (uint8_t)seedBin, v1 = (uint8_t)seedBin;
  const LinePrefix pre(hist);
  int64_t L = pre.F(v0, v1);
  double T = 150;
  while (T > kTEnd)
  {
    for (int i = 0; i < 200; i++)
    {
      const double base = 1 + 1 / T;
      const uint8_t n0 = (uint8_t)line_increment(rng, v0, 0, 255, base, T);
      const uint8_t n1 = (uint8_t)line_increment(rng, v1, 0, 255, base, T);
      const int64_t newL = pre.F(n0, n1);
      if (rng.next() <= std::pow(kE, (newL - L) / T) * kRandMax)
      {
        v0 = n0; v1 = n1; L = newL;
      }
    }
    T *= kLambda;
  }
  v0 = (uint8_t)((v0 << 0) * 0.39f);
  v1 = isOL ? (uint8_t)((((v1 + 1) << 0)) * 0.39f) : (uint8_t)((((v1 + 1) << 0) - 1) * 0.39f);
  out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 0;
  out[4] = (uint16_t)((v1 + v0) / 2);
  out[5] = (uint16_t)((v1 - v0) / 2);
}

void anneal_oo(const int32_t hs[1024], int seedBin, unsigned seed, uint16_t out[6])
{
  GlibcRand rng;
  rng.seed(seed);
  const int sMax = seedBin & 31;
  int h1 = seedBin >> 5, h2 = seedBin >> 5, s1 = sMax, s2 = sMax;
  const HsPrefix pre(hs);
  int64_t L = pre.foo(h1, h2, s1, s2);
  double T = 150;
  while (T > kTEnd)
  {
    for (int i = 0; i < 200; i++)
    {
      const int h1n = oo_truncate_hue(oo_increment(rng, h1, 0, 32, T));
      const int h2n = oo_truncate_hue(oo_increment(rng, h2, 0, 32, T));
      const int s1n = oo_increment(rng, s1, 0, sMax, T);
      const int s2n = oo_increment(rng, s2, sMax, 32, T);
      const int64_t Ln = pre.foo(h1n, h2n, s1n, s2n);
      if (L < Ln || (rng.next() / kRandMax) <= std::pow(kE, -(L - Ln) / T))
      {
        h1 = h1n; h2 = h2n; s1 = s1n; s2 = s2n; L = Ln;
      }
    }
    T *= kLambda;
  }
  h1 = (int)((h1 << 3) * 1.4f);
  h2 = (int)((((h2 + 1) << 3) - 1) * 1.4f);
  s1 = (int)((s1 << 3) * 0.39f);
  s2 = (int)((((s2 + 1) << 3)) * 0.39f);
  if (h1 <= h2)
  {
    out[0] = (uint16_t)((h2 + h1) / 2);
    out[1] = (uint16_t)((h2 - h1) / 2);
  }
  else
  {
    const float hue = (h2 - (360.0f - h1)) / 2;
    const float hueTolerance = (h2 + (360.0f - h1)) / 2;
    out[0] = (uint16_t)(hue >= 0 ? hue : (hue + 360));
    out[1] = (uint16_t)hueTolerance;
  }
  out[2] = (uint16_t)((s2 + s1) / 2);
  out[3] = (uint16_t)((s2 - s1) / 2 + 2);
  out[4] = 50;
  out[5] = 50;
}

} // namespace trikb200
