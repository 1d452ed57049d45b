/*
 * TEST INFRASTRUCTURE (oracle/): see trik_oracle.h.  Every function cites the reference
 * file:line it restates (paths relative to /root/reference/trik/).  Abbreviations:
 *   WO webcam/object_sensor   WL webcam/line_sensor
 *   OO ov7670/object_sensor   OL ov7670/line_sensor   OM ov7670/mxn_sensor
 *   inc = include/internal
 *
 * Third-party arithmetic on the path: libc srand/rand (glibc 2.39,
 * TYPE_3 additive-feedback generator) and libm pow in the annealed auto-calibration;
 * sqrtf/ceilf (IEEE, exact) in the size outputs; libstdc++ 13.3 std::sort (introsort,
 * threshold 16) in the OO cluster ranking.  rand/pow are CALLED here as the reference calls
 * them; std::sort is restated below (oracle_sort) because its tie order is observable.
 */
#include "trik_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * pixel functions
 * ---------------------------------------------------------------------------------------- */

static inline uint32_t sat8(int32_t v) { return v < 0 ? 0u : (v > 255 ? 255u : (uint32_t)v); }

/* WO/inc/cv_ball_detector_seqpass.hpp:181-205 (identical in all five sensors).  Closed form of
 * the packed 16-bit lane arithmetic: every sum wraps to int16 before the arithmetic >>6 and
 * the signed-16 -> unsigned-8 saturation of _spacku4.  Only the blue lane can actually wrap. */
uint32_t trik_oracle_yuv_to_rgb888(uint32_t y, uint32_t u, uint32_t v)
{
  const int16_t r16 = (int16_t)(uint16_t)(102 * (int32_t)v - 14248 + 74 * (int32_t)y);
  const int16_t g16 = (int16_t)(uint16_t)(-52 * (int32_t)v - 25 * (int32_t)u + 8696 + 74 * (int32_t)y);
  const int16_t b16 = (int16_t)(uint16_t)(129 * (int32_t)u - 17672 + 74 * (int32_t)y);
  return (sat8(r16 >> 6) << 16) | (sat8(g16 >> 6) << 8) | sat8(b16 >> 6);
}

/* WO/inc/cv_ball_detector_seqpass.hpp:207-249 with the LUTs of :400-406
 * (s_mult43_div[i] = 43*256/i, s_mult255_div[i] = 255*256/i, entry 0 = 0).  Result 0x00VVSSHH. */
uint32_t trik_oracle_rgb888_to_hsv(uint32_t rgb)
{
  const int32_t r = (rgb >> 16) & 0xff, g = (rgb >> 8) & 0xff, b = rgb & 0xff;
  const int32_t mx = r > g ? (r > b ? r : b) : (g > b ? g : b);
  const int32_t mn = r < g ? (r < b ? r : b) : (g < b ? g : b);
  const int32_t d = mx - mn;
  const uint32_t t255 = mx ? (uint32_t)(65280 / mx) : 0u;      /* uint16 table entry */
  const int32_t  t43  = d ? (11008 / d) : 0;
  const uint32_t sat_x256 = t255 * (uint32_t)d;
  int32_t hue_x256;
  const int g_is_max = (mx == g), b_is_max = (mx == b);        /* _cmpeq2 bit1, bit0 (:230) */
  if (!g_is_max && !b_is_max)
    hue_x256 = 0 + t43 * (g - b);                              /* :231-234 */
  else if (b_is_max && !g_is_max)
    hue_x256 = 43690 + t43 * (r - g);                          /* :235-238 */
  else
    hue_x256 = 21845 + t43 * (b - r);                          /* :239-242 */
  return ((uint32_t)mx << 16) | (((sat_x256 >> 8) & 0xffu) << 8) | (((uint32_t)hue_x256 >> 8) & 0xffu);
}

/* WO/inc/cv_ball_detector_seqpass.hpp:171-179: byte-wise unsigned compares give a 4-bit mask
 * (bit0 = H, bit1 = S, bit2 = V, bit3 = the always-zero top byte). */
int trik_oracle_detect(uint32_t hsv, uint32_t from, uint32_t to, uint32_t expected)
{
  uint32_t mask = 0;
  int i;
  for (i = 0; i < 4; ++i)
  {
    const uint32_t p = (hsv >> (8 * i)) & 0xff, f = (from >> (8 * i)) & 0xff, t = (to >> (8 * i)) & 0xff;
    if (p < f || p > t)
      mask |= 1u << i;
  }
  return mask == expected;
}

void trik_oracle_yuv_to_rgb888_range(uint32_t first, uint32_t count, uint32_t* out)
{
  uint32_t i;
  for (i = 0; i < count; ++i)
  {
    const uint32_t idx = first + i;                  /* Y | U << 8 | V << 16 */
    out[i] = trik_oracle_yuv_to_rgb888(idx & 0xff, (idx >> 8) & 0xff, (idx >> 16) & 0xff);
  }
}

void trik_oracle_rgb888_to_hsv_range(uint32_t first, uint32_t count, uint32_t* out)
{
  uint32_t i;
  for (i = 0; i < count; ++i)
    out[i] = trik_oracle_rgb888_to_hsv(first + i);
}

/* OM/inc/cv_ball_detector_seqpass.hpp:480-517.  float divisions promoted to double, as written. */
uint32_t trik_oracle_hsv_to_rgb_mxn(int H, int S, int V)
{
  double r = 0, g = 0, b = 0;
  double h = H / 255.0f;
  double s = S / 255.0f;
  double v = V / 255.0f;
  v = v < 0.2 ? 0 : v;
  s = s < 0.2 ? 0 : 1;
  {
    int i = h * 6;
    double f = h * 6 - i;
    double p = v * (1 - s);
    double q = v * (1 - f * s);
    double t = v * (1 - (1 - f) * s);
    switch (i % 6)
    {
      case 0: r = v; g = t; b = p; break;
      case 1: r = q; g = v; b = p; break;
      case 2: r = p; g = v; b = t; break;
      case 3: r = p; g = q; b = v; break;
      case 4: r = t; g = p; b = v; break;
      case 5: r = v; g = p; b = q; break;
    }
  }
  {
    int ri = r * 255, gi = g * 255, bi = b * 255;
    return (uint32_t)(((int32_t)ri << 16) + ((int32_t)gi << 8) + ((int32_t)bi));
  }
}

/* ------------------------------------------------------------------------------------------
 * sensor object
 * ---------------------------------------------------------------------------------------- */

typedef struct { int32_t x, y, size; } cluster_t;   /* OO/inc/cv_clusterizer_reference.hpp:22-26 */

struct trik_oracle_sensor {
  int kind, w, h, line;
  uint32_t* hsv;              /* w*h, 0x00VVSSHH (the reference's s_rgb888hsv low words) */
  /* OO: range persists while setHsvRange == 0 (cv_bitmap_builder_reference.hpp:110-130);
   * the object is value-initialised by "new T()" (src/vidtranscode_cv.cpp:58), hence zero. */
  uint32_t oo_from, oo_to, oo_expected;
  uint16_t* bitmap;           /* (w/4)*(h/4) */
  uint16_t* cmap;             /* (w/4)*(h/4) */
  uint16_t* equal;            /* label equivalence, one level (cv_clusterizer_reference.hpp:36) */
  cluster_t* clusters;
  int ncl;
  /* OL: m_hStart/m_hStop are assigned AFTER the pixel pass (OL/inc/cv_line_detector_seqpass.hpp:449-450),
   * so each frame uses the previous frame's values; zero on a fresh object. */
  uint32_t ol_hstart, ol_hstop;
  int flags;                  /* TRIK_ORACLE_FLAG_* of the last run */
};

static int is_planar(int kind) { return kind == TRIK_ORACLE_OO || kind == TRIK_ORACLE_OL || kind == TRIK_ORACLE_OM; }

/* CVAlgorithm::setup(): WO/inc/cv_ball_detector_seqpass.hpp:358-410 (size rules :365-369). */
trik_oracle_sensor* trik_oracle_create(int kind, int width, int height, int lineLength)
{
  trik_oracle_sensor* s;
  if (kind < TRIK_ORACLE_WO || kind > TRIK_ORACLE_OM)
    return NULL;
  if (width < 0 || height < 0 || width % 32 != 0 || height % 4 != 0)
    return NULL;
  s = (trik_oracle_sensor*)calloc(1, sizeof(*s));
  if (!s)
    return NULL;
  s->kind = kind; s->w = width; s->h = height; s->line = lineLength;
  s->hsv = (uint32_t*)calloc((size_t)width * height + 1, sizeof(uint32_t));
  if (kind == TRIK_ORACLE_OO)
  {
    const size_t cells = (size_t)(width / 4) * (height / 4) + 1;
    s->bitmap   = (uint16_t*)calloc(cells, sizeof(uint16_t));
    s->cmap     = (uint16_t*)calloc(cells, sizeof(uint16_t));
    s->equal    = (uint16_t*)calloc(cells + 1, sizeof(uint16_t));
    s->clusters = (cluster_t*)calloc(cells + 1, sizeof(cluster_t));
  }
  return s;
}

void trik_oracle_destroy(trik_oracle_sensor* s)
{
  if (!s) return;
  free(s->hsv); free(s->bitmap); free(s->cmap); free(s->equal); free(s->clusters);
  free(s);
}

const uint32_t* trik_oracle_last_hsv(const trik_oracle_sensor* s) { return s->hsv; }
int trik_oracle_last_flags(const trik_oracle_sensor* s) { return s->flags; }

/* pass 1: convertImageYuyvToHsv.
 * YUYV form   WO/inc/cv_ball_detector_seqpass.hpp:251-284 : bytes Y0 U Y1 V, row stride lineLength.
 * YUV422P form OO/inc/cv_ball_detector_seqpass.hpp:343-387 : luma plane at 0, chroma plane at
 *   lineLength*height; after _swap4 and the unpacks (:369-373) a pixel pair takes V = C[0], U = C[1]. */
static void convert_image(trik_oracle_sensor* s, const uint8_t* frame)
{
  int row, col;
  uint32_t* out = s->hsv;
  if (!is_planar(s->kind))
  {
    for (row = 0; row < s->h; ++row)
    {
      const uint8_t* p = frame + (size_t)row * s->line;
      for (col = 0; col < s->w; col += 2, p += 4)
      {
        *out++ = trik_oracle_rgb888_to_hsv(trik_oracle_yuv_to_rgb888(p[0], p[1], p[3]));
        *out++ = trik_oracle_rgb888_to_hsv(trik_oracle_yuv_to_rgb888(p[2], p[1], p[3]));
      }
    }
  }
  else
  {
    const uint8_t* chroma = frame + (size_t)s->line * s->h;
    for (row = 0; row < s->h; ++row)
    {
      const uint8_t* py = frame + (size_t)row * s->line;
      const uint8_t* pc = chroma + (size_t)row * s->line;
      for (col = 0; col < s->w; col += 2, py += 2, pc += 2)
      {
        *out++ = trik_oracle_rgb888_to_hsv(trik_oracle_yuv_to_rgb888(py[0], pc[1], pc[0]));
        *out++ = trik_oracle_rgb888_to_hsv(trik_oracle_yuv_to_rgb888(py[1], pc[1], pc[0]));
      }
    }
  }
}

static int32_t range_i32(int32_t lo, int32_t v, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* range packing: WO/inc/cv_ball_detector_seqpass.hpp:425-445 (also OO cv_bitmap_builder_reference.hpp:65-77) */
static void pack_range(uint32_t hf, uint32_t ht, uint32_t sf, uint32_t st, uint32_t vf, uint32_t vt,
                       uint32_t* from, uint32_t* to, uint32_t* expected)
{
  if (hf <= ht)
  {
    *from = (vf << 16) | (sf << 8) | hf;
    *to   = (vt << 16) | (st << 8) | ht;
    *expected = 0;
  }
  else
  {
    *from = (vf << 16) | (sf << 8) | (ht + 1);
    *to   = (vt << 16) | (st << 8) | (hf - 1);
    *expected = 1;
  }
}

/* ------------------------------------------------------------------------------------------
 * auto-calibration ("HsvRangeDetector::detect"), three families
 * ---------------------------------------------------------------------------------------- */

/* WO/inc/cv_hsv_range_detector.hpp:78-201: three 256-bin histograms over the open centre window,
 * mode = first bin to reach the final maximum; the uint16 ROI bounds are compared after integer
 * promotion (:153-156). */
static void detect_wo(const trik_oracle_sensor* s, trik_oracle_target_out* o)
{
  static uint32_t hh[256], hs[256], hv[256];
  const uint16_t hH = (uint16_t)(s->h / 2), hW = (uint16_t)(s->w / 2), step = (uint16_t)(s->h / 6);
  const uint16_t left = (uint16_t)(hW - step), right = (uint16_t)(hW + step);
  const uint16_t top = (uint16_t)(hH - step), bot = (uint16_t)(hH + step);
  uint32_t maxHue = 0, maxSat = 0, maxVal = 0;
  int32_t maxHueVal = 0, maxSatVal = 0, maxValVal = 0;
  const uint32_t* img = s->hsv;
  int row, col;
  memset(hh, 0, sizeof(hh)); memset(hs, 0, sizeof(hs)); memset(hv, 0, sizeof(hv));
  for (row = 0; row < (uint16_t)s->h; ++row)
    for (col = 0; col < (uint16_t)s->w; ++col)
    {
      const uint32_t p = *img++;
      const uint8_t hue = (uint8_t)p, sat = (uint8_t)(p >> 8), val = (uint8_t)(p >> 16);
      if (left < col && right > col && top < row && bot > row)
      {
        hh[hue]++; hs[sat]++; hv[val]++;
        if (hh[hue] > (uint32_t)maxHueVal) { maxHue = hue; maxHueVal = (int32_t)hh[hue]; }
        if (hs[sat] > (uint32_t)maxSatVal) { maxSat = sat; maxSatVal = (int32_t)hs[sat]; }
        if (hv[val] > (uint32_t)maxValVal) { maxVal = val; maxValVal = (int32_t)hv[val]; }
      }
    }
  o->detectHue = (uint16_t)((double)maxHue * 1.4f);
  o->detectHueTolerance = 15;
  o->detectSat = (uint16_t)((double)maxSat * 0.39f);
  o->detectSatTolerance = 30;
  o->detectVal = (uint16_t)((double)maxVal * 0.39f);
  o->detectValTolerance = 30;
}

/* --- WL / OL: WL/inc/cv_hsv_range_detector.hpp:81-305 (OL differs at :281 only) --- */

static int32_t s_vcl[256];

static int wl_do_get_increment(int val, int mn, int mx, double base, double t)   /* :95-112 */
{
  for (;;)
  {
    int res;
    double alpha, degree;
    if (mn == mx)
      return mn;
    alpha = rand() / (double)RAND_MAX;
    degree = 2 * alpha - 1;
    res = val + ((pow(base, degree) - 1) * t) * (double)(mx - mn);
    if (!((res < mn) || (res > mx)))
      return res;
  }
}

static int64_t wl_F(uint8_t v0, uint8_t v1)                                       /* :133-159 */
{
  int64_t res = 0;
  int v;
  for (v = v0; v <= v1; v++)
    res += s_vcl[v] != 0 ? s_vcl[v] : -1;
  return res;
}

static void detect_line(trik_oracle_sensor* s, trik_oracle_target_out* o, long long seed, int is_ol)
{
  const int step = 40;                                        /* WL/inc/cv_line_detector_seqpass.hpp:370,382 */
  const uint16_t hW = (uint16_t)(s->w / 2);
  const uint16_t left_p = (uint16_t)(hW - step), right_p = (uint16_t)(hW + step);
  const uint16_t left_n = (uint16_t)(left_p - step), right_n = (uint16_t)(right_p + step);
  const double T_end = 0.0005, lambda = 0.76, e = 2.718281828;
  const uint32_t* img = s->hsv;
  int32_t maxFillValue = 0;
  uint8_t maxFillV = 0; /* m_maxFillCluster is indeterminate in the reference if nothing is counted; tests avoid that */
  uint8_t Cv0, Cv1;
  int64_t L, newL;
  double T = 150;
  int row, col, i;

  srand((unsigned)seed);                                      /* :200, time() wrapped */
  memset(s_vcl, 0, sizeof(s_vcl));
  for (row = 0; row < (uint16_t)s->h; row++)
    for (col = 0; col < (uint16_t)s->w; col++)
    {
      const uint8_t v = (uint8_t)(*img++ >> 16);
      if (left_p < col && right_p > col)                      /* :225-234 */
      {
        s_vcl[v] += 1;
        if (s_vcl[v] > maxFillValue) { maxFillValue = s_vcl[v]; maxFillV = v; }
      }
      else if (left_n > col || right_n < col)                 /* :235-238 */
        s_vcl[v] -= 2;
    }

  if (maxFillValue == 0)
    s->flags |= TRIK_ORACLE_FLAG_LINE_SEED_INDETERMINATE;
  Cv0 = Cv1 = maxFillV;                                       /* :252-253 */
  L = wl_F(Cv0, Cv1);
  while (T > T_end)                                           /* :259-272 */
  {
    for (i = 0; i < 200; i++)
    {
      const double base = 1 + 1 / T;                          /* getIncrement :117-131, v0 first then v1 */
      const uint8_t n0 = (uint8_t)wl_do_get_increment(Cv0, 0, 255, base, T);
      const uint8_t n1 = (uint8_t)wl_do_get_increment(Cv1, 0, 255, base, T);
      newL = wl_F(n0, n1);
      if (rand() <= pow(e, (newL - L) / T) * RAND_MAX)
      {
        Cv0 = n0; Cv1 = n1; L = newL;
      }
    }
    T *= lambda;
  }
  Cv0 = (uint8_t)((Cv0 << 0) * 0.39f);                        /* :280 */
  if (is_ol)
    Cv1 = (uint8_t)((((Cv1 + 1) << 0)) * 0.39f);              /* OL :281 */
  else
    Cv1 = (uint8_t)((((Cv1 + 1) << 0) - 1) * 0.39f);          /* WL :281 */
  o->detectHue = 0; o->detectHueTolerance = 0; o->detectSat = 0; o->detectSatTolerance = 0;
  o->detectVal = (uint16_t)((Cv1 + Cv0) / 2);                 /* :302-303 */
  o->detectValTolerance = (uint16_t)((Cv1 - Cv0) / 2);
}

/* --- OO: OO/inc/cv_hsv_range_detector.hpp:34-294 --- */

static int32_t s_hs[32][32];

static int oo_get_increment(int val, int mn, int mx, double t)                    /* :77-98 */
{
  for (;;)
  {
    int res;
    double base, alpha, degree;
    if (mn == mx)
      return mn;
    base = 1 + 1 / t;
    alpha = rand() / (double)RAND_MAX;
    degree = 2 * alpha - 1;
    res = val + ((pow(base, degree) - 1) * t) * (double)(mx - mn);
    if ((mn <= res) && (res < mx))
      return res;
  }
}

static int oo_truncate_hue(int v) { int r = v % 32; if (r < 0) r += 32; return r; }  /* :101-107 */

static int64_t oo_foo(int h1, int h2, int s1, int s2)                             /* :109-153 */
{
  int64_t res = 0;
  int h, sp;
  if (h1 <= h2)
  {
    for (h = h1; h <= h2; h++)
      for (sp = s1; sp <= s2; sp++)
        res += s_hs[h][sp] != 0 ? s_hs[h][sp] : -2;
  }
  else
  {
    for (h = h1; h < 32; h++)
      for (sp = s1; sp <= s2; sp++)
        res += s_hs[h][sp] != 0 ? s_hs[h][sp] : -2;
    for (h = 0; h <= h2; h++)
      for (sp = s1; sp <= s2; sp++)
        res += s_hs[h][sp] != 0 ? s_hs[h][sp] : -2;
  }
  return res;
}

static void detect_oo(const trik_oracle_sensor* s, trik_oracle_obj_out* o, long long seed)
{
  const int width = s->w, height = s->h;
  const int hH = height / 2, hW = width / 2, step = height / 6;     /* :156-175, m_detectZoneScale = 6 */
  const int pos_l = hW - step, pos_r = hW + step, pos_t = hH - step, pos_b = hH + step;
  const int neg_l = hW - 2 * step, neg_r = hW + 2 * step, neg_t = hH - 2 * step, neg_b = hH + 2 * step;
  const double T_end = 0.0005, e = 2.718281828, lambda = 0.76;
  const uint32_t* img = s->hsv;
  int h1, h2, s1, s2, h_max = 0, s_max = 0, max_value = 0;
  int row, col, i;
  int64_t L;
  double T = 150;

  srand((unsigned)seed);                                      /* :180 */
  memset(s_hs, 0, sizeof(s_hs));
  for (row = 0; row < height; row++)                          /* :199-227 */
    for (col = 0; col < width; col++)
    {
      const uint32_t p = *img++;
      const int hp = (p & 0xff) >> 3, sp = ((p >> 8) & 0xff) >> 3;
      if (pos_l < col && col < pos_r && pos_t < row && row < pos_b)
      {
        s_hs[hp][sp] += 1;
        if (s_hs[hp][sp] > max_value) { max_value = s_hs[hp][sp]; h_max = hp; s_max = sp; }
      }
      else if (neg_r < col || col < neg_l || neg_b < row || row < neg_t)
        s_hs[hp][sp] -= 2;
    }

  h1 = h2 = h_max; s1 = s2 = s_max;
  L = oo_foo(h1, h2, s1, s2);
  while (T > T_end)                                           /* :246-269 */
  {
    for (i = 0; i < 200; i++)
    {
      const int h1n = oo_truncate_hue(oo_get_increment(h1, 0, 32, T));
      const int h2n = oo_truncate_hue(oo_get_increment(h2, 0, 32, T));
      const int s1n = oo_get_increment(s1, 0, s_max, T);
      const int s2n = oo_get_increment(s2, s_max, 32, T);
      const int64_t Ln = oo_foo(h1n, h2n, s1n, s2n);
      if (L < Ln || (rand() / (double)RAND_MAX) <= pow(e, -(L - Ln) / T))
      {
        h1 = h1n; h2 = h2n; s1 = s1n; s2 = s2n; L = Ln;
      }
    }
    T *= lambda;
  }
  h1 = (h1 << 3) * 1.4f;                                      /* :271-275 */
  h2 = (((h2 + 1) << 3) - 1) * 1.4f;
  s1 = (s1 << 3) * 0.39f;
  s2 = (((s2 + 1) << 3)) * 0.39f;
  if (h1 <= h2)
  {
    o->detectHue = (uint16_t)((h2 + h1) / 2);
    o->detectHueTolerance = (uint16_t)((h2 - h1) / 2);
  }
  else
  {
    float hue = (h2 - (360.0f - h1)) / 2;
    float hueTolerance = (h2 + (360.0f - h1)) / 2;
    o->detectHue = (uint16_t)(hue >= 0 ? hue : (hue + 360));
    o->detectHueTolerance = (uint16_t)hueTolerance;
  }
  o->detectSat = (uint16_t)((s2 + s1) / 2);
  o->detectSatTolerance = (uint16_t)((s2 - s1) / 2 + 2);
  o->detectVal = 50;
  o->detectValTolerance = 50;
}

/* ------------------------------------------------------------------------------------------
 * per-sensor pass 2 + tail
 * ---------------------------------------------------------------------------------------- */

/* WO/inc/cv_ball_detector_seqpass.hpp:412-508 */
static void run_wo(trik_oracle_sensor* s, const trik_oracle_range_in* in, trik_oracle_target_out* o)
{
  const uint32_t hf = (uint32_t)range_i32(0, ((int32_t)in->detectHueFrom * 255) / 359, 255);
  const uint32_t ht = (uint32_t)range_i32(0, ((int32_t)in->detectHueTo   * 255) / 359, 255);
  const uint32_t sf = (uint32_t)range_i32(0, ((int32_t)in->detectSatFrom * 255) / 100, 255);
  const uint32_t st = (uint32_t)range_i32(0, ((int32_t)in->detectSatTo   * 255) / 100, 255);
  const uint32_t vf = (uint32_t)range_i32(0, ((int32_t)in->detectValFrom * 255) / 100, 255);
  const uint32_t vt = (uint32_t)range_i32(0, ((int32_t)in->detectValTo   * 255) / 100, 255);
  uint32_t from, to, expected;
  int32_t tx = 0, ty = 0;
  uint32_t points = 0;
  pack_range(hf, ht, sf, st, vf, vt, &from, &to, &expected);

  if (s->h > 0 && s->w > 0)
  {
    const uint32_t* img = s->hsv;
    uint32_t row, col;
    if (in->autoDetectHsv)
      detect_wo(s, o);
    for (row = 0; row < (uint32_t)s->h; ++row)               /* proceedImageHsv :316-354 */
    {
      uint32_t perRow = 0, colSum = 0;
      for (col = 0; col < (uint32_t)s->w; ++col)
      {
        const int det = trik_oracle_detect(*img++, from, to, expected);
        perRow += det;
        colSum += det ? col : 0;
      }
      tx += colSum; ty += row * perRow; points += perRow;
    }
  }
  if (points > 0)                                            /* :486-499 */
  {
    const int32_t targetX = (uint32_t)tx / points;
    const int32_t targetY = (uint32_t)ty / points;
    const uint32_t radius = ceilf(sqrtf((float)points / 3.1415927f));
    o->targetX = ((targetX - (int32_t)s->w / 2) * 100 * 2) / (int32_t)s->w;
    o->targetY = ((targetY - (int32_t)s->h / 2) * 100 * 2) / (int32_t)s->h;
    o->targetSize = (uint32_t)(radius * 100 * 4) / (uint32_t)(s->w + s->h);
  }
  else
  {
    o->targetX = 0; o->targetY = 0; o->targetSize = 0;
  }
}

/* WL/inc/cv_line_detector_seqpass.hpp:325-420 and OL/inc/cv_line_detector_seqpass.hpp:372-476 */
static void run_line(trik_oracle_sensor* s, const trik_oracle_range_in* in, trik_oracle_target_out* o,
                     long long seed, int is_ol)
{
  const uint32_t vf = (uint32_t)range_i32(0, ((int32_t)in->detectValFrom * 255) / 100, 255);
  const uint32_t vt = (uint32_t)range_i32(0, ((int32_t)in->detectValTo   * 255) / 100, 255);
  uint32_t from, to, expected;
  int32_t tx = 0;
  uint32_t points = 0, cross = 0;
  pack_range(0, 255, 0, 255, vf, vt, &from, &to, &expected);  /* H and S arguments are ignored: WL :345-348 */

  if (s->h > 0 && s->w > 0)
  {
    const uint32_t* img = s->hsv;
    const uint32_t width = (uint32_t)s->w;
    uint32_t row, col;
    if (in->autoDetectHsv)
      detect_line(s, o, seed, is_ol);
    for (row = 0; row < (uint32_t)s->h; ++row)               /* WL :232-269, OL :258-301 */
    {
      uint32_t perRow = 0, colSum = 0;
      for (col = 0; col < width; ++col)
      {
        const uint32_t p = *img++;
        if (!is_ol || (col >= 5 && col <= width - 5))         /* OL :288 */
        {
          const int det = trik_oracle_detect(p, from, to, expected);
          perRow += det;
          colSum += det ? col : 0;
        }
      }
      tx += colSum; points += perRow;
      if (is_ol && row >= s->ol_hstart && row <= s->ol_hstop) /* OL :296-297 */
        cross += perRow;
    }
  }
  o->targetX = 0; o->targetY = 0; o->targetSize = 0;
  if (is_ol)
  {
    s->ol_hstart = (uint32_t)(s->h / 2);                      /* OL :449-450 */
    s->ol_hstop  = (uint32_t)(s->h / 2 + 2 * 40);
  }
  if (points > 10)                                           /* WL :401-417, OL :459-473 */
  {
    const int32_t inImagePixels = s->h * s->w;
    const int32_t targetX = (uint32_t)tx / points;
    o->targetX = ((targetX - (int32_t)s->w / 2) * 100 * 2) / (int32_t)s->w;
    if (is_ol)
      o->targetY = (int)((uint32_t)(cross * 100) / (uint32_t)(s->w * 2 * 40));   /* OL :452,471 */
    o->targetSize = (uint32_t)(points * 100 * 1) / (uint32_t)inImagePixels;
  }
}

/* --- OO --- */

static int mk_range(int v, int adj, int mn, int mx)           /* OO/inc/stdcpp.hpp:65-74 */
{
  v += adj;
  return v > mx ? mx : (v < mn ? mn : v);
}
static int mk_wrap(int v, int adj, int mn, int mx)            /* OO/inc/stdcpp.hpp:76-85 */
{
  v += adj;
  while (v > mx) v -= (mx - mn + 1);
  while (v < mn) v += (mx - mn + 1);
  return v;
}
static int16_t range_i16(int16_t lo, int16_t v, int16_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

static uint16_t pop16(uint16_t x)                             /* OO/inc/stdcpp.hpp:54-63 */
{
  x = x - ((x >> 1) & 0x5555);
  x = (x & 0x3333) + ((x >> 2) & 0x3333);
  x = (x + (x >> 4)) & 0x0f0f;
  x = x + (x >> 8);
  return x & 0x003f;
}

/* libstdc++ 13 std::sort with compareTargetBySize (a.size > b.size), restated:
 * bits/stl_algo.h __introsort_loop / __final_insertion_sort, _S_threshold = 16;
 * bits/stl_heap.h for the depth-limit fallback. */
static int cl_less(const cluster_t* a, const cluster_t* b) { return a->size > b->size; }
static void cl_swap(cluster_t* a, cluster_t* b) { cluster_t t = *a; *a = *b; *b = t; }

static void cl_push_heap(cluster_t* first, long hole, long top, cluster_t value)
{
  long parent = (hole - 1) / 2;
  while (hole > top && cl_less(first + parent, &value))
  {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void cl_adjust_heap(cluster_t* first, long hole, long len, cluster_t value)
{
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2)
  {
    second = 2 * (second + 1);
    if (cl_less(first + second, first + (second - 1)))
      second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2)
  {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  cl_push_heap(first, hole, top, value);
}
static void cl_heap_sort(cluster_t* first, cluster_t* last)
{
  const long len = last - first;
  long parent;
  if (len >= 2)
    for (parent = (len - 2) / 2;; parent--)
    {
      cluster_t value = first[parent];
      cl_adjust_heap(first, parent, len, value);
      if (parent == 0)
        break;
    }
  while (last - first > 1)
  {
    cluster_t value;
    --last;
    value = *last;
    *last = *first;
    cl_adjust_heap(first, 0, last - first, value);
  }
}
static void cl_median_to_first(cluster_t* result, cluster_t* a, cluster_t* b, cluster_t* c)
{
  if (cl_less(a, b))
  {
    if (cl_less(b, c)) cl_swap(result, b);
    else if (cl_less(a, c)) cl_swap(result, c);
    else cl_swap(result, a);
  }
  else if (cl_less(a, c)) cl_swap(result, a);
  else if (cl_less(b, c)) cl_swap(result, c);
  else cl_swap(result, b);
}
static cluster_t* cl_partition(cluster_t* first, cluster_t* last, cluster_t* pivot)
{
  for (;;)
  {
    while (cl_less(first, pivot)) ++first;
    --last;
    while (cl_less(pivot, last)) --last;
    if (!(first < last))
      return first;
    cl_swap(first, last);
    ++first;
  }
}
static void cl_introsort_loop(cluster_t* first, cluster_t* last, long depth)
{
  while (last - first > 16)
  {
    cluster_t *mid, *cut;
    if (depth == 0)
    {
      cl_heap_sort(first, last);
      return;
    }
    --depth;
    mid = first + (last - first) / 2;
    cl_median_to_first(first, first + 1, mid, last - 1);
    cut = cl_partition(first + 1, last, first);
    cl_introsort_loop(cut, last, depth);
    last = cut;
  }
}
static void cl_linear_insert(cluster_t* last)
{
  cluster_t val = *last;
  cluster_t* next = last - 1;
  while (cl_less(&val, next))
  {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}
static void cl_insertion_sort(cluster_t* first, cluster_t* last)
{
  cluster_t* i;
  if (first == last) return;
  for (i = first + 1; i != last; ++i)
  {
    if (cl_less(i, first))
    {
      cluster_t val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(cluster_t));
      *first = val;
    }
    else
      cl_linear_insert(i);
  }
}
static void oracle_sort(cluster_t* first, cluster_t* last)
{
  long n = last - first, lg = 0;
  cluster_t* i;
  if (first == last) return;
  while ((n >> (lg + 1)) != 0) lg++;                          /* std::__lg */
  cl_introsort_loop(first, last, lg * 2);
  if (last - first > 16)
  {
    cl_insertion_sort(first, first + 16);
    for (i = first + 16; i != last; ++i)
      cl_linear_insert(i);
  }
  else
    cl_insertion_sort(first, last);
}

/* OO/inc/cv_ball_detector_seqpass.hpp:514-598 with BitmapBuilder (cv_bitmap_builder_reference.hpp:107-217)
 * and Clusterizer (cv_clusterizer_reference.hpp:38-202). */
static void run_oo(trik_oracle_sensor* s, const trik_oracle_obj_in* in, trik_oracle_obj_out* o, long long seed)
{
  const int bw = s->w / 4, bh = s->h / 4;
  int i;
  memset(s->cmap, 0, (size_t)bw * bh * sizeof(uint16_t));
  memset(s->bitmap, 0, (size_t)bw * bh * sizeof(uint16_t));
  s->ncl = 0;

  if (s->h > 0 && s->w > 0)
  {
    int row, col;
    if (in->autoDetectHsv)
      detect_oo(s, o, seed);

    /* BitmapBuilder::run */
    if (in->setHsvRange)
    {
      const int32_t hueFrom = mk_wrap(in->detectHue, -in->detectHueTol, 0, 359);
      const int32_t hueTo   = mk_wrap(in->detectHue, +in->detectHueTol, 0, 359);
      const int32_t satFrom = mk_range(in->detectSat, -in->detectSatTol, 0, 100);
      const int32_t satTo   = mk_range(in->detectSat, +in->detectSatTol, 0, 100);
      const int32_t valFrom = mk_range(in->detectVal, -in->detectValTol, 0, 100);
      const int32_t valTo   = mk_range(in->detectVal, +in->detectValTol, 0, 100);
      const uint32_t hf = (uint32_t)range_i16(0, (int16_t)((hueFrom * 255) / 359), 255);
      const uint32_t ht = (uint32_t)range_i16(0, (int16_t)((hueTo   * 255) / 359), 255);
      const uint32_t sf = (uint32_t)range_i16(0, (int16_t)((satFrom * 255) / 100), 255);
      const uint32_t st = (uint32_t)range_i16(0, (int16_t)((satTo   * 255) / 100), 255);
      const uint32_t vf = (uint32_t)range_i16(0, (int16_t)((valFrom * 255) / 100), 255);
      const uint32_t vt = (uint32_t)range_i16(0, (int16_t)((valTo   * 255) / 100), 255);
      pack_range(hf, ht, sf, st, vf, vt, &s->oo_from, &s->oo_to, &s->oo_expected);
    }
    {
      const uint32_t* img = s->hsv;
      for (row = 0; row < s->h; row++)
      {
        /* s_hi2ho[row] = (row/4)*bitmapWidth as uint16 (:93-96) */
        uint16_t* out = s->bitmap + (uint16_t)((row / 4) * bw);
        const int shifter = (row % 4) * 4;
        int filler = 0;
        for (col = 0; col < s->w; col++)
        {
          const int det = trik_oracle_detect(*img++, s->oo_from, s->oo_to, s->oo_expected);
          *out = (uint16_t)(*out + (det << (shifter + filler++)));
          if (filler == 4) { out++; filler = 0; }
        }
      }
    }

    /* Clusterizer::run */
    s->equal[0] = 0;
    memset(&s->clusters[0], 0, sizeof(cluster_t));
    s->ncl = 1;
    {
      uint16_t maxCluster = 1;
      const uint16_t* src = s->bitmap;
      uint16_t* dst = s->cmap;
      for (row = 0; row < bh; row++)
        for (col = 0; col < bw; col++, dst++)
        {
          uint16_t a[4] = {0, 0, 0, 0}, v;
          int n;
          if (!(pop16(*src++) > 4 / 2))
            continue;
          if (row != 0)                                       /* setPixelEnvironment :69-83 */
          {
            a[2] = *(dst - bw);
            if (col != 0) a[1] = *(dst - bw - 1);
            if (col != bw - 1) a[3] = *(dst - bw + 1);
          }
          if (col != 0) a[0] = *(dst - 1);
          v = a[0];                                           /* min() :44-52 */
          for (n = 1; n < 4; n++)
            if ((a[n] < v && a[n] != 0) || v == 0)
              v = a[n];
          if (v)                                              /* setClusterNum :92-102 */
          {
            *dst = v;
            s->clusters[v].x += col;
            s->clusters[v].y += row;
            s->clusters[v].size++;
            for (n = 0; n < 4; n++)
              if (a[n])
                if (!(a[n] == v || s->equal[a[n]] == s->equal[v]))
                  s->equal[a[n]] = s->equal[v];
          }
          else                                                /* :104-111 */
          {
            *dst = maxCluster;
            s->equal[s->ncl] = maxCluster;
            memset(&s->clusters[s->ncl], 0, sizeof(cluster_t));
            s->ncl++;
            maxCluster++;
          }
        }
    }
    for (i = 0; i < s->ncl; i++)                              /* postProcessing :115-127 */
      if (i != s->equal[i])
      {
        cluster_t* d = &s->clusters[s->equal[i]];
        d->x += s->clusters[i].x;
        d->y += s->clusters[i].y;
        d->size += s->clusters[i].size;
        s->clusters[i].size = 0;
      }
    oracle_sort(s->clusters, s->clusters + s->ncl);
  }

  memset(o->target, 0, sizeof(o->target));                    /* OO/inc/cv_ball_detector_seqpass.hpp:569-597 */
  {
    int noObjects = 1;
    for (i = 0; i < 8; i++)
    {
      /* the reference reads clusters[i] even past the end of the vector (undefined behaviour);
       * the oracle defines those slots as empty and the test vectors keep >= 8 labels. */
      const cluster_t c = (i < s->ncl) ? s->clusters[i] : (cluster_t){0, 0, 0};
      if (i >= s->ncl)
        s->flags |= TRIK_ORACLE_FLAG_OO_FEWER_THAN_8;
      int size = sqrtf((float)(uint16_t)c.size);
      const uint32_t radius = ceilf(size / 3.1415927f);
      size = (uint32_t)(radius * 100 * 4) / (uint32_t)(bw + bh);
      if (size > 4)
      {
        const int x = (c.x / (c.size + 1)) * 4;
        const int y = (c.y / (c.size + 1)) * 4;
        noObjects = 0;
        o->target[i].size = size;
        o->target[i].x = ((x - (int32_t)s->w / 2) * 100 * 2) / (int32_t)s->w;
        o->target[i].y = ((y - (int32_t)s->h / 2) * 100 * 2) / (int32_t)s->h;
      }
    }
    if (noObjects)
    {
      o->target[0].x = 0; o->target[0].y = 0; o->target[0].size = 0;
    }
  }
}

/* OM/inc/cv_ball_detector_seqpass.hpp:560-621 with GetImgColor2 (:411-452) */
static void run_om(trik_oracle_sensor* s, const trik_oracle_mxn_in* in, trik_oracle_mxn_out* o)
{
  static int hist[32][4][4];
  const uint8_t heightM = (uint8_t)in->widthM;                /* names swapped, :585-586 */
  const uint8_t widthN  = (uint8_t)in->heightN;
  const uint16_t widthStep  = (uint16_t)(s->w / widthN);
  const uint16_t heightStep = (uint16_t)(s->h / heightM);
  int counter = 0, rowStart = 0, i, j;
  for (i = 0; i < heightM; ++i)
  {
    int colStart = 0;
    for (j = 0; j < widthN; ++j)
    {
      int ch_max = 0, cs_max = 0, cv_max = 0, maxEntry = 0, row, col;
      memset(hist, 0, sizeof(hist));
      for (row = 0; row < heightStep; row++)
        for (col = 0; col < widthStep; col++)
        {
          const uint32_t p = s->hsv[(size_t)(rowStart + row) * s->w + colStart + col];
          const int ch = (uint8_t)p / 8, cs = (uint8_t)(p >> 8) / 64, cv = (uint8_t)(p >> 16) / 64;
          hist[ch][cs][cv]++;
          if (hist[ch][cs][cv] > maxEntry)
          {
            maxEntry = hist[ch][cs][cv];
            ch_max = ch; cs_max = cs; cv_max = cv;
          }
        }
      o->outColor[counter++] = (int32_t)trik_oracle_hsv_to_rgb_mxn(ch_max * 8, cs_max * 64, cv_max * 64);
      colStart += widthStep;
    }
    rowStart += heightStep;
  }
}

int trik_oracle_run(trik_oracle_sensor* s, const uint8_t* frame, int numBytes,
                    const void* inArgsAlg, void* outArgsAlg, long long seed)
{
  s->flags = 0;
  if (s->h * s->line > numBytes)                              /* WO/inc/cv_ball_detector_seqpass.hpp:415-416 */
    return 0;
  if (s->h > 0 && s->w > 0)
    convert_image(s, frame);
  switch (s->kind)
  {
    case TRIK_ORACLE_WO: run_wo(s, (const trik_oracle_range_in*)inArgsAlg, (trik_oracle_target_out*)outArgsAlg); break;
    case TRIK_ORACLE_WL: run_line(s, (const trik_oracle_range_in*)inArgsAlg, (trik_oracle_target_out*)outArgsAlg, seed, 0); break;
    case TRIK_ORACLE_OL: run_line(s, (const trik_oracle_range_in*)inArgsAlg, (trik_oracle_target_out*)outArgsAlg, seed, 1); break;
    case TRIK_ORACLE_OO: run_oo(s, (const trik_oracle_obj_in*)inArgsAlg, (trik_oracle_obj_out*)outArgsAlg, seed); break;
    case TRIK_ORACLE_OM: run_om(s, (const trik_oracle_mxn_in*)inArgsAlg, (trik_oracle_mxn_out*)outArgsAlg); break;
    default: return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------
 * glibc 2.39 stdlib/random_r.c, TYPE_3 (x^31 + x^3 + 1), restated.  state34[0..30] = r[],
 * state34[31] = front index, state34[32] = rear index.
 * ---------------------------------------------------------------------------------------- */
void trik_oracle_srand(uint32_t* st, unsigned seed)
{
  int32_t word;
  int i;
  if (seed == 0) seed = 1;
  st[0] = seed;
  word = (int32_t)seed;
  for (i = 1; i < 31; ++i)
  {
    const long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    st[i] = (uint32_t)word;
  }
  st[31] = 3; st[32] = 0;
  for (i = 0; i < 310; ++i)
    (void)trik_oracle_rand(st);
}

int trik_oracle_rand(uint32_t* st)
{
  uint32_t f = st[31], r = st[32];
  const uint32_t val = st[f] += st[r];
  if (++f >= 31) f = 0;
  if (++r >= 31) r = 0;
  st[31] = f; st[32] = r;
  return (int)(val >> 1);
}

/* =============================================================================================
 * ov7670/edge_line_sensor
 *   convertImageYuyvToRgb: include/internal/cv_ball_detector_seqpass.hpp:151-205 (Sobel, threshold, counting loop)
 *   run tail:              :386-414
 * The Sobel and threshold kernels are TI IMGLIB (closed, absent): restated in oracle/imglib_open.c, PARITY UNPINNED.
 * The reference hands the input to IMG_sobel_3x3_8 as a dense width x height array (it ignores inputLineLength there);
 * here rows are lineLength bytes apart and are gathered densely first, which is the same thing when lineLength == width.
 * The reference's work buffer s_y is a zero-initialised static whose last two rows (and first byte) the Sobel never
 * writes, so they stay 0 for ever: calloc reproduces that.
 * ============================================================================================= */
void IMG_sobel_3x3_8(const unsigned char* in, unsigned char* out, short cols, short rows);
void IMG_thr_gt2max_8(const unsigned char* in_data, unsigned char* out_data, short cols, short rows, unsigned char threshold);

void trik_oracle_edge_line(const uint8_t* frame, int width, int height, int lineLength, void* outArgsAlg)
{
  trik_oracle_target_out* o = (trik_oracle_target_out*)outArgsAlg;
  const size_t n = (size_t)width * (size_t)height;
  uint8_t* in = (uint8_t*)malloc(n ? n : 1);
  uint8_t* sy = (uint8_t*)calloc(n ? n : 1, 1);
  int32_t tx = 0;
  uint32_t points = 0;
  int r, c;
  o->targetX = 0; o->targetY = 0; o->targetSize = 0;
  if (!in || !sy || width <= 0 || height <= 0) { free(in); free(sy); return; }
  for (r = 0; r < height; ++r)
    memcpy(in + (size_t)r * width, frame + (size_t)r * lineLength, (size_t)width);
  IMG_sobel_3x3_8(in, sy, (short)width, (short)height);                       /* :176-178 */
  IMG_thr_gt2max_8(sy, sy, (short)width, (short)height, 50);                  /* :180-182 */
  for (r = 0; r < height; ++r)                                                /* :186-205 */
  {
    uint16_t perRow = 0, colSum = 0;                                          /* uint16_t in the reference: sums wrap */
    for (c = 0; c < width; ++c)
      if (c > 15 && c < width - 15)
      {
        const int det = sy[(size_t)r * width + c] == 0xFF;
        perRow = (uint16_t)(perRow + det);
        colSum = (uint16_t)(colSum + (det ? c : 0));
      }
    tx += colSum;
    points += perRow;
  }
  if (points > 0)                                                             /* :388-403; m_targetY is never added to */
  {
    const int32_t targetX = (int32_t)((uint32_t)tx / points);
    const int32_t targetY = 0;
    const uint32_t radius = (uint32_t)ceilf(sqrtf((float)points / 3.1415927f));
    o->targetX = (int8_t)(((targetX - (int32_t)width / 2) * 100 * 2) / (int32_t)width);
    o->targetY = (int8_t)(((targetY - (int32_t)height / 2) * 100 * 2) / (int32_t)height);
    o->targetSize = (uint8_t)((uint32_t)(radius * 100 * 4) / (uint32_t)(width + height));
  }
  free(in);
  free(sy);
}
