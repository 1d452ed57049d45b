// trik_kernels_detect.cu -- the histogram halves of the auto-calibration ("HsvRangeDetector::detect").
//
//   WO  webcam/object_sensor/include/internal/cv_hsv_range_detector.hpp:116-201
//       three 256-bin histograms over the open centre window, mode = first bin to reach the final
//       maximum.  Deterministic and complete on the device.
//   WL / OL  webcam/line_sensor/include/internal/cv_hsv_range_detector.hpp:195-240
//       256-bin V histogram: +1 inside the centre band, -2 outside the double band, and the bin
//       that first reaches the running maximum seeds the annealing.
//   OO  ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:177-227
//       the same on a 32x32 (H>>3, S>>3) grid with a centre window / double window.
//
// The +1/-2 histograms can fall, so "first bin to exceed the running maximum" depends on raster
// order.  Within one row the order is: left negatives, positives, right negatives.  A bin's value
// at its k-th positive pixel of row r is  base_b(r) - 2*NL_b(r) + k, which peaks at the bin's LAST
// positive pixel of the row.  So per row it is enough to know, per bin, NL, P, NR and the column
// of the last positive pixel: the row maximum M_r = max_b peak_b, and if M_r beats the running
// maximum the new seed is the bin with peak_b == M_r whose last positive column is smallest.
// Rows are replayed in order by one CTA per frame; everything inside a row is parallel.
//
// The annealing that follows (9200 moves driven by rand()/pow()) runs on the host: see trik_host.cpp.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_detect;
std::atomic<long long> g_launches_detect{0};

// ---------------------------------------------------------------------------------------------
// WO
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wo_detect_kernel(const Geometry g, const uint8_t* __restrict__ frames, const int* __restrict__ frameIdx,
                 TargetOut* __restrict__ out, const int left, const int right, const int top, const int bot)
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  __shared__ uint32_t s_cnt[3][256];
  __shared__ uint32_t s_pos[3][256];
  __shared__ unsigned long long s_best[8];
  __shared__ uint32_t s_bestBin[8];
  __shared__ uint32_t s_mode[3];
  const int t = threadIdx.x;
  const int frame = frameIdx[blockIdx.x];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  for (int i = t; i < 3 * 256; i += blockDim.x)
  {
    (&s_cnt[0][0])[i] = 0u;
    (&s_pos[0][0])[i] = 0u;
  }
  __syncthreads();

  // window: left < col < right, top < row < bot (strict, :153-156); bounds are the reference's uint16 values
  const int c0 = max(left + 1, 0), c1 = min(right - 1, g.width - 1);
  const int r0 = max(top + 1, 0),  r1 = min(bot - 1, g.height - 1);
  if (c0 <= c1 && r0 <= r1)
  {
    const int p0 = c0 >> 1, p1 = c1 >> 1;                  // pixel pairs touching the window
    const int pairs = p1 - p0 + 1;
    const int total = pairs * (r1 - r0 + 1);
    const uint8_t* base = frames + (size_t)frame * g.frameStride;
    for (int i = t; i < total; i += blockDim.x)
    {
      const int row = r0 + i / pairs, pr = p0 + i % pairs;
      const uint32_t w = *reinterpret_cast<const uint32_t*>(base + (size_t)row * g.lineLength + (size_t)pr * 4u);
      uint32_t hsv[2];
      hsv_pair(w & 0x00FF00FFu, w, coef_yuyv(), s_lutHue, s_lut255, hsv[0], hsv[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int col = 2 * pr + e;
        if (col < c0 || col > c1)
          continue;
        const uint32_t pos = (uint32_t)row * (uint32_t)g.width + (uint32_t)col;
        const uint32_t h = hsv[e] & 0xFFu, s = (hsv[e] >> 8) & 0xFFu, v = hsv[e] >> 16;
        atomicAdd(&s_cnt[0][h], 1u); atomicMax(&s_pos[0][h], pos);
        atomicAdd(&s_cnt[1][s], 1u); atomicMax(&s_pos[1][s], pos);
        atomicAdd(&s_cnt[2][v], 1u); atomicMax(&s_pos[2][v], pos);
      }
    }
  }
  __syncthreads();
  for (int c = 0; c < 3; ++c)                              // mode = first bin to reach the final maximum (:167-178)
  {
    unsigned long long best = ((unsigned long long)s_cnt[c][t] << 32) | (unsigned long long)(0xFFFFFFFFu - s_pos[c][t]);
    uint32_t bestBin = (uint32_t)t;
    for (int off = 16; off > 0; off >>= 1)
    {
      const unsigned long long ok = __shfl_down_sync(0xFFFFFFFFu, best, off);
      const uint32_t ob = __shfl_down_sync(0xFFFFFFFFu, bestBin, off);
      if (ok > best || (ok == best && ob < bestBin)) { best = ok; bestBin = ob; }
    }
    if ((t & 31) == 0) { s_best[t >> 5] = best; s_bestBin[t >> 5] = bestBin; }
    __syncthreads();
    if (t == 0)
    {
      for (int w2 = 1; w2 < 8; ++w2)
        if (s_best[w2] > best || (s_best[w2] == best && s_bestBin[w2] < bestBin)) { best = s_best[w2]; bestBin = s_bestBin[w2]; }
      s_mode[c] = ((best >> 32) == 0ull) ? 0u : bestBin;   // nothing counted: m_maxHue stays 0
    }
    __syncthreads();
  }
  if (t == 0)
  {
    TargetOut* o = out + frame;
    o->detectHue = (uint16_t)((double)s_mode[0] * (double)1.4f);          // :193-198
    o->detectHueTolerance = 15;
    o->detectSat = (uint16_t)((double)s_mode[1] * (double)0.39f);
    o->detectSatTolerance = 30;
    o->detectVal = (uint16_t)((double)s_mode[2] * (double)0.39f);
    o->detectValTolerance = 30;
  }
}

// ---------------------------------------------------------------------------------------------
// WL / OL / OO: +1 / -2 histogram with ordered running maximum
// ---------------------------------------------------------------------------------------------
struct DetectWindow {
  int posL, posR, posT, posB;      // positive: posL < col < posR (and posT < row < posB for OO)
  int negL, negR, negT, negB;      // negative: col < negL or col > negR (or row < negT or row > negB for OO)
  int useRows;                     // 0: line sensors (columns only)
};

// result record per frame: hist[BINS] int32, then seedBin, seedValue
template <int KIND>
__global__ void __launch_bounds__(256)
ordered_hist_kernel(const Geometry g, const uint8_t* __restrict__ frames, const int* __restrict__ frameIdx,
                    int32_t* __restrict__ results, const DetectWindow win)
{
  constexpr bool OO = (KIND == KIND_OO);
  constexpr bool PLANAR = (KIND != KIND_WL);
  constexpr int BINS = OO ? 1024 : 256;
  constexpr int PER = BINS / 256;                          // bins per thread in the update step
  __shared__ HueLutEntry s_lutHue[OO ? 256 : 1];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  __shared__ int32_t  s_base[BINS];
  __shared__ uint32_t s_nl[BINS], s_p[BINS], s_nr[BINS], s_last[BINS];
  __shared__ long long s_best[8];
  __shared__ int s_seedBin, s_seedVal;

  const int t = threadIdx.x;
  const int frame = frameIdx[blockIdx.x];
  if (OO) { fill_div_luts(s_lut43, s_lut255); fill_hue_lut(s_lutHue); }
  for (int i = t; i < BINS; i += 256)
  {
    s_base[i] = 0; s_nl[i] = 0; s_p[i] = 0; s_nr[i] = 0; s_last[i] = 0;
  }
  if (t == 0) { s_seedBin = 0; s_seedVal = 0; }
  __syncthreads();

  const int W = g.width, H = g.height;
  const uint8_t* base = frames + (size_t)frame * g.frameStride;
  const size_t chromaOfs = (size_t)H * g.lineLength;
  const int pairsPerRow = W / 2;

  for (int row = 0; row < H; ++row)
  {
    const bool rowNegAll = win.useRows && (row < win.negT || row > win.negB);
    const bool rowPos = !win.useRows || (win.posT < row && row < win.posB);
    for (int pr = t; pr < pairsPerRow; pr += 256)
    {
      uint32_t yy, cw;
      ChromaCoef coef = coef_yuyv();
      if (!PLANAR)
      {
        cw = *reinterpret_cast<const uint32_t*>(base + (size_t)row * g.lineLength + (size_t)pr * 4u);
        yy = cw & 0x00FF00FFu;
      }
      else
      {
        const uint32_t l2 = *reinterpret_cast<const uint16_t*>(base + (size_t)row * g.lineLength + (size_t)pr * 2u);
        const uint32_t c2 = *reinterpret_cast<const uint16_t*>(base + chromaOfs + (size_t)row * g.lineLength + (size_t)pr * 2u);
        yy = (l2 & 0xFFu) | ((l2 >> 8) << 16);
        cw = c2;                                           // [V U 0 0]
        coef = coef_planar0();
      }
      uint32_t bin[2];
      if (OO)
      {
        uint32_t h0, h1;
        hsv_pair(yy, cw, coef, s_lutHue, s_lut255, h0, h1);
        bin[0] = (((h0 & 0xFFu) >> 3) << 5) | (((h0 >> 8) & 0xFFu) >> 3);
        bin[1] = (((h1 & 0xFFu) >> 3) << 5) | (((h1 >> 8) & 0xFFu) >> 3);
      }
      else
      {
        uint32_t kr, kg, kb;
        rgb_keys(yy, cw, coef, kr, kg, kb);
        const uint32_t v2 = chan8_from_key(__vimax3_u16x2(kr, kg, kb));   // V = max(R,G,B)
        bin[0] = v2 & 0xFFFFu;
        bin[1] = v2 >> 16;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int col = 2 * pr + e;
        if (rowNegAll)
          atomicAdd(&s_nl[bin[e]], 1u);
        else if (rowPos && win.posL < col && col < win.posR)
        {
          atomicAdd(&s_p[bin[e]], 1u);
          atomicMax(&s_last[bin[e]], (uint32_t)col);
        }
        else if (col < win.negL || col > win.negR)
        {
          // a negative pixel is applied before or after this row's positives according to where it
          // sits relative to the positive band (robust to the reference's uint16 bound wrap on narrow images)
          if (col <= win.posL || !rowPos)
            atomicAdd(&s_nl[bin[e]], 1u);
          else
            atomicAdd(&s_nr[bin[e]], 1u);
        }
      }
    }
    __syncthreads();

    // per-bin update + row arg-max (largest peak, then smallest last column)
    long long best = -1;                                    // key = peak << 20 | (0xFFFFF - lastcol); peak > 0 only
#pragma unroll
    for (int j = 0; j < PER; ++j)
    {
      const int b = t + j * 256;
      const int32_t nl = (int32_t)s_nl[b], p = (int32_t)s_p[b], nr = (int32_t)s_nr[b];
      if ((nl | p | nr) != 0)
      {
        const int32_t before = s_base[b] - 2 * nl;
        if (p > 0)
        {
          const int32_t peak = before + p;
          if (peak > 0)
          {
            const long long key = ((long long)peak << 32) | ((long long)(0xFFFFFu - s_last[b]) << 12) | (long long)(0xFFF - b);
            if (key > best) best = key;
          }
        }
        s_base[b] = before + p - 2 * nr;
        s_nl[b] = 0; s_p[b] = 0; s_nr[b] = 0; s_last[b] = 0;
      }
    }
    if (rowPos && !rowNegAll)
    {
      for (int off = 16; off > 0; off >>= 1)
      {
        const long long o = __shfl_down_sync(0xFFFFFFFFu, best, off);
        if (o > best) best = o;
      }
      if ((t & 31) == 0) s_best[t >> 5] = best;
      __syncthreads();
      if (t == 0)
      {
        for (int w2 = 1; w2 < 8; ++w2)
          if (s_best[w2] > best) best = s_best[w2];
        if (best >= 0)
        {
          const int peak = (int)(best >> 32);
          if (peak > s_seedVal)                              // strict: the first bin to EXCEED the running maximum
          {
            s_seedVal = peak;
            s_seedBin = 0xFFF - (int)(best & 0xFFF);
          }
        }
      }
    }
    __syncthreads();
  }

  int32_t* res = results + (size_t)blockIdx.x * (BINS + 2);
  for (int i = t; i < BINS; i += 256)
    res[i] = s_base[i];
  if (t == 0)
  {
    res[BINS] = s_seedBin;
    res[BINS + 1] = s_seedVal;
  }
}

cudaError_t launch_wo_detect(const Geometry& g, int numFlagged, const uint8_t* frames, const int* frameIdx,
                             TargetOut* out, cudaStream_t stream)
{
  if (numFlagged <= 0)
    return cudaSuccess;
  // initImg(): uint16 arithmetic (webcam/object_sensor/include/internal/cv_hsv_range_detector.hpp:84-105)
  const uint16_t hH = (uint16_t)(g.height / 2), hW = (uint16_t)(g.width / 2), step = (uint16_t)(g.height / 6);
  const int left = (uint16_t)(hW - step), right = (uint16_t)(hW + step);
  const int top = (uint16_t)(hH - step), bot = (uint16_t)(hH + step);
  wo_detect_kernel<<<numFlagged, 256, 0, stream>>>(g, frames, frameIdx, out, left, right, top, bot);
  ++g_launches_detect;
  return cudaGetLastError();
}

cudaError_t launch_ordered_hist(int kind, const Geometry& g, int numFlagged, const uint8_t* frames, const int* frameIdx,
                                int32_t* results, cudaStream_t stream)
{
  if (numFlagged <= 0)
    return cudaSuccess;
  DetectWindow w{};
  if (kind == KIND_OO)
  {
    // ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:156-175 (int arithmetic, zone scale 6)
    const int hH = g.height / 2, hW = g.width / 2, step = g.height / 6;
    w.posL = hW - step; w.posR = hW + step; w.posT = hH - step; w.posB = hH + step;
    w.negL = hW - 2 * step; w.negR = hW + 2 * step; w.negT = hH - 2 * step; w.negB = hH + 2 * step;
    w.useRows = 1;
    ordered_hist_kernel<KIND_OO><<<numFlagged, 256, 0, stream>>>(g, frames, frameIdx, results, w);
  }
  else
  {
    // webcam/line_sensor/include/internal/cv_hsv_range_detector.hpp:161-184, step = 40, uint16 arithmetic
    const uint16_t hW = (uint16_t)(g.width / 2);
    const uint16_t lp = (uint16_t)(hW - 40), rp = (uint16_t)(hW + 40);
    w.posL = lp; w.posR = rp;
    w.negL = (uint16_t)(lp - 40); w.negR = (uint16_t)(rp + 40);
    w.useRows = 0;
    if (kind == KIND_WL)
      ordered_hist_kernel<KIND_WL><<<numFlagged, 256, 0, stream>>>(g, frames, frameIdx, results, w);
    else
      ordered_hist_kernel<KIND_OL><<<numFlagged, 256, 0, stream>>>(g, frames, frameIdx, results, w);
  }
  ++g_launches_detect;
  return cudaGetLastError();
}

} // namespace trikb200
