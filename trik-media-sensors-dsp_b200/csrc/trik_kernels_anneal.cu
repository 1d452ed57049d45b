// trik_kernels_anneal.cu -- the annealing tail of the auto-calibration on the device.
//
// WL / OL (webcam/line_sensor/include/internal/cv_hsv_range_detector.hpp:95-159,203-250) and OO
// (ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:77-153,180-236) finish their auto-detect with
// 46 x 200 simulated-annealing moves driven by rand() and pow().  The host version (trik_host.cpp: anneal_line /
// anneal_oo, glibc's generator restated, libm pow called) is the default and is what the parity tests pin.  This
// file runs the same chain on the device, one thread per calibrating frame, so that calibration needs no host
// step (TRIKB200_BATCH_DEVICE_TAIL; works with TRIKB200_BATCH_ASYNC and device-resident results):
//   * glibc TYPE_3 rand(): integer arithmetic, identical;
//   * every double operation is an explicit round-to-nearest intrinsic (no FMA contraction), identical;
//   * pow(): CUDA's double pow is within 2 ulp, glibc's within 1 ulp, so single results can differ in the last
//     bit.  A different bit only matters when it moves a value across an integer (the (int) truncation in the
//     increment) or across the integer rand() value it is compared with -- ~1e-13 per call.  tests/test_anneal_gpu.py
//     runs both tails on the same histograms and seeds and reports the mismatching frames (none observed).
// The chain is inherently sequential (each move depends on the previous acceptance), so one frame costs tens of
// milliseconds on one thread; the kernel pays off for batches of calibrating frames, not for one.
#include <atomic>
#include "trik_kernels.cuh"

namespace trikb200 {

std::atomic<long long> g_launches_anneal{0};

namespace {

struct DevRand {                         // glibc 2.39 stdlib/random_r.c, TYPE_3 (degree 31, separation 3)
  int32_t r[31];
  int f, b;
  __device__ void seed(unsigned s)
  {
    if (s == 0) s = 1;
    r[0] = (int32_t)s;
    int32_t word = (int32_t)s;
    for (int i = 1; i < 31; ++i)
    {
      const int32_t hi = word / 127773, lo = word % 127773;
      word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      r[i] = word;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; ++i)
      (void)next();
  }
  __device__ int next()
  {
    const uint32_t val = (uint32_t)r[f] + (uint32_t)r[b];
    r[f] = (int32_t)val;
    if (++f >= 31) f = 0;
    if (++b >= 31) b = 0;
    return (int)(val >> 1);
  }
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

constexpr double kTEnd = 0.0005, kLambda = 0.76, kE = 2.718281828;
constexpr double kRandMax = 2147483647.0;

// (int)(val + ((pow(base, degree) - 1) * t) * (double)(mx - mn)) with alpha = rand / RAND_MAX, degree = 2 alpha - 1
__device__ int increment_value(DevRand& rng, int val, int mn, int mx, double base, double t)
{
  const double alpha = ddiv((double)rng.next(), kRandMax);
  const double degree = dadd(dmul(2.0, alpha), -1.0);
  const double step = dmul(dmul(dadd(pow(base, degree), -1.0), t), (double)(mx - mn));
  return (int)dadd((double)val, step);
}

// do_getIncrement (WL/.../cv_hsv_range_detector.hpp:95-112): rejection sampling, inclusive bounds
__device__ int line_increment(DevRand& rng, int val, int mn, int mx, double base, double t)
{
  for (;;)
  {
    if (mn == mx)
      return mn;
    const int res = increment_value(rng, val, mn, mx, base, t);
    if (!((res < mn) || (res > mx)))
      return res;
  }
}

// getIncrement (OO/.../cv_hsv_range_detector.hpp:77-98): half-open upper bound
__device__ int oo_increment(DevRand& rng, int val, int mn, int mx, double t)
{
  for (;;)
  {
    if (mn == mx)
      return mn;
    const double base = dadd(1.0, ddiv(1.0, t));
    const int res = increment_value(rng, val, mn, mx, base, t);
    if ((mn <= res) && (res < mx))
      return res;
  }
}

__device__ __forceinline__ int truncate_hue(int v) { int r = v % 32; if (r < 0) r += 32; return r; }

} // namespace

// hist: per flagged frame 256 bins + seed bin + seed value (ordered_hist_kernel); prefix sums live in local memory
__global__ void __launch_bounds__(64)
anneal_line_kernel(const int numFlagged, const int* __restrict__ frameIdx, const int32_t* __restrict__ hist,
                   const uint32_t* __restrict__ seeds, const int isOL, TargetOut* __restrict__ out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= numFlagged)
    return;
  const int32_t* h = hist + (size_t)k * 258;
  long long pre[257];                              // F(v0, v1) = pre[v1 + 1] - pre[v0], bins that are 0 count -1 (:133-159)
  pre[0] = 0;
  for (int v = 0; v < 256; ++v)
    pre[v + 1] = pre[v] + (h[v] != 0 ? h[v] : -1);
  DevRand rng;
  rng.seed(seeds[k]);
  int v0 = h[256] & 0xFF, v1 = v0;
  long long L = pre[v1 + 1] - pre[v0];
  double T = 150;
  while (T > kTEnd)
  {
    const double base = dadd(1.0, ddiv(1.0, T));
    for (int i = 0; i < 200; i++)
    {
      const int n0 = line_increment(rng, v0, 0, 255, base, T) & 0xFF;
      const int n1 = line_increment(rng, v1, 0, 255, base, T) & 0xFF;
      const long long newL = n0 <= n1 ? pre[n1 + 1] - pre[n0] : 0;
      if ((double)rng.next() <= dmul(pow(kE, ddiv((double)(newL - L), T)), kRandMax))
      {
        v0 = n0; v1 = n1; L = newL;
      }
    }
    T = dmul(T, kLambda);
  }
  const int o0 = (int)(uint8_t)((float)v0 * 0.39f);
  const int o1 = isOL ? (int)(uint8_t)((float)(v1 + 1) * 0.39f) : (int)(uint8_t)((float)v1 * 0.39f);
  TargetOut* o = out + frameIdx[k];
  o->detectHue = 0; o->detectHueTolerance = 0; o->detectSat = 0; o->detectSatTolerance = 0;
  o->detectVal = (uint16_t)((o1 + o0) / 2);
  o->detectValTolerance = (uint16_t)((o1 - o0) / 2);
}

struct ObjOutRec {                                  // == TRIKB200_ObjOutArgsAlg
  int8_t  t[24];
  uint16_t detectHue, detectHueTolerance, detectSat, detectSatTolerance, detectVal, detectValTolerance;
};

// hist: per flagged frame 32 x 32 hue x saturation cells + seed bin + seed value; summed-area table in local memory
__global__ void __launch_bounds__(64)
anneal_oo_kernel(const int numFlagged, const int* __restrict__ frameIdx, const int32_t* __restrict__ hist,
                 const uint32_t* __restrict__ seeds, ObjOutRec* __restrict__ out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= numFlagged)
    return;
  const int32_t* hs = hist + (size_t)k * 1026;
  long long sat[33 * 33];                          // cells that are 0 count -2 (:109-153)
  for (int i = 0; i <= 32; ++i) { sat[i * 33] = 0; sat[i] = 0; }
  for (int hh = 0; hh < 32; ++hh)
    for (int s = 0; s < 32; ++s)
      sat[(hh + 1) * 33 + s + 1] = sat[hh * 33 + s + 1] + sat[(hh + 1) * 33 + s] - sat[hh * 33 + s]
                                 + (hs[hh * 32 + s] != 0 ? hs[hh * 32 + s] : -2);
  auto rect = [&](int ha, int hb, int s1, int s2) -> long long
  {
    if (ha > hb || s1 > s2) return 0;
    return sat[(hb + 1) * 33 + s2 + 1] - sat[ha * 33 + s2 + 1] - sat[(hb + 1) * 33 + s1] + sat[ha * 33 + s1];
  };
  auto foo = [&](int h1, int h2, int s1, int s2) -> long long
  {
    return h1 <= h2 ? rect(h1, h2, s1, s2) : rect(h1, 31, s1, s2) + rect(0, h2, s1, s2);
  };
  DevRand rng;
  rng.seed(seeds[k]);
  const int seedBin = hs[1024];
  const int sMax = seedBin & 31;
  int h1 = seedBin >> 5, h2 = seedBin >> 5, s1 = sMax, s2 = sMax;
  long long L = foo(h1, h2, s1, s2);
  double T = 150;
  while (T > kTEnd)
  {
    for (int i = 0; i < 200; i++)
    {
      const int h1n = truncate_hue(oo_increment(rng, h1, 0, 32, T));
      const int h2n = truncate_hue(oo_increment(rng, h2, 0, 32, T));
      const int s1n = oo_increment(rng, s1, 0, sMax, T);
      const int s2n = oo_increment(rng, s2, sMax, 32, T);
      const long long Ln = foo(h1n, h2n, s1n, s2n);
      if (L < Ln || ddiv((double)rng.next(), kRandMax) <= pow(kE, ddiv(-(double)(L - Ln), T)))
      {
        h1 = h1n; h2 = h2n; s1 = s1n; s2 = s2n; L = Ln;
      }
    }
    T = dmul(T, kLambda);
  }
  h1 = (int)((float)(h1 << 3) * 1.4f);
  h2 = (int)((float)(((h2 + 1) << 3) - 1) * 1.4f);
  s1 = (int)((float)(s1 << 3) * 0.39f);
  s2 = (int)((float)((s2 + 1) << 3) * 0.39f);
  ObjOutRec* o = out + frameIdx[k];
  if (h1 <= h2)
  {
    o->detectHue = (uint16_t)((h2 + h1) / 2);
    o->detectHueTolerance = (uint16_t)((h2 - h1) / 2);
  }
  else
  {
    const float hue = ((float)h2 - (360.0f - (float)h1)) / 2;
    const float hueTolerance = ((float)h2 + (360.0f - (float)h1)) / 2;
    o->detectHue = (uint16_t)(hue >= 0 ? hue : (hue + 360));
    o->detectHueTolerance = (uint16_t)hueTolerance;
  }
  o->detectSat = (uint16_t)((s2 + s1) / 2);
  o->detectSatTolerance = (uint16_t)((s2 - s1) / 2 + 2);
  o->detectVal = 50;
  o->detectValTolerance = 50;
}

cudaError_t launch_anneal(int kind, int numFlagged, const int* frameIdx, const int32_t* hist, const uint32_t* seeds,
                          void* out, cudaStream_t stream)
{
  if (numFlagged <= 0)
    return cudaSuccess;
  const int threads = 64, grid = (numFlagged + threads - 1) / threads;
  if (kind == KIND_OO)
    anneal_oo_kernel<<<grid, threads, 0, stream>>>(numFlagged, frameIdx, hist, seeds, static_cast<ObjOutRec*>(out));
  else if (kind == KIND_WL || kind == KIND_OL)
    anneal_line_kernel<<<grid, threads, 0, stream>>>(numFlagged, frameIdx, hist, seeds, kind == KIND_OL ? 1 : 0,
                                                     static_cast<TargetOut*>(out));
  else
    return cudaErrorInvalidValue;
  ++g_launches_anneal;
  return cudaGetLastError();
}

} // namespace trikb200
