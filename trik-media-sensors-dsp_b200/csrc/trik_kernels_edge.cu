// trik_kernels_edge.cu -- ov7670/edge_line_sensor: Sobel 3x3 on the luma plane, threshold, column
// centroid of the edge pixels.
//
// Reference: ov7670/edge_line_sensor/include/internal/cv_ball_detector_seqpass.hpp
//   :176-182  IMG_sobel_3x3_8 + IMG_thr_gt2max_8(.., 50) on the luma plane          (TI IMGLIB: closed, not vendored)
//   :186-205  count / column-sum of the pixels == 0xFF over columns 16 .. W-16, per row in uint16_t
//   :386-414  targetX = centroid column, targetY from a sum that is never added to (-100 whenever anything is found),
//             targetSize = ceil(sqrt(points / pi)) * 400 / (W + H)
// PARITY UNPINNED for the two IMGLIB kernels: they are restated from TI's published natural-C models (oracle/imglib_open.c),
// and the reference's own sensor code is built against that restatement as the checker (oracle/_ref/libtrikref_oe.so).
//
// What the IMGLIB pair amounts to for the pixels the sensor looks at: IMG_sobel_3x3_8 treats the image as one raster line
// and writes output i + 1 from the 3x3 block whose top-left input is i, so output (r, c) is the Sobel magnitude centred on
// input (r + 1, c); its wrap-around garbage lands in columns 0 and W-1, which the sensor's 16-column margins never read,
// and the last two rows of the work buffer are never written (zero for ever).  "== 0xFF after gt2max(50)" is |H| + |V| > 50.
// So: points = #{(r, c): 0 <= r <= H-3, 16 <= c <= W-16, |H| + |V| > 50 at (r+1, c)} -- no intermediate image at all.
//
// Kernel: one CTA per frame, one thread per column walking down the rows with a three-row sliding window of the separable
// parts (A = l + 2m + r, D = r - l): one byte load per row and thread, the neighbours' bytes by shuffle.  A thread's
// column never changes, so its column sum is count * column -- except that the reference sums each ROW in uint16_t, which
// wraps from W = 384 on; then (WRAP) the row sums are formed exactly, per row, in shared memory.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_edge;
std::atomic<long long> g_launches_edge{0};

template <bool WRAP>
__global__ void __launch_bounds__(1024)
edge_line_kernel(const uint8_t* __restrict__ frames, const long long frameStride, const int lineLength,
                 const int W, const int H, TargetOut* __restrict__ out, const int outStride)
{
  extern __shared__ uint32_t s_rows[];                     // WRAP: column sum of every row
  __shared__ uint32_t s_cnt[32], s_sx[32];
  const int frame = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = (int)(blockDim.x >> 5);
  const int c = 16 + t;                                    // this thread's column
  const bool counted = c <= W - 16;
  const uint8_t* base = frames + (long long)frame * frameStride;
  if (WRAP)
  {
    for (int i = t; i < H; i += blockDim.x) s_rows[i] = 0u;
    __syncthreads();
  }
  // row k of this thread: its own byte, the left neighbour (lane - 1; lane 0 loads column c - 1 itself) and the right one
  // (lane + 1; lane 31 loads c + 1).  (Fetching eight rows ahead of the combine step was measured: slower, 0.60 against
  // 0.47 ms per 4096 x 320x240 -- the walk is bound by its ~25 instructions per row and warp, not by load latency.)
  auto load_row = [&](int k, int& A, int& D)
  {
    const uint8_t* rowp = base + (size_t)k * lineLength;
    const int m = c < W ? (int)__ldg(rowp + c) : 0;
    int l = __shfl_up_sync(0xFFFFFFFFu, m, 1);
    int r = __shfl_down_sync(0xFFFFFFFFu, m, 1);
    if (lane == 0) l = c - 1 < W ? (int)__ldg(rowp + c - 1) : 0;   // c >= 16
    if (lane == 31) r = c + 1 < W ? (int)__ldg(rowp + c + 1) : 0;
    A = l + 2 * m + r;
    D = r - l;
  };
  uint32_t cnt = 0u;
  if (H >= 3)
  {
    int A0, D0, A1, D1;
    load_row(0, A0, D0);
    load_row(1, A1, D1);
    for (int r = 0; r + 2 < H; ++r)
    {
      int A2, D2;
      load_row(r + 2, A2, D2);
      const int Hs = A2 - A0, Vs = D0 + 2 * D1 + D2;
      const bool det = counted && (abs(Hs) + abs(Vs) > 50);
      cnt += det ? 1u : 0u;
      if (WRAP)
      {
        const uint32_t rowPart = __reduce_add_sync(0xFFFFFFFFu, det ? (uint32_t)c : 0u);
        if (lane == 0 && rowPart) atomicAdd(&s_rows[r], rowPart);
      }
      A0 = A1; D0 = D1; A1 = A2; D1 = D2;
    }
  }
  uint32_t sx = cnt * (uint32_t)c;
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  sx = __reduce_add_sync(0xFFFFFFFFu, sx);
  if (lane == 0) { s_cnt[warp] = cnt; s_sx[warp] = sx; }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t points = lane < nwarps ? s_cnt[lane] : 0u;
    uint32_t tx = lane < nwarps ? s_sx[lane] : 0u;
    points = __reduce_add_sync(0xFFFFFFFFu, points);
    tx = __reduce_add_sync(0xFFFFFFFFu, tx);
    if (WRAP)
    {
      uint32_t w = 0u;
      for (int i = lane; i < H; i += 32) w += s_rows[i] & 0xFFFFu;            // each row's sum as the reference's uint16_t
      tx = __reduce_add_sync(0xFFFFFFFFu, w);
    }
    if (lane == 0)
    {
      TargetOut o;
      o.targetX = 0; o.targetY = 0; o.targetSize = 0; o.pad = 0;
      o.detectHue = o.detectHueTolerance = o.detectSat = o.detectSatTolerance = o.detectVal = o.detectValTolerance = 0;
      if (points > 0u)                                                        // :388-403
      {
        const int32_t targetX = (int32_t)(tx / points);
        const uint32_t radius = (uint32_t)ceilf(sqrtf((float)points / 3.1415927f));
        o.targetX = (int8_t)(((targetX - W / 2) * 100 * 2) / W);
        o.targetY = (int8_t)(((0 - H / 2) * 100 * 2) / H);                    // m_targetY is never accumulated
        o.targetSize = (uint8_t)((radius * 100u * 4u) / (uint32_t)(W + H));
      }
      *reinterpret_cast<TargetOut*>(reinterpret_cast<uint8_t*>(out) + (size_t)frame * outStride) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// packed form: four pixels (one aligned luma word) per thread, even / odd bytes as two 16-bit lanes per register
// ---------------------------------------------------------------------------------------------
// Per lane everything stays non-negative through biases (D + 256, V and H + 1024), and |x - 1024| = max(x, 2048 - x) - 1024,
// so "|H| + |V| > 50" is  max(Hb, 2048 - Hb) + max(Vb, 2048 - Vb) > 2098  on packed lanes.  A thread's four columns never
// change: it keeps four 16-bit hit counters (rows < 65536) and the column sums follow at the end.  Threads are laid out as
// `groups` row bands x `wpr` words, a band per group of threads, so that a frame offers several thousand threads.
__device__ __forceinline__ uint32_t edge_hits2(uint32_t A0, uint32_t A2, uint32_t D0, uint32_t D1, uint32_t D2)
{
  const uint32_t Hb = A2 + 0x04000400u - A0;                    // lanes in [4, 2044]
  const uint32_t Vb = D0 + 2u * D1 + D2;                        // three biased differences + one more: bias 4 * 256
  const uint32_t aH = __vmaxu2(Hb, 0x08000800u - Hb);
  const uint32_t aV = __vmaxu2(Vb, 0x08000800u - Vb);
  // lane sum = 2048 + |H| + |V| <= 4088: adding 32768 - 2099 carries into bit 15 exactly when |H| + |V| > 50
  return ((aH + aV + 0x77CD77CDu) >> 15) & 0x00010001u;
}

template <bool WRAP>
__global__ void __launch_bounds__(1024)
edge_line4_kernel(const uint8_t* __restrict__ frames, const long long frameStride, const int lineLength,
                  const int W, const int H, TargetOut* __restrict__ out, const int outStride, const int wpr, const int groups)
{
  extern __shared__ uint32_t s_rows[];                     // WRAP: column sum of every row
  __shared__ uint32_t s_cnt[32], s_sx[32];
  const int frame = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = (int)(blockDim.x >> 5);
  const int grp = t / wpr, wi = t - grp * wpr;              // wpr is a multiple of 32: a warp never straddles two bands
  // a warp covers 30 words; its lanes 0 and 31 only carry the words next to them (no extra loads, no special cases)
  const int j = 4 + (wi >> 5) * 30 + lane - 1;              // luma word index: columns 4j .. 4j+3, first counted word is 4
  const int lastWord = (W - 16) >> 2;                       // holds column W-16, its only counted one
  const bool loads = grp < groups && j >= 3 && 4 * j < W;
  const bool live = grp < groups && lane != 0 && lane != 31 && j <= lastWord;
  const uint8_t* base = frames + (long long)frame * frameStride;
  if (WRAP)
  {
    for (int i = t; i < H; i += blockDim.x) s_rows[i] = 0u;
    __syncthreads();
  }
  // output rows of this band: [rBeg, rEnd), from input rows rBeg .. rEnd+1
  const int outRows = H - 2;
  const int per = (outRows + groups - 1) / groups;
  const int rBeg = min(grp * per, outRows), rEnd = min(rBeg + per, outRows);
  // per input row: the separable parts of the four pixels as packed lanes (lo = bytes 0,2; hi = bytes 1,3)
  auto load_row = [&](int k, uint32_t& Alo, uint32_t& Ahi, uint32_t& Dlo, uint32_t& Dhi)
  {
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(base + (size_t)k * lineLength);
    const uint32_t m = loads ? __ldg(rowp + j) : 0u;
    const uint32_t lw = __shfl_up_sync(0xFFFFFFFFu, m, 1);
    const uint32_t rw = __shfl_down_sync(0xFFFFFFFFu, m, 1);
    const uint32_t l = __byte_perm(lw, m, 0x6543);            // bytes of columns 4j-1 .. 4j+2
    const uint32_t r = __byte_perm(m, rw, 0x4321);            // bytes of columns 4j+1 .. 4j+4
    const uint32_t llo = l & 0x00FF00FFu, lhi = (l >> 8) & 0x00FF00FFu;
    const uint32_t mlo = m & 0x00FF00FFu, mhi = (m >> 8) & 0x00FF00FFu;
    const uint32_t rlo = r & 0x00FF00FFu, rhi = (r >> 8) & 0x00FF00FFu;
    Alo = llo + 2u * mlo + rlo;  Ahi = lhi + 2u * mhi + rhi;
    Dlo = rlo + 0x01000100u - llo;  Dhi = rhi + 0x01000100u - lhi;
  };
  uint32_t hitsLo = 0u, hitsHi = 0u;                        // 16-bit counters: bytes 0,2 / bytes 1,3
  if (rBeg < rEnd)
  {
    uint32_t A0l, A0h, D0l, D0h, A1l, A1h, D1l, D1h;
    load_row(rBeg, A0l, A0h, D0l, D0h);
    load_row(rBeg + 1, A1l, A1h, D1l, D1h);
    const uint32_t keepLo = j == lastWord ? 0x0000FFFFu : 0xFFFFFFFFu;         // the last word counts column W-16 only
    const uint32_t keepHi = j == lastWord ? 0u : 0xFFFFFFFFu;
    for (int r = rBeg; r < rEnd; ++r)
    {
      uint32_t A2l, A2h, D2l, D2h;
      load_row(r + 2, A2l, A2h, D2l, D2h);
      const uint32_t hl = edge_hits2(A0l, A2l, D0l, D1l, D2l) & keepLo;
      const uint32_t hh = edge_hits2(A0h, A2h, D0h, D1h, D2h) & keepHi;
      hitsLo += live ? hl : 0u; hitsHi += live ? hh : 0u;
      if (WRAP)
      {
        const uint32_t c0 = 4u * (uint32_t)j;
        const uint32_t part = live ? ((hl & 1u) * c0 + (hl >> 16) * (c0 + 2u) + (hh & 1u) * (c0 + 1u) + (hh >> 16) * (c0 + 3u)) : 0u;
        const uint32_t rowPart = __reduce_add_sync(0xFFFFFFFFu, part);
        if (lane == 0 && rowPart) atomicAdd(&s_rows[r], rowPart);
      }
      A0l = A1l; A0h = A1h; D0l = D1l; D0h = D1h; A1l = A2l; A1h = A2h; D1l = D2l; D1h = D2h;
    }
  }
  uint32_t cnt = 0u, sx = 0u;
  if (live)
  {
    const uint32_t n0 = hitsLo & 0xFFFFu, n2 = hitsLo >> 16, n1 = hitsHi & 0xFFFFu, n3 = hitsHi >> 16, c0 = 4u * (uint32_t)j;
    cnt = n0 + n1 + n2 + n3;
    sx = cnt * c0 + n1 + 2u * n2 + 3u * n3;
  }
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  sx = __reduce_add_sync(0xFFFFFFFFu, sx);
  if (lane == 0) { s_cnt[warp] = cnt; s_sx[warp] = sx; }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t points = lane < nwarps ? s_cnt[lane] : 0u;
    uint32_t tx = lane < nwarps ? s_sx[lane] : 0u;
    points = __reduce_add_sync(0xFFFFFFFFu, points);
    tx = __reduce_add_sync(0xFFFFFFFFu, tx);
    if (WRAP)
    {
      uint32_t w = 0u;
      for (int i = lane; i < H; i += 32) w += s_rows[i] & 0xFFFFu;            // each row's sum as the reference's uint16_t
      tx = __reduce_add_sync(0xFFFFFFFFu, w);
    }
    if (lane == 0)
    {
      TargetOut o;
      o.targetX = 0; o.targetY = 0; o.targetSize = 0; o.pad = 0;
      o.detectHue = o.detectHueTolerance = o.detectSat = o.detectSatTolerance = o.detectVal = o.detectValTolerance = 0;
      if (points > 0u)                                                        // :388-403
      {
        const int32_t targetX = (int32_t)(tx / points);
        const uint32_t radius = (uint32_t)ceilf(sqrtf((float)points / 3.1415927f));
        o.targetX = (int8_t)(((targetX - W / 2) * 100 * 2) / W);
        o.targetY = (int8_t)(((0 - H / 2) * 100 * 2) / H);                    // m_targetY is never accumulated
        o.targetSize = (uint8_t)((radius * 100u * 4u) / (uint32_t)(W + H));
      }
      *reinterpret_cast<TargetOut*>(reinterpret_cast<uint8_t*>(out) + (size_t)frame * outStride) = o;
    }
  }
}

static int g_edgeVariant = 0;        // 0 = packed (when the rows are 4-byte aligned), 1 = one thread per column
void set_edge_variant(int v) { g_edgeVariant = v; }

cudaError_t launch_edge_line(const uint8_t* frames, long long frameStride, int lineLength, int width, int height,
                             int numFrames, TargetOut* out, int outStride, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  if (width < 32 || height < 4 || width % 32 != 0 || height % 4 != 0 || lineLength < width || outStride < (int)sizeof(TargetOut)
      || width - 31 > 1024)
    return cudaErrorInvalidValue;
  // the reference's uint16_t row sums can wrap once sum_{c=16}^{W-16} c exceeds 65535
  const long long maxRowSum = ((long long)(width - 16) * (width - 15) - 16 * 15) / 2;
  const bool aligned4 = (lineLength % 4 == 0) && (frameStride % 4 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 3u) == 0);
  if (g_edgeVariant == 0 && aligned4 && height < 65536)
  {
    const int words = (width >> 2) - 7;                     // words 4 .. (W-16)/4
    const int wpr = ((words + 29) / 30) * 32;               // 30 counted words per warp + two halo lanes
    int groups = 512 / wpr;
    if (groups < 1) groups = 1;
    if (groups > (height - 2 + 7) / 8) groups = (height - 2 + 7) / 8;      // at least 8 output rows per band
    if (groups < 1) groups = 1;
    const int threads = wpr * groups;
    if (maxRowSum > 65535)
      edge_line4_kernel<true><<<numFrames, threads, (size_t)height * sizeof(uint32_t), stream>>>(frames, frameStride, lineLength, width, height, out, outStride, wpr, groups);
    else
      edge_line4_kernel<false><<<numFrames, threads, 0, stream>>>(frames, frameStride, lineLength, width, height, out, outStride, wpr, groups);
    ++g_launches_edge;
    return cudaGetLastError();
  }
  const int cols = width - 31;                              // columns 16 .. W-16
  const int threads = ((cols + 31) / 32) * 32;
  if (maxRowSum > 65535)
    edge_line_kernel<true><<<numFrames, threads, (size_t)height * sizeof(uint32_t), stream>>>(frames, frameStride, lineLength, width, height, out, outStride);
  else
    edge_line_kernel<false><<<numFrames, threads, 0, stream>>>(frames, frameStride, lineLength, width, height, out, outStride);
  ++g_launches_edge;
  return cudaGetLastError();
}

} // namespace trikb200
